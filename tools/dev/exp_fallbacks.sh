# fallback engines (converter-only 3xFP16, 3xTF32 per phase, persistent 3xTF32) of the in-tree build against the v7 library, same box
for cfg in "BAE_TC_PRECONV=0" "BAE_TC_TF32=1" "BAE_PHASE_LAUNCH=0"; do
  for lib in "BAE_B200_LIB=$PWD/build/lib_v7.so" "X=1"; do
    echo "== $cfg $lib"; env $cfg $lib BAE_TC_LAG=0 EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  done
done
