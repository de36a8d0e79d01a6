# step time (Madelon, 64 chains) of library variants under build/ (A/B experiments): bash tools/dev/exp_variants.sh lib1.so lib2.so ...
for l in "$@"; do
  echo "== $l"
  BAE_B200_LIB=$PWD/build/$l EXP_CHAINS=64 BAE_TC_PRECONV=1 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
done
echo "== default"; EXP_CHAINS=64 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
