# step time (graph replay, Madelon, 64 chains, swap interval 5) of library variants under build/ and of the in-tree build: bash tools/dev/exp_libs.sh a.so b.so ...
for i in 1 2; do
  for l in "$@"; do
    echo "== $l"; BAE_B200_LIB=$PWD/build/$l EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  done
  echo "== in-tree"; EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
done
