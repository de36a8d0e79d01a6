# in-box A/B (graph replay, Madelon, 64 chains, swap interval 5): the round-2 v7 library (build/lib_v7.so, commit 99fae66) against the in-tree build
for i in 1 2; do
  echo "== v7 library"; BAE_B200_LIB=$PWD/build/lib_v7.so EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  echo "== in-tree, BAE_TC_PRECONV=0 BAE_TC_LAG=0"; BAE_TC_PRECONV=0 BAE_TC_LAG=0 EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  echo "== in-tree, defaults"; EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
done
