for l in "$@"; do
  echo "== $l PRECONV=0 LAG=0"; BAE_TC_PRECONV=0 BAE_TC_LAG=0 BAE_B200_LIB=$PWD/build/$l EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
done
echo "== in-tree PRECONV=0 LAG=0"; BAE_TC_PRECONV=0 BAE_TC_LAG=0 EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
