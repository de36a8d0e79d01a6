# step time (graph replay, Madelon, 64 chains, swap interval 5) of the in-tree build under environment settings: bash tools/dev/exp_env.sh "A=1" "B=2 C=3" ...
for i in 1 2; do
  for cfg in "$@"; do
    echo "== $cfg"; env $cfg EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  done
done
