# in-box A/B (graph replay, Madelon, 64 chains, swap interval 5): round-2 v7 engine settings against the current defaults, twice
for i in 1 2; do
  for cfg in "BAE_TC_PRECONV=0 BAE_TC_LAG=0" "BAE_TC_HIMG=0 BAE_TC_LAG=0" "BAE_TC_LAG=0" ""; do
    echo "== ${cfg:-defaults}"
    env $cfg EXP_CHAINS=64 EXP_SI=5 timeout 120 python tools/dev/exp_preconv.py madelon child /tmp/x.npz 2>&1 | tail -1
  done
done
