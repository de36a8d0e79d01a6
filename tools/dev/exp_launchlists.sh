# launch lists (ncu gpu__time_duration.sum) of one production bae_run(5) call with and without the operand images
for pc in 0 1; do
BAE_TC_PRECONV=$pc BAE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/pc${pc}_launches.csv \
  python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1 > gpurun_out/pc${pc}_launches.log 2>&1
tail -2 gpurun_out/pc${pc}_launches.log
done
