# launch list (ncu gpu__time_duration.sum) of one production bae_run(5) call of the current build -> gpurun_out/$1_launches.csv
tag=${1:-cur}
BAE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv \
  python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1 > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_launches.log
