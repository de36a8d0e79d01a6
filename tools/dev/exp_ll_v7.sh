# launch lists of one bae_run(5) call: v7 library against the in-tree build with the images and the lagged order switched off
BAE_B200_LIB=$PWD/build/lib_v7.so BAE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/llv7_launches.csv \
  python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1 > gpurun_out/llv7.log 2>&1
BAE_TC_PRECONV=0 BAE_TC_LAG=0 BAE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/llpc0_launches.csv \
  python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1 > gpurun_out/llpc0.log 2>&1
tail -2 gpurun_out/llv7.log gpurun_out/llpc0.log
