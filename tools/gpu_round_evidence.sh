#!/bin/bash
# One GPU-box session that regenerates the evidence of the current build: GPU tests, default bench line, launch list, ncu full-set export.
# Usage (from the repo root, under gpurun): bash tools/gpu_round_evidence.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench rc $?" >> gpurun_out/${tag}_bench.err
BAE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv \
  python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1 > gpurun_out/${tag}_launches.log 2>&1
BAE_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bae_ -c 52 -o /tmp/${tag}_full -f \
  python tools/profile_step.py --chains 8 --rounds 2 --steps-per-round 1 --profile-round 1 > gpurun_out/${tag}_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_full.log
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_bench.err | tail -3; head -c 600 gpurun_out/${tag}_bench.log
