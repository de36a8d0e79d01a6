#!/bin/bash
# Timing experiments on the tcgen05 engine (debug flags make results wrong; timing only).
for f in ${FLAGS:-0 1 8 16 17 25}; do
  echo "== BAE_TC_DBGFLAGS=$f"
  BAE_TC_DBGFLAGS=$f timeout 120 python tests/dev_trace.py 2>&1 | tail -12
done
