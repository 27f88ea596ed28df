#!/bin/bash
# The two Nsight Compute passes of B200_PROFILING.md for the current kernel set (run on the GPU box through gpurun):
#   1. launch list  (gpu__time_duration per launch; cold-cache, serialised: only the SHARES are comparable with the CUDA-event times)
#   2. --set full of one learner step + the stand-alone V-trace kernels -> gpurun_out/${TAG}_full.ncu-rep (read with tools/ncu_summary.py)
set -u
cd "$(dirname "$0")/.."
TAG=${TAG:-r02}
mkdir -p gpurun_out
NCU=${NCU:-/usr/local/cuda/bin/ncu}
SRL_NO_GRAPH=1 timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "capture/" --csv \
    --log-file gpurun_out/${TAG}_ncu_launches.csv python tests/diag/ncu_step.py > gpurun_out/${TAG}_ncu_launches.out 2>&1
echo "launch list: exit $? ($(grep -c gpu__time_duration gpurun_out/${TAG}_ncu_launches.csv) rows)"
SRL_NO_GRAPH=1 timeout 1500 $NCU --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" -f \
    -o gpurun_out/${TAG}_full python tests/diag/ncu_step.py > gpurun_out/${TAG}_ncu_full.out 2>&1
echo "full capture: exit $?"; ls -la gpurun_out/${TAG}_full.ncu-rep 2>/dev/null
