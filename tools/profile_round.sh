#!/bin/bash
# Round-end evidence (run under gpurun, one GPU): bench lines first (no profiler), then the ncu launch list of the
# same command, then one full capture of the two hot kernels of one step.  Summarise with tools/ncu_summarise.py.
set -x
TAG=${1:-r01e}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
# skip the 3 warm-up steps' launches (3 per step of each kernel), capture the 3 radii of one timed step
ncu --set full --clock-control none --import-source on -k regex:"moments_kernel|dkdn_kernel" -s 18 -c 6 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
cat gpurun_out/${TAG}_bench.json
