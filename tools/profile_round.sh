#!/bin/bash
# Round-end evidence (run under gpurun, one GPU): bench lines first (no profiler), then the ncu launch list of the
# same command line, then one full capture of the two hot kernels of one step.  Summarise with tools/ncu_summarise.py.
set -x
TAG=${1:-r02}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-configs --strong-points 0"
python bench.py $FAST > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $FAST > gpurun_out/${TAG}_ncu1.log 2>&1
# skip the warm-up launches (4 pipelines x 12 hot launches), capture the 3 radii of one timed step of the hot kernels:
# moments, screening walk and the two refinement passes of dk/dn (the gated fallback returns at once: not captured)
python bench.py $FAST > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"moments_kernel|dkdn_screen_kernel|dk_refine_kernel|dkdn_refine_kernel" -s 48 -c 12 \
    -o gpurun_out/${TAG}_prof -f python bench.py $FAST > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
