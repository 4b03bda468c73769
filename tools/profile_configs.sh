#!/bin/bash
# ncu launch lists of BASELINE.json configs[0], [2], [3] at full size (tools/bench_configs.py), one list per config:
# run plain first, then under ncu with per-launch durations only (cold-cache, serialised: compare shares).
TAG=${1:-r03}
for c in 0 2 3; do
  python tools/bench_configs.py $c > gpurun_out/${TAG}_config$c.json 2> gpurun_out/${TAG}_config$c.err || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_config${c}_launches.csv \
      python tools/bench_configs.py $c > gpurun_out/${TAG}_config${c}_ncu.log 2>&1
done
