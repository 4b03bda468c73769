#!/bin/bash
# development aid: downsample parity tests + config 3 timings at 50 M points (run under gpurun)
python -m pytest tests/test_gpu_downsample.py tests/test_gpu_compare.py tests/test_gpu_saliency.py -m gpu -x -q 2>&1 | tail -3
MORE3D_BENCH_N3=${1:-5e7} python tools/bench_configs.py 3 2>/dev/null > gpurun_out/c3b.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/c3b.json"))
print("tree", d["tree_build_ms"], {k: (round(v["saliency_downsample_device_ms"], 1), round(v["voxel_ms"], 1),
      v["cells_partition_cloud"], v["all_cell_radii_below_lod_or_leaf"]) for k, v in d["lods"].items()})
PY
