#!/bin/bash
# development aid: saliency parity tests + per-stage timings at 10 M points (run under gpurun)
python -m pytest tests/test_gpu_saliency.py tests/test_gpu_edges.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -5
python tools/gpu_dev.py 1e7 2>&1 | grep -v "^gen"
