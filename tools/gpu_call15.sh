#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CTCB200_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_head.py -q --tb=short -p no:cacheprovider > gpurun_out/c15_pytest_head.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_head.py > gpurun_out/c15_pytest_all.log 2>&1
python tools/launch_list_head.py > gpurun_out/c15_head.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
echo done
