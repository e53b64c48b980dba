#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_head.py -q --tb=short -p no:cacheprovider > gpurun_out/c20_pytest_head.log 2>&1
tail -3 gpurun_out/c20_pytest_head.log
CTCB200_HEAD_PRESPLIT=1 timeout 600 python -m pytest tests/test_gpu_head.py -q --tb=short -p no:cacheprovider -k "3xtf32 or gradient_buffer" > gpurun_out/c20_pytest_head_presplit.log 2>&1
tail -2 gpurun_out/c20_pytest_head_presplit.log
echo "--- in-ring"; python tools/launch_list_head.py 2>&1 | tail -2
echo "--- pre-split"; CTCB200_HEAD_PRESPLIT=1 python tools/launch_list_head.py 2>&1 | tail -2
python tools/head_check.py 2>&1 | tail -12
