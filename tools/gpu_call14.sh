#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 280 python tools/gemm_check.py > gpurun_out/c14_gemm.txt 2>&1
CTCB200_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_head.py -q --tb=short -p no:cacheprovider > gpurun_out/c14_pytest.log 2>&1
python tools/launch_list_head.py > gpurun_out/c14_head.txt 2>&1
echo done
