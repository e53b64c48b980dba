#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bwd_gemm_probe.py > gpurun_out/c12_bwd_gemm.txt 2>&1
python tools/launch_list_head.py > gpurun_out/c12_head.txt 2>&1
python tools/launch_list_head.py tf32 >> gpurun_out/c12_head.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_parity.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider > gpurun_out/c12_pytest.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo done
