#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integration.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider > gpurun_out/c16_pytest.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
echo done
