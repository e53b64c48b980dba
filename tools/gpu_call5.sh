#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for nw in 4 5; do
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S > gpurun_out/c5_nw${nw}_var.json 2>> gpurun_out/c5_ab.err
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S --lengths full > gpurun_out/c5_nw${nw}_full.json 2>> gpurun_out/c5_ab.err
done
CTCB200_K1W_NW=4 CTCB200_K1W_NSLOT=5 timeout 300 python bench.py $S --lengths full > gpurun_out/c5_nw4s5_full.json 2>> gpurun_out/c5_ab.err
CTCB200_SWEEP_WARP=0 CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -k "log_space_recursion_forced_c2" -q --tb=long -p no:cacheprovider > gpurun_out/c5_log_old.log 2>&1
CTCB200_SWEEP_WARP=1 CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -k "log_space_recursion_forced_c2" -q --tb=long -p no:cacheprovider > gpurun_out/c5_log_new.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_integration.py tests/test_gpu_head.py -q --tb=short -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo done
