#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
timeout 300 python bench.py $S > gpurun_out/c9_var.json 2>> gpurun_out/c9_ab.err
timeout 300 python bench.py $S --lengths full > gpurun_out/c9_full.json 2>> gpurun_out/c9_ab.err
CTCB200_K1F_NST=3 CTCB200_K1F_CPS=3 timeout 300 python bench.py $S --lengths full > gpurun_out/c9_full_33.json 2>> gpurun_out/c9_ab.err
CTCB200_K1F_NT=64 CTCB200_K1F_NST=3 CTCB200_K1F_CPS=4 timeout 300 python bench.py $S --lengths full > gpurun_out/c9_full_nt64.json 2>> gpurun_out/c9_ab.err
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c9_pytest.log 2>&1
echo done
