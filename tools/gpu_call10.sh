#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for cps in 4 3 5; do
  CTCB200_K1D_CPS=$cps timeout 300 python bench.py $S > gpurun_out/c10_cps${cps}_var.json 2>> gpurun_out/c10_ab.err
  CTCB200_K1D_CPS=$cps timeout 300 python bench.py $S --lengths full > gpurun_out/c10_cps${cps}_full.json 2>> gpurun_out/c10_ab.err
done
CTCB200_SWEEP_DIRECT=0 timeout 300 python bench.py $S --lengths full > gpurun_out/c10_old_full.json 2>> gpurun_out/c10_ab.err
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c10_pytest.log 2>&1
echo done
