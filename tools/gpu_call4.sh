#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c4_pytest.log
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
CTCB200_SWEEP_WARP=0 timeout 300 python bench.py $S > gpurun_out/c4_old_var.json 2>> gpurun_out/c4_ab.err
CTCB200_SWEEP_WARP=0 timeout 300 python bench.py $S --lengths full > gpurun_out/c4_old_full.json 2>> gpurun_out/c4_ab.err
for nw in 3 4 5 6; do
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S > gpurun_out/c4_nw${nw}_var.json 2>> gpurun_out/c4_ab.err
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S --lengths full > gpurun_out/c4_nw${nw}_full.json 2>> gpurun_out/c4_ab.err
done
timeout 300 python tools/launch_list_head.py > gpurun_out/c4_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c4_head_launches.csv \
    python tools/launch_list_head.py > gpurun_out/c4_ncu.log 2>&1
echo done
