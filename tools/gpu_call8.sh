#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CTCB200_DEBUG=0
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for nw in 4 6 8; do
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S > gpurun_out/c8_nw${nw}_var.json 2>> gpurun_out/c8_ab.err
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S --lengths full > gpurun_out/c8_nw${nw}_full.json 2>> gpurun_out/c8_ab.err
done
CTCB200_K1W_NW=8 CTCB200_K1W_NSLOT=4 timeout 300 python bench.py $S --lengths full > gpurun_out/c8_nw8s4_full.json 2>> gpurun_out/c8_ab.err
timeout 300 python tools/repro_log_c4.py > gpurun_out/c8_repro_full.log 2>&1; echo "rc $?" >> gpurun_out/c8_repro_full.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c8_pytest.log 2>&1
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity --lengths full"
timeout 300 python bench.py $P > gpurun_out/c8_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1w_sweep' -s 3 -c 1 -o gpurun_out/r02_k1w_regs_full \
    python bench.py $P > gpurun_out/c8_ncu.log 2>&1
echo done
