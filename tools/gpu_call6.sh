#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/repro_log_c4.py small > gpurun_out/c6_repro_small.log 2>&1; echo "rc $?" >> gpurun_out/c6_repro_small.log
timeout 300 python tools/repro_log_c4.py > gpurun_out/c6_repro_full.log 2>&1; echo "rc $?" >> gpurun_out/c6_repro_full.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/repro_log_c4.py small > gpurun_out/c6_memcheck.log 2>&1; echo "rc $?" >> gpurun_out/c6_memcheck.log
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity --lengths full"
CTCB200_K1W_NW=5 timeout 300 python bench.py $P > gpurun_out/c6_plain.log 2>&1 &&
CTCB200_K1W_NW=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1w_sweep' -s 3 -c 1 -o gpurun_out/r02_k1w_full \
    python bench.py $P > gpurun_out/c6_ncu.log 2>&1
echo done
