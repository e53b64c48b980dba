#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/direct_probe.bin random > gpurun_out/c11_direct_probe.txt 2>&1
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity --lengths full"
timeout 300 python bench.py $P > gpurun_out/c11_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_lse_gather|k2_lattice|k3p_patch' -s 9 -c 3 -o gpurun_out/r02_step_full \
    python bench.py $P > gpurun_out/c11_ncu.log 2>&1
CTCB200_SWEEP_DIRECT=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k1d_sweep' -c 6 --csv --log-file gpurun_out/c11_k1d.csv python bench.py $P > gpurun_out/c11_ncu2.log 2>&1
echo done
