#!/bin/bash
# GPU call 1 of round 2: full GPU test suite, the new bench line, the label_keep_l2 A/B, and an ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench exit $?" >> gpurun_out/c1_bench.err
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for keep in 0 1; do
  for len in var full; do
    CTCB200_LABEL_KEEP_L2=$keep timeout 300 python bench.py $S --lengths $len > gpurun_out/c1_keep${keep}_${len}.json 2>> gpurun_out/c1_ab.err
  done
done
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity"
timeout 300 python bench.py $P > gpurun_out/c1_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k0_prep|k1_lse|k2_lattice|k3p_patch|k4_rescale' -c 40 --csv --log-file gpurun_out/c1_ncu_launches.csv \
    python bench.py $P > gpurun_out/c1_ncu.log 2>&1
echo done
