#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/c22_pytest.log 2>&1
tail -5 gpurun_out/c22_pytest.log
timeout 900 python bench.py --no-e2e --no-cpu > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1p_sweep|k2_lattice|k3p_patch' -s 9 -c 3 \
    -o gpurun_out/r02_step_var_k1p python bench.py $P > gpurun_out/c22_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1p_sweep' -s 3 -c 1 \
    -o gpurun_out/r02_step_full_k1p python bench.py $P --lengths full > gpurun_out/c22_ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/c22_launches.csv python bench.py $P > gpurun_out/c22_ncu1.log 2>&1
echo done
