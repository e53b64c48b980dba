#!/bin/bash
# final round-2 evidence of the k1p (bulk-store) build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c25_pytest.log 2>&1
tail -3 gpurun_out/c25_pytest.log
timeout 900 python bench.py > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
timeout 300 python bench.py --lengths full --no-e2e --no-cpu --no-configs > gpurun_out/c25_bench_full.json 2>> gpurun_out/c25_bench.err
timeout 300 python tools/k1_context_probe.py > gpurun_out/c25_context.txt 2>&1
timeout 300 python tools/k1_context_probe.py full >> gpurun_out/c25_context.txt 2>&1
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1p_sweep|k2_lattice|k3p_patch' -s 9 -c 3 \
    -o gpurun_out/r02_step_var_final python bench.py $P > gpurun_out/c25_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1p_sweep' -s 3 -c 1 \
    -o gpurun_out/r02_sweep_full_final python bench.py $P --lengths full > gpurun_out/c25_ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/c25_launches.csv python bench.py $P > gpurun_out/c25_ncu1.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
echo done
