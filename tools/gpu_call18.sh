#!/bin/bash
# round-2 evidence: full GPU test suite, the default bench line, ncu launch lists and --set full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c18_pytest.log 2>&1
timeout 900 python bench.py > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
timeout 300 python bench.py --lengths full --no-e2e --no-cpu --no-configs > gpurun_out/c18_bench_full.json 2>> gpurun_out/c18_bench.err
python tools/launch_list_head.py > gpurun_out/c18_head.txt 2>&1
P="--steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity"
timeout 300 python bench.py $P > gpurun_out/c18_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/c18_launches.csv python bench.py $P > gpurun_out/c18_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_lse_gather|k2_lattice|k3p_patch' -s 9 -c 3 \
    -o gpurun_out/r02_step_var python bench.py $P > gpurun_out/c18_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/c18_launches_head.csv python tools/launch_list_head.py > gpurun_out/c18_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_head|k_gemm3' -s 4 -c 4 \
    -o gpurun_out/r02_head python tools/launch_list_head.py > gpurun_out/c18_ncu4.log 2>&1
echo done
