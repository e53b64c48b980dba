#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c3_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench exit $?" >> gpurun_out/c3_bench.err
timeout 300 python tools/head_check.py 6 > gpurun_out/c3_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_head' -c 2 -o gpurun_out/r02_k_head_c2 \
    python tools/head_check.py 6 > gpurun_out/c3_ncu.log 2>&1
echo done
