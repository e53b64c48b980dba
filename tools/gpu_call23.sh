#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
timeout 300 python bench.py --lengths full --no-e2e --no-cpu --no-configs > gpurun_out/c23_bench_full.json 2>> gpurun_out/c23_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/c23_bench_ref.json 2>> gpurun_out/c23_bench.err
timeout 300 python tools/k1_context_probe.py > gpurun_out/c23_context.txt 2>&1
timeout 300 python tools/k1_context_probe.py full >> gpurun_out/c23_context.txt 2>&1
echo done
