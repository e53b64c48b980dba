#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { env "$@" timeout 200 python bench.py $P 2>> gpurun_out/c21.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$*', 'k1 ms', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],4), 'step', round(d['ms_per_step'],4), d.get('parity',{}).get('pass'))"; }
P="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --lengths full"
run CTCB200_K1F_NT=128
run CTCB200_K1F_NT=256
P="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs"
run CTCB200_K1F_NT=128
run CTCB200_K1F_NT=256
tail -3 gpurun_out/c21.err
