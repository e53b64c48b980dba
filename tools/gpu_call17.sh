#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for pol in 0 1 2; do
  CTCB200_SCRATCH_POLICY=$pol timeout 300 python bench.py $P > gpurun_out/c17_pol${pol}_var.json 2>> gpurun_out/c17.err
  CTCB200_SCRATCH_POLICY=$pol timeout 300 python bench.py $P --lengths full > gpurun_out/c17_pol${pol}_full.json 2>> gpurun_out/c17.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c17_pol*.json')):
    try:
        d=json.load(open(f)); print(f, 'step', round(d['ms_per_step'],4), 'k1', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],4), 'rest', round(d['roofline_step']['lattice_plus_patch_ms'],4))
    except Exception as e: print(f, 'ERR', e)
PY
