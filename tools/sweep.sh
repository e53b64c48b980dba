#!/bin/bash
# Developer tool (GPU box): sweep ring depth / CTAs per SM / chunk count of the streaming kernels.
run() { timeout 300 python bench.py --lengths ${LEN:-full} --steps 20 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_step']; print('  utt/s',round(d['value']), 'ms',round(d['ms_per_step'],4), 'sweep frac',round(d['roofline']['frac'],3), 'sweep ms',round(r['sweep_ms'],4),'lattice+patch ms',round(r['lattice_plus_patch_ms'],4))"; }
for cfg in "$@"; do echo "$cfg"; env $cfg bash -c "$(declare -f run); run"; done
