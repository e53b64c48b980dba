#!/bin/bash
# Developer tool (GPU box): sweep-kernel-only timing (lattice skipped; results are garbage, timing only).
export CTCB200_DEBUG_SKIP_LATTICE=1 CTCB200_CHUNKS=1
for cfg in "$@"; do echo "$cfg"; env $cfg timeout 300 python bench.py --lengths full --steps 20 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_step']; print('  k0+k1 ms',round(r['unpipelined_forward_ms'],4),' k3 ms',round(r['unpipelined_backward_ms'],4))"; done
