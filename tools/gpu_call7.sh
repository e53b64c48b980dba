#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="--steps 100 --warmup 5 --no-e2e --no-cpu --no-configs --no-parity"
for nw in 4 6 8; do
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S > gpurun_out/c7_nw${nw}_var.json 2>> gpurun_out/c7_ab.err
  CTCB200_K1W_NW=$nw timeout 300 python bench.py $S --lengths full > gpurun_out/c7_nw${nw}_full.json 2>> gpurun_out/c7_ab.err
done
CTCB200_K1W_NW=8 CTCB200_K1W_NSLOT=4 timeout 300 python bench.py $S --lengths full > gpurun_out/c7_nw8s4_full.json 2>> gpurun_out/c7_ab.err
timeout 300 python tools/repro_log_c4.py > gpurun_out/c7_repro_full.log 2>&1; echo "rc $?" >> gpurun_out/c7_repro_full.log
CTCB200_PDL=0 timeout 300 python tools/repro_log_c4.py > gpurun_out/c7_repro_nopdl.log 2>&1; echo "rc $?" >> gpurun_out/c7_repro_nopdl.log
CTCB200_DEBUG_SKIP_LATTICE=1 timeout 300 python tools/repro_log_c4.py > gpurun_out/c7_repro_skiplat.log 2>&1; echo "rc $?" >> gpurun_out/c7_repro_skiplat.log
CTCB200_SWEEP_WARP=0 timeout 300 python tools/repro_log_c4.py > gpurun_out/c7_repro_oldsweep.log 2>&1; echo "rc $?" >> gpurun_out/c7_repro_oldsweep.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c7_pytest.log 2>&1
echo done
