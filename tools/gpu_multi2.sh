#!/bin/bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
S="--steps 200 --warmup 10 --no-e2e --no-configs --no-parity"
for cps in 32 24 20 16 12; do
  CTCB200_K3P_CPS=$cps timeout 600 $TR bench.py --gpus $N $S 2> gpurun_out/mm${N}_cps${cps}.err | grep '^{' > gpurun_out/mm${N}_cps${cps}.json
done
echo done
