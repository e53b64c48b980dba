#!/bin/bash
# multi-GPU measurements: bench at N = $1 ranks (weak scaling), default seeds and --same-seed; plus the NCCL test
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_c5.py tests/test_gpu_integration.py -q --tb=short -p no:cacheprovider > gpurun_out/m${N}_pytest.log 2>&1
fi
timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "rc $?" >> gpurun_out/m${N}_bench.err
timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --per-rank-lengths --no-e2e --no-configs > gpurun_out/m${N}_bench_perrank.json 2> gpurun_out/m${N}_perrank.err
echo "rc $?" >> gpurun_out/m${N}_perrank.err
echo done
