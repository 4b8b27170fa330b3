#!/bin/bash
# usage: bash tools/multi_gpu_check.sh <N> <tag>: N-GPU bench line (torchrun, as the driver launches it) + the tests that need >= 2 GPUs
N=${1:-2}; tag=${2:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_boundary.py -q -m gpu -k another_gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_${tag}_${N}gpu.err | tail -1 > gpurun_out/bench_${tag}_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_${N}gpu.json'))
print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'allreduce_us', d.get('stats_allreduce_us'), 'episodes', d['episode_stats']['episodes'])"
