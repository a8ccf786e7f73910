#!/bin/bash
# N GPUs: bench line only.
N=${1:-2}; tag=${2:-r2nb$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; grep -a '^{' gpurun_out/${tag}_bench.json | cut -c1-260; tail -2 gpurun_out/${tag}_bench.err
