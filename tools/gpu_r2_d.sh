#!/bin/bash
# N GPUs: per-family device-time breakdown of one sharded UNet call (tools/sharded_breakdown.py).
N=${1:-4}; tag=${2:-r2d$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    tools/sharded_breakdown.py ${3:-} > gpurun_out/${tag}_breakdown.md 2> gpurun_out/${tag}_breakdown.err
echo "rc=$?"; grep -av "^\*\*\*\|OMP_NUM" gpurun_out/${tag}_breakdown.md | head -60; tail -5 gpurun_out/${tag}_breakdown.err
