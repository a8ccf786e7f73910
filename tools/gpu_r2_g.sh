#!/bin/bash
# N GPUs: breakdown with the layout exchange fused into the GEMM epilogues vs as separate kernels.
N=${1:-4}; tag=${2:-r2g$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for fused in 1 0; do
VMV_SHARD_FUSED_EXCHANGE=$fused timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$fused \
    tools/sharded_breakdown.py > gpurun_out/${tag}_breakdown_fused$fused.md 2> gpurun_out/${tag}_breakdown_fused$fused.err
echo "fused=$fused rc=$?"; grep -av "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/${tag}_breakdown_fused$fused.md | head -48
done
