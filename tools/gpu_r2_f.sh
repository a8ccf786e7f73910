#!/bin/bash
# N GPUs: sharded parity + breakdown + bench (headline + modes).
N=${1:-4}; tag=${2:-r2f$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_sharded_check.py > gpurun_out/${tag}_sharded_check.log 2>&1
echo "sharded check rc=$? $(( $(date +%s) - t0 ))s"; grep -a "rank 0" gpurun_out/${tag}_sharded_check.log | sed 's/\[sharded\] rank/\n[sharded] rank/g' | grep -a "rank 0" | cut -c1-330 | tail -12; grep -aE "Error|error|Traceback" gpurun_out/${tag}_sharded_check.log | head -5
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    tools/sharded_breakdown.py > gpurun_out/${tag}_breakdown.md 2> gpurun_out/${tag}_breakdown.err
echo "breakdown rc=$? $(( $(date +%s) - t0 ))s"; grep -av "^\*\*\*\|OMP_NUM" gpurun_out/${tag}_breakdown.md | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 3 --warmup 3 ${3:-} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; grep -a '^{' gpurun_out/${tag}_bench.json | cut -c1-1300; tail -3 gpurun_out/${tag}_bench.err
