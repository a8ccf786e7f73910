#!/bin/bash
# N GPUs: simulated-rank tests on GPU 0, then sharded parity + fused breakdown + bench.
N=${1:-4}; tag=${2:-r2h$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_peer_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_sharded_check.py > gpurun_out/${tag}_sharded_check.log 2>&1
echo "sharded check rc=$? $(( $(date +%s) - t0 ))s"; grep -a "rank 0" gpurun_out/${tag}_sharded_check.log | sed 's/\[sharded\] rank/\n[sharded] rank/g' | grep -a "rank 0" | cut -c1-330 | tail -12; grep -aE "Error|error|Traceback" gpurun_out/${tag}_sharded_check.log | head -5
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    tools/sharded_breakdown.py > gpurun_out/${tag}_breakdown.md 2> gpurun_out/${tag}_breakdown.err
echo "breakdown rc=$? $(( $(date +%s) - t0 ))s"; grep -av "^\*\*\*\|OMP_NUM" gpurun_out/${tag}_breakdown.md | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 3 --warmup 3 ${3:-} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; grep -a '^{' gpurun_out/${tag}_bench.json | cut -c1-200; tail -3 gpurun_out/${tag}_bench.err
