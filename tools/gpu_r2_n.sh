#!/bin/bash
# Multi-GPU call: parity of the sharded modes vs single GPU (tests/dist_sharded_check.py), then the bench line (headline =
# CFG split x frame sharding, `modes` = pure frames + replicas).  usage: gpu_r2_n.sh <N> <tag> [extra bench args]
N=${1:-2}; tag=${2:-r2n$N}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_sharded_check.py > gpurun_out/${tag}_sharded_check.log 2>&1
echo "sharded check rc=$? $(( $(date +%s) - t0 ))s"; grep -E "^\[sharded\] rank 0|Error|error|Traceback" gpurun_out/${tag}_sharded_check.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 3 --warmup 3 ${3:-} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; grep -a '^{' gpurun_out/${tag}_bench.json | cut -c1-2600; tail -5 gpurun_out/${tag}_bench.err
