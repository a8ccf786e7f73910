#!/bin/bash
tag=${1:-n2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep "sharded\]" gpurun_out/${tag}_pytest.log | tail -14; tail -3 gpurun_out/${tag}_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --quick --parallel frames > gpurun_out/${tag}_bench_frames.json 2> gpurun_out/${tag}_bench_frames.err
echo "bench frames rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-900 gpurun_out/${tag}_bench_frames.json; tail -5 gpurun_out/${tag}_bench_frames.err
