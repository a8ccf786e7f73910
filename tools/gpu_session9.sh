#!/bin/bash
tag=${1:-s9}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gemm_breakdown.py > gpurun_out/${tag}_shapes.md 2>&1
echo "shapes rc=$? $(( $(date +%s) - t0 ))s"; head -14 gpurun_out/${tag}_shapes.md; tail -1 gpurun_out/${tag}_shapes.md
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
