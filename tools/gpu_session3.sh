#!/bin/bash
tag=${1:-s3}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_norm_attn_misc_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gn_bench.py 0,1,4,8,16 > gpurun_out/${tag}_gn_bench.md 2>&1
echo "gn rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/${tag}_gn_bench.md
timeout 300 python tools/gemm_breakdown.py > gpurun_out/${tag}_shapes_la.md 2>&1
echo "shapes rc=$? $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/${tag}_shapes_la.md
VMV_GEMM_DEBUG=16 timeout 300 python tools/gemm_breakdown.py > gpurun_out/${tag}_shapes_nola.md 2>&1
echo "shapes rc=$? $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/${tag}_shapes_nola.md
