#!/bin/bash
# 1 GPU: GN micro-benchmark sweep (min KB per CTA) at B=2 and B=1, parity of norm tests, quick bench.
tag=${1:-r2c}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_norm_attn_misc_gpu.py tests/test_gemm_gpu.py tests/test_peer_gpu.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gn_bench.py 0,32,48,96 2 > gpurun_out/${tag}_gn_b2.md 2>&1; echo "gn b2 rc=$?"; cat gpurun_out/${tag}_gn_b2.md
timeout 300 python tools/gn_bench.py 0,32,48,96 1 > gpurun_out/${tag}_gn_b1.md 2>&1; echo "gn b1 rc=$?"; cat gpurun_out/${tag}_gn_b1.md
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
