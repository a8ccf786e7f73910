#!/bin/bash
# 1 GPU: GEMM tests + e2e tests, then quick bench A/B of the A-resident schedule, then the per-shape table.
tag=${1:-r2i}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; grep -aE "passed|failed|^E  |^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -8
VMV_GEMM_ARES=0 timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_off.json 2> gpurun_out/${tag}_bench_off.err
echo "bench ares=0 rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_off.json
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_on.json 2> gpurun_out/${tag}_bench_on.err
echo "bench ares=1 rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_on.json; tail -3 gpurun_out/${tag}_bench_on.err
timeout 400 python tools/gemm_breakdown.py 32 > gpurun_out/${tag}_breakdown_b2.md 2>&1
echo "breakdown rc=$?"; grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/${tag}_breakdown_b2.md; head -14 gpurun_out/${tag}_breakdown_b2.md | tail -12
