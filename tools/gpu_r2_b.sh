#!/bin/bash
# 1 GPU: new conv modes + full parity suite, quick bench A/B (legacy resample vs new), B=1 / B=2 per-shape breakdowns.
tag=${1:-r2b}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep -aE "passed|failed|^E  |^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -30
VMV_RESAMPLE_LEGACY=1 timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_legacy.json 2> gpurun_out/${tag}_bench_legacy.err
echo "bench legacy rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_legacy.json
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_new.json 2> gpurun_out/${tag}_bench_new.err
echo "bench new rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_new.json; tail -3 gpurun_out/${tag}_bench_new.err
timeout 400 python tools/gemm_breakdown.py 32 --b1 > gpurun_out/${tag}_breakdown_b1.md 2>&1
echo "breakdown b1 rc=$? $(( $(date +%s) - t0 ))s"; grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/${tag}_breakdown_b1.md
timeout 400 python tools/gemm_breakdown.py 32 > gpurun_out/${tag}_breakdown_b2.md 2>&1
echo "breakdown b2 rc=$? $(( $(date +%s) - t0 ))s"; grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/${tag}_breakdown_b2.md
