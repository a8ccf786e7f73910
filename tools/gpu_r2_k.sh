#!/bin/bash
tag=${1:-r2k}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_norm_attn_misc_gpu.py -m gpu -q -k "attention" > gpurun_out/${tag}_pytest.log 2>&1
echo "attention pytest rc=$?"; grep -aE "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/${tag}_pytest.log | tail -10
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_all.log 2>&1
echo "full pytest rc=$?"; grep -aE "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/${tag}_pytest_all.log | tail -10
VMV_ATTN_TC=0 timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_off.json 2> gpurun_out/${tag}_bench_off.err
echo "bench mma.sync for short rows rc=$?"; cut -c1-160 gpurun_out/${tag}_bench_off.json
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_on.json 2> gpurun_out/${tag}_bench_on.err
echo "bench all-tcgen05 attention rc=$?"; cut -c1-160 gpurun_out/${tag}_bench_on.json; tail -2 gpurun_out/${tag}_bench_on.err
timeout 400 python tools/gemm_breakdown.py 32 > gpurun_out/${tag}_breakdown_b2.md 2>&1
grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/${tag}_breakdown_b2.md; sed -n '/^attention/,$p' gpurun_out/${tag}_breakdown_b2.md | head -18
