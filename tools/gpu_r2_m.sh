#!/bin/bash
# 1 GPU: full GPU suite, then same-box A/B vs the previous build (per-shape table + quick bench).
tag=${1:-r2m}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; grep -aE "passed|failed|^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -4
bash tools/gpu_ab.sh ${tag}_ab | head -3
VMV_LIB=$PWD/videomv_b200/lib/libvideomv_b200_prev.so timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_prev.json 2> gpurun_out/${tag}_bench_prev.err
echo "bench prev rc=$?"; cut -c1-150 gpurun_out/${tag}_bench_prev.json
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_new.json 2> gpurun_out/${tag}_bench_new.err
echo "bench new rc=$?"; cut -c1-150 gpurun_out/${tag}_bench_new.json; tail -2 gpurun_out/${tag}_bench_new.err
