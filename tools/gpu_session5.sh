#!/bin/bash
tag=${1:-s5}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -15 gpurun_out/${tag}_pytest.log
VMV_LN_FUSED=0 timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_nolnf.json 2> gpurun_out/${tag}_bench_nolnf.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench_nolnf.json; tail -3 gpurun_out/${tag}_bench_nolnf.err
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
