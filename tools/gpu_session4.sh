#!/bin/bash
tag=${1:-s4}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_norm_attn_misc_gpu.py tests/test_gemm_gpu.py::test_dependent_launch_chain -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gn_bench.py 0 > gpurun_out/${tag}_gn_smem.md 2>&1
echo "gn rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/${tag}_gn_smem.md
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest_unet.log 2>&1
echo "pytest unet rc=$? $(( $(date +%s) - t0 ))s"; tail -5 gpurun_out/${tag}_pytest_unet.log
