#!/bin/bash
tag=${1:-s6}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_norm_attn_misc_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -8 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${tag}_kernels python tools/prof_kernels.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}_kernels.ncu-rep
