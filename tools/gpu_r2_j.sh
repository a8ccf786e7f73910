#!/bin/bash
# 1 GPU: epilogue warp split A/B (2 / 3 / 4 warps per TMEM lane quarter): GEMM parity tests with each build, quick bench each.
tag=${1:-r2j}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for S in 3 4; do
VMV_LIB=$PWD/videomv_b200/lib/libvideomv_b200_s$S.so timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest_s$S.log 2>&1
echo "pytest split=$S rc=$?"; tail -2 gpurun_out/${tag}_pytest_s$S.log
done
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_s2.json 2> gpurun_out/${tag}_bench_s2.err
echo "bench split=2 rc=$?"; cut -c1-160 gpurun_out/${tag}_bench_s2.json
for S in 3 4; do
VMV_LIB=$PWD/videomv_b200/lib/libvideomv_b200_s$S.so timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_s$S.json 2> gpurun_out/${tag}_bench_s$S.err
echo "bench split=$S rc=$?"; cut -c1-160 gpurun_out/${tag}_bench_s$S.json; tail -2 gpurun_out/${tag}_bench_s$S.err
done
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_s2b.json 2> gpurun_out/${tag}_bench_s2b.err
echo "bench split=2 again rc=$?"; cut -c1-160 gpurun_out/${tag}_bench_s2b.json
