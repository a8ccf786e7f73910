#!/bin/bash
# 1 GPU: epilogue variants (warps per lane quarter x columns per step): GEMM parity with each build, quick bench each.
tag=${1:-r2l}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for V in s2c16 s3c16 s4c16; do
VMV_LIB=$PWD/videomv_b200/lib/libvmv_$V.so timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest_$V.log 2>&1
echo "pytest $V rc=$?"; tail -1 gpurun_out/${tag}_pytest_$V.log
done
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_base.json 2> gpurun_out/${tag}_bench_base.err
echo "bench s2c32 (base) rc=$?"; cut -c1-150 gpurun_out/${tag}_bench_base.json
for V in s2c16 s3c16 s4c16; do
VMV_LIB=$PWD/videomv_b200/lib/libvmv_$V.so timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_$V.json 2> gpurun_out/${tag}_bench_$V.err
echo "bench $V rc=$?"; cut -c1-150 gpurun_out/${tag}_bench_$V.json; tail -1 gpurun_out/${tag}_bench_$V.err
done
timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_base2.json 2> gpurun_out/${tag}_bench_base2.err
echo "bench s2c32 again rc=$?"; cut -c1-150 gpurun_out/${tag}_bench_base2.json
