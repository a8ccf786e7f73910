#!/bin/bash
# same-box A/B of two library builds on the per-shape GEMM table: A = videomv_b200/lib/libvideomv_b200_prev.so, B = current
tag=${1:-ab}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2; do
  VMV_LIB=$PWD/videomv_b200/lib/libvideomv_b200_prev.so timeout 300 python tools/gemm_breakdown.py > gpurun_out/${tag}_A$i.md 2>&1
  timeout 300 python tools/gemm_breakdown.py > gpurun_out/${tag}_B$i.md 2>&1
done
python tools/ab_compare.py gpurun_out/${tag}_A1.md,gpurun_out/${tag}_A2.md gpurun_out/${tag}_B1.md,gpurun_out/${tag}_B2.md | tee gpurun_out/${tag}_compare.txt
