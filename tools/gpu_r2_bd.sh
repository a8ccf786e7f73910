#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/gemm_breakdown.py 32 --b1 > gpurun_out/r2bd_b1.md 2>&1; grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/r2bd_b1.md
timeout 300 python tools/gemm_breakdown.py 32 > gpurun_out/r2bd_b2.md 2>&1; grep -aE "gemm total|groupnorm:|attention:|graph replay" gpurun_out/r2bd_b2.md
