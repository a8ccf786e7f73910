#!/bin/bash
# 1 GPU: ncu --set full of the hot kernels on their level-0 shapes (tools/prof_kernels.py); report comes back in gpurun_out/.
tag=${1:-r2ncu}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tc2|gn_smem|attention" \
    -o gpurun_out/${tag}_kernels -f python tools/prof_kernels.py > gpurun_out/${tag}_prof.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_prof.log; ls -la gpurun_out/${tag}_kernels.ncu-rep
