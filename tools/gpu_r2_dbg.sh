#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== steps 1"; timeout 300 python tools/debug_loop_graph.py 1 2>&1 | tail -5
echo "== steps 3"; timeout 300 python tools/debug_loop_graph.py 3 2>&1 | tail -5
echo "== steps 1, VMV_PDL=0"; VMV_PDL=0 timeout 300 python tools/debug_loop_graph.py 1 2>&1 | tail -5
echo "== pytest"; timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -k "loop or graph or sampler" 2>&1 | tail -3
