#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python tools/debug_loop_graph.py 1 2>&1 | grep "replay"
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -k "loop or sampler" 2>&1 | tail -2
