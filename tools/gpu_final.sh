#!/bin/bash
# Round-end sanity on one GPU: build check is done on CPU; here smoke(), the full GPU test suite and the default bench line.
tag=${1:-fin}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-700 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
