#!/bin/bash
tag=${1:-s10}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s10_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["clocks"])
for k in ("roofline","breakdown_ms_per_forward_b2","roofline_hbm_norms","attention_tflops"):
    print(k, json.dumps(d.get(k))[:600])
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_norm_attn_misc_gpu.py -m gpu -q -x -k "groupnorm or conv_in" > gpurun_out/${tag}_memcheck_norm.log 2>&1
echo "memcheck norm rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/${tag}_memcheck_norm.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "rowstats or static_weight or geglu or folded" > gpurun_out/${tag}_memcheck_gemm.log 2>&1
echo "memcheck gemm rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/${tag}_memcheck_gemm.log
