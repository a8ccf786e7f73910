#!/bin/bash
# Round-2 evidence run (1 GPU): smoke, GPU test suite, ncu launch list with DRAM traffic of one forward, full bench line,
# quick benches of the other configs, reference arm.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2ev}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep -aE "passed|failed|^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/prof_forward.py 3 > gpurun_out/${tag}_prof.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 ))s"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.md | head -24
python tools/summarize_traffic.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_traffic.json gpurun_out/${tag}_traffic.md
cp gpurun_out/${tag}_traffic.json profiles/r2_traffic.json
timeout 900 python bench.py --steps 3 --warmup 3 --shapes-out gpurun_out/${tag}_gemm_shapes.md > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --steps 2 --warmup 3 --quick --workload i2v256 > gpurun_out/${tag}_bench_i2v256.json 2> gpurun_out/${tag}_bench_i2v256.err
echo "bench i2v rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench_i2v256.json
timeout 600 python bench.py --steps 2 --warmup 3 --quick --workload t2v512 > gpurun_out/${tag}_bench_t2v512.json 2> gpurun_out/${tag}_bench_t2v512.err
echo "bench 512 rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-200 gpurun_out/${tag}_bench_t2v512.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "bench ref rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-700 gpurun_out/${tag}_bench_reference.json
