#!/bin/bash
# One gpurun call: GPU parity tests, bench A/B (PDL off/on), ncu launch list.  Outputs under gpurun_out/<tag>_*.
tag=${1:-s}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
VMV_PDL=0 timeout 400 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_bench_nopdl.json 2> gpurun_out/${tag}_bench_nopdl.err
echo "bench nopdl rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/${tag}_bench_nopdl.json | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 --shapes-out gpurun_out/${tag}_gemm_shapes.md > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/${tag}_bench.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/prof_forward.py 3 > gpurun_out/${tag}_prof.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 ))s"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.md | head -30
