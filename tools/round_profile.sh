#!/bin/bash
# Everything profiles/ needs for a round, in ONE gpurun call (1 GPU), bounded by timeouts:
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/round_profile.sh r02'
# Produces under gpurun_out/ (copy what should be judged into profiles/):
#   <tag>_pytest_gpu.txt            tail of `pytest -m gpu`
#   <tag>_bench_1gpu.json           the bench line (never taken under a profiler)
#   <tag>_launches_bench.csv        per-launch durations of one bench step (ncu, --clock-control none)
#   <tag>_<kernel>.ncu-rep          one `--set full` capture per hot kernel
# Read the captures here with: python tools/ncu_summary.py gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
timeout 420 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
for k in msm_accumulate_kernel ntt_tile_kernel quotient_kernel msm_reduce_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${tag}_$k \
        python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
done
# feasibility of batched-affine bucket accumulation (tools/microbench4.cu; build it first with the nvcc line in its header)
[ -x tools/microbench4 ] && timeout 180 ./tools/microbench4 24 > $out/${tag}_microbench_affine.txt 2>&1
ls -la $out | tail -20
