#!/bin/bash
# Everything profiles/ needs for a round, in ONE gpurun call (1 GPU), bounded by timeouts:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_profile.sh r02'
# Produces under gpurun_out/ (copy what should be judged into profiles/):
#   <tag>_pytest_gpu.txt            tail of `pytest -m gpu`
#   <tag>_bench_1gpu.json           the bench line (never taken under a profiler)
#   <tag>_launches_bench.csv        per-launch durations of one bench step (ncu, --clock-control none)
#   <tag>_<kernel>.ncu-rep          one `--set full` capture per hot kernel
# Read the captures here with: python tools/ncu_summary.py gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
quick="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_bench.csv $quick > /dev/null 2>&1
# the four passes of one 2^25 coset transform (launches 0-20 of ntt_tile_kernel are the 2^22 passes of the warm-up step)
timeout 240 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 25 -c 4 -f -o $out/${tag}_ntt_tile_2p25 $quick > /dev/null 2>&1
for k in msm_accumulate_kernel quotient_kernel msm_reduce_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/${tag}_$k $quick > /dev/null 2>&1
done
ls -la $out | tail -20
