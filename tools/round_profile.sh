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
# the three passes of one 2^25 coset transform (launches 0-20 of ntt_tile_kernel are the 7 x 3 passes of the 2^22 transforms
# of the warm-up step; 21-23 the first 2^25 transform)
timeout 240 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 24 -c 3 -f -o $out/${tag}_ntt_tile_2p25 $quick > /dev/null 2>&1
for k in msm_accumulate_kernel msm_reduce_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/${tag}_$k $quick > /dev/null 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:quotient_kernel -s 1 -c 1 -f -o $out/${tag}_quotient_kernel $quick > /dev/null 2>&1
# BASELINE config 1 (2^20 gates, one GPU) as a bench line
timeout 300 python bench.py --steps 3 --warmup 3 --log-n 20 --no-cpu > $out/${tag}_bench_1gpu_2p20.json 2> /dev/null
ls -la $out | tail -20
