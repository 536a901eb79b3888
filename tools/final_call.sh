#!/bin/bash
# Last GPU call of round 2 (a few minutes of budget), most important first:
#   1. the GPU tests that cover what changed since r02g (MSM digit sort on its own stream, quotient table variant, resident prover)
#   2. A/B of the two new defaults (decides whether they stay): sort stream, quotient table; times of the next-row kernels
#   3. the bench line of the final build
#   4. batched-affine level, second attempt (tools/microbench6.cu)
#   5. ncu --set full captures of the "next"-row kernels, which round 2 had only for quotient_kernel
#   /usr/local/graft/bin/gpurun --timeout 540 -- 'bash tools/final_call.sh r02h'
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
date +%s > $out/${tag}_t0
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_rounds.py tests/test_zzz_gpu_round2.py -m gpu -x -q -k "(msm or commit or quotient or resident or rounds or satisfied or host_schedules) and not full_size" 2>&1 | tail -8 > $out/${tag}_pytest_changed.txt
date +%s > $out/${tag}_t1
timeout 90 python tools/ab_sort_stream.py > $out/${tag}_ab_sort_stream.txt 2>&1
timeout 90 python tools/f_kernels.py > $out/${tag}_f_kernels.txt 2>&1
date +%s > $out/${tag}_t2
timeout 240 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
date +%s > $out/${tag}_t3
timeout 60 ./tools/microbench6 24 > $out/${tag}_microbench_affine2.txt 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'perm_|poly_|quotient_kernel|quotient_inv|g1_decompress' -c 26 -f -o $out/${tag}_f_kernels python tools/f_kernels.py > $out/${tag}_f_kernels_ncu.log 2>&1
date +%s > $out/${tag}_t4
ls -la $out | tail -14
