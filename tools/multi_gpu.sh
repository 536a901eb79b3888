#!/bin/bash
# One gpurun call on N GPUs: the cross-process GPU tests, then bench.py at the given sizes.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/multi_gpu.sh 2 r02 22'
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/multi_gpu.sh 8 r02 22 26'
N=$1; tag=$2; shift 2
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $out/${tag}_${N}gpu_smi.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_${N}gpu_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "nccl or device_side or device_barrier or two_ranks" 2>&1 | tail -15 > $out/${tag}_${N}gpu_pytest.txt
  cat $out/${tag}_${N}gpu_pytest.txt
fi
port=29500
for logn in "$@"; do
  port=$((port+1))
  steps=3; [ "$logn" -ge 24 ] && steps=2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps $steps --warmup 3 --log-n $logn $BENCH_FLAGS > $out/${tag}_bench_${N}gpu_2p${logn}.json 2> $out/${tag}_bench_${N}gpu_2p${logn}.err
  echo "== 2^$logn on $N GPUs: rc=$?"
  tail -c 1500 $out/${tag}_bench_${N}gpu_2p${logn}.err | grep -v "^W0\|^\*\*\*\*\|OMP_NUM_THREADS" | tail -12
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$out/${tag}_bench_${N}gpu_2p${logn}.json") if l.startswith("{")][-1]
    print({k: d.get(k) for k in ("value", "ms_per_step", "verify")}, d.get("e2e", {}) and {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "schedule")}, d.get("breakdown_ms"))
except Exception as e:
    print("no line:", e)
PY
done
