q="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-verify"
for b in 2 3; do for pf in 0 1; do
  echo "== DP_NTT_BLOCKS=$b DP_NTT_PREFETCH=$pf"
  DP_NTT_BLOCKS=$b DP_NTT_PREFETCH=$pf DP_BENCH_SKIP_ROUNDS=1 $q 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],1),'ntt8n_transform_ms',round(d['roofline_ntt']['transform_ms'],3),'intt_n',round(d['breakdown_ms']['intt_n_total'],2),'sections',d['breakdown_ms']['sections_max_over_ranks'])
"
done; done
