# bench.py with other pool shapes: lanes x cluster x threads [x CUDA_DEVICE_MAX_CONNECTIONS]
COMMON="--no-cpu-baseline --no-kaplan --no-cufft --stress-recordings 0 --ingest-seconds 0 --steps 24 --warmup 12"
while read -r lanes cluster threads conn extra; do
  [ -z "$lanes" ] && continue
  echo "== lanes $lanes cluster $cluster threads $threads connections $conn $extra"
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python bench.py $COMMON --lanes $lanes --cluster $cluster --threads $threads $extra 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        k = d['roofline']['kernels']['trk_borre_kernel']
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'trk in-region ms', round(k['ms'], 1), 'alone', round(k['alone_ms'], 1), 'agg_frac', round(d['roofline']['aggregate_frac'], 3))
    elif 'Error' in l:
        print(l[:300])
"
done
