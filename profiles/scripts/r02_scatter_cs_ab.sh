for i in 1 2 3; do
for cs in 0 1; do
  if [ $cs = 1 ]; then export PILLARS_SCATTER_CS=1; else unset PILLARS_SCATTER_CS; fi
  python bench.py --steps 20 --warmup 5 --repeats 5 --no-cpu --no-extra-workloads --no-extractor --no-e2e --no-tokens --no-backbone 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']
print('cs=$cs: value %.0f ms_per_step %.4f serial %.4f features_ms %.4f group_ms %.4f scatter_ms %.4f' % (d['value'], d['ms_per_step'], s['serial_ms_per_step'], s['features_ms'], s['group_ms'], s['scatter_ms']))"
done; done
