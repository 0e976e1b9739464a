# stage times of the default workload from bench.py (no cpu legs), N repeats
for i in 1 2 3; do
python bench.py --steps 20 --warmup 5 --no-cpu --no-extra-workloads --no-extractor 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']; print('group_ms %.4f features_ms %.4f scatter_ms %.4f serial %.4f value %.0f' % (s['group_ms'], s['features_ms'], s['scatter_ms'], s['serial_ms_per_step'], d['value']))"
done
