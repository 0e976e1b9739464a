#!/bin/bash
# feature-kernel time against the share of chunks handed out through the cursor (PILLARS_WALK_POOL, in 1/256)
for wl in cfg2_nuscenes32_b16_pillar0.2_bev512 cfg3_10sweep_p32_b8 cfg4_waymo64_pillar0.1_bev1024; do
  for pool in ${POOLS:-0 48 96 160 256}; do
    PILLARS_WALK_POOL=$pool python bench.py --workload $wl --steps 10 --warmup 3 --repeats 3 --no-cpu --no-extra-workloads --no-extractor --no-e2e --no-tokens --no-backbone 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']
print('$wl pool $pool/256: features_ms %.4f group_ms %.4f scatter_ms %.4f value %.0f' % (s['features_ms'], s['group_ms'], s['scatter_ms'], d['value']))"
  done
done
