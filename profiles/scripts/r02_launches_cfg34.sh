# ncu launch lists (gpu__time_duration) of cfg3 / cfg4 encoder steps and of the two-layer [64,64] stack at cfg2 / cfg4
summ() { python - "$1" <<'PY'
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].split("::")[-1]
    tot[name] += float(r[-1]); cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(10):
    print(f"{k:40s} launches {cnt[k]:4d}  mean {v / cnt[k] / 1e3:8.2f} us  share {100 * v / s:5.1f} %")
PY
}
for wl in cfg3_10sweep_p32_b8 cfg4_waymo64_pillar0.1_bev1024; do
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --repeats 1 --no-cpu --no-extra-workloads --no-extractor --no-e2e --no-tokens --no-backbone"
  $CMD > gpurun_out/plain_$wl.json 2> gpurun_out/plain_$wl.err || { tail -5 gpurun_out/plain_$wl.err; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_ncu_launches_$wl.csv $CMD > gpurun_out/ncu_l.log 2>&1
  echo "== $wl"; summ gpurun_out/r02_ncu_launches_$wl.csv
done
for wl in cfg2_nuscenes32_b16_pillar0.2_bev512 cfg4_waymo64_pillar0.1_bev1024; do
  CMD="python profiles/scripts/stack_times.py $wl"
  $CMD > gpurun_out/plain_stack_$wl.txt 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_ncu_launches_stack_$wl.csv $CMD > gpurun_out/ncu_l.log 2>&1
  echo "== stack $wl"; summ gpurun_out/r02_ncu_launches_stack_$wl.csv
done
