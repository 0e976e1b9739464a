# ncu launch lists of the late additions: DynamicPillarVFE (cfg2), two-layer stack (cfg4), padded voxels (cfg2)
summ() { python - "$1" <<'PY'
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].split("::")[-1]
    tot[name] += float(r[-1]); cnt[name] += 1
for k, v in tot.most_common(16):
    print(f"{k:44s} launches {cnt[k]:4d}  mean {v / cnt[k] / 1e3:8.2f} us")
PY
}
python profiles/scripts/dyn_times.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_dynamic_cfg2.csv python profiles/scripts/dyn_times.py > gpurun_out/ncu_l.log 2>&1
echo "== DynamicPillarVFE [64] then [64,64], cfg2"; summ gpurun_out/r02_ncu_launches_dynamic_cfg2.csv
python profiles/scripts/stack_times.py cfg4_waymo64_pillar0.1_bev1024 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_ncu_launches_stack_cfg4.csv python profiles/scripts/stack_times.py cfg4_waymo64_pillar0.1_bev1024 > gpurun_out/ncu_l.log 2>&1
echo "== encode_stack [64] (general kernel) then [64,64] (two-layer streaming kernel), cfg4"; summ gpurun_out/r02_ncu_launches_stack_cfg4.csv
