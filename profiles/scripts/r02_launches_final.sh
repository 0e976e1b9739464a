# ncu launch list of the bench command (encoder step only; same command runs first without ncu and must exit 0)
CMD="python bench.py --steps 3 --warmup 3 --repeats 1 --no-cpu --no-extra-workloads --no-extractor --no-e2e --no-tokens --no-backbone"
$CMD > gpurun_out/r02_plain_short.json 2> gpurun_out/r02_plain_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_final.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_ncu_launches_final.csv")) if len(r) > 5 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].split("::")[-1]
    tot[name] += float(r[-1]); cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(12):
    print(f"{k:40s} launches {cnt[k]:4d}  mean {v / cnt[k] / 1e3:8.2f} us  share {100 * v / s:5.1f} %")
PY
