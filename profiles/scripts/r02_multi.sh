# usage: bash profiles/scripts/r02_multi.sh N   (under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-tokens > gpurun_out/r02_bench_cfg2_n$N.json 2> gpurun_out/err_cfg2_n$N.txt; tail -c 600 gpurun_out/err_cfg2_n$N.txt
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-tokens --no-e2e --gather-dtype float16 > gpurun_out/r02_bench_cfg2_n${N}_fp16wire.json 2> gpurun_out/err_cfg2f_n$N.txt; tail -c 300 gpurun_out/err_cfg2f_n$N.txt
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-tokens --no-e2e --workload cfg4_waymo64_pillar0.1_bev1024 --streams 2 > gpurun_out/r02_bench_cfg4_n$N.json 2> gpurun_out/err_cfg4_n$N.txt; tail -c 600 gpurun_out/err_cfg4_n$N.txt
if [ "$N" -ge 8 ]; then
$TR bench.py --gpus $N --steps 4 --workload cfg5_e2e_b256_fusion > gpurun_out/r02_bench_cfg5_n$N.json 2> gpurun_out/err_cfg5_n$N.txt; tail -c 600 gpurun_out/err_cfg5_n$N.txt
$TR bench.py --gpus $N --steps 4 --workload cfg5_e2e_b256_fusion --cfg5-mode sharded > gpurun_out/r02_bench_cfg5_sharded_n$N.json 2> gpurun_out/err_cfg5s_n$N.txt; tail -c 600 gpurun_out/err_cfg5s_n$N.txt
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_n$N*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "NO LINE", e); continue
    g=d.get("gather_to_fusion_rank") or {}
    e=d.get("e2e") or {}
    print(f.split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "with_gather", round(g.get("value_with_gather",0)), "check", g.get("gather_check"), "xfer_ms", g.get("transfer_only_ms_per_step"), "e2e", round(e.get("value",0)), "ceil", (e.get("h2d_copy_ceiling") or {}).get("sweeps_per_s"))
PY
