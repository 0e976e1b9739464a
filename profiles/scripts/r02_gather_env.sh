run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-tokens --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
g=d['gather_to_fusion_rank']; print(round(d['value']), round(g['value_with_gather']), 'xfer ms', round(g['transfer_only_ms_per_step'],4), 'check', g['gather_check'])"; }
run X=1
run NCCL_P2P_USE_CUDA_MEMCPY=1
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=2
run NCCL_MIN_NCHANNELS=32
