"""Timing of the BEV tokeniser (csrc/tokens.cu) on the cfg2 canvas: pillars_encode_bev -> tokens from the encoder's index
map.  Usage: python profiles/time_tokens.py [frames] [d_model] [--dense] [--fma|--mma]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lidar_vision_vqa_b200 as L  # noqa: E402
from lidar_vision_vqa_b200 import ops, synth  # noqa: E402
from lidar_vision_vqa_b200 import tokens as T  # noqa: E402
from oracle import pillar_oracle as po  # noqa: E402  (random weights only)
from oracle import tokens_oracle as to  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
model, gc, _ = synth.WORKLOADS["cfg2_nuscenes32_b16_pillar0.2_bev512"]
grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, 32, 30000)
pts, offs = synth.make_batch(nb, model, 5)
sdp = po.random_pfn_params(11, [64], True, seed=0)
pfn = ops.fold_pfn(sdp["pfn_layers.0.linear.weight"], (sdp["pfn_layers.0.norm.weight"], sdp["pfn_layers.0.norm.bias"],
                   sdp["pfn_layers.0.norm.running_mean"], sdp["pfn_layers.0.norm.running_var"], 1e-3), None, c_point=5,
                   use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                   point_cloud_range=grid.point_cloud_range, device=dev)
p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
bufs = ops.EncodeBuffers(len(pts), nb, grid, 64, dev)
res = ops.encode_bev(p, o, grid, pfn, buffers=bufs)
sd = to.random_token_params(64, d, seed=11)
proj = "mma" if "--mma" in sys.argv else "fma" if "--fma" in sys.argv else "umma"
tk = T.VATLiDARTokenizer(64, d, projection=proj)
tk.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
tk = tk.eval().to(dev)
cell_row = T.encode_index_map(bufs, len(pts), nb, grid)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); tk.tables(512, 512); e.record(); torch.cuda.synchronize()
t_prep = s.elapsed_time(e)
out = torch.empty((nb, 512 * 512, d), dtype=torch.float32, device=dev)
occ = (cell_row >= 0).float()[:, None]
active = float((torch.nn.functional.max_pool2d(occ, 3, stride=1, padding=1) > 0).float().mean())
dense = "--dense" in sys.argv
fn = (lambda: tk(res["bev"], out=out)) if dense else (lambda: tk.forward_index_map(res["pillar_features"], cell_row, out=out))
for _ in range(3):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = float(np.median(ts))
gb = out.numel() * 4 / 1e9
pe_gb = 512 * 512 * d * 4 / 1e9
print(f"tokens nb={nb} d={d} {'dense' if dense else 'rows'} {proj}: {ms * 1e3:.0f} us, write {gb / ms * 1e3:.0f} GB/s "
      f"(+PE read {pe_gb:.2f} GB), cells with a non-empty window {100 * active:.1f} %, tables prepared in {t_prep:.1f} ms",
      flush=True)
