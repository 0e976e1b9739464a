"""Phase stamps (ns, %globaltimer) of ONE CTA of a backbone convolution (pillars_set_debug_times hook in conv_umma.cu).
Usage: python profiles/scripts/conv_timeline.py c_in c_out k stride h w"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidar_vision_vqa_b200 import _native, backbone as B
c_in, c_out, k, stride, h, w = [int(a) for a in sys.argv[1:7]]
dev = torch.device("cuda:0")
conv = torch.nn.Conv2d(c_in, c_out, k, stride=stride, padding=0, bias=False)
bn = torch.nn.BatchNorm2d(c_out, eps=1e-3)
l = B._Layer(conv, bn.eval(), False); l.prepare(dev)
nb = 16
x = torch.rand((nb, h, w, c_in), device=dev)
gather = len(sys.argv) > 7 and sys.argv[7] == "gather"
if gather:  # 5 % occupied cells, given as pillar rows + index map
    occ = torch.rand((nb, h, w), device=dev) < 0.05
    idx = torch.full((nb, h, w), -1, dtype=torch.int32, device=dev)
    n_occ = int(occ.sum())
    idx[occ] = torch.arange(n_occ, dtype=torch.int32, device=dev)
    rows_t = torch.rand((n_occ, c_in), device=dev)
oh, ow = B.BaseBEVBackbone._out_hw(h, w, l.desc)
y = torch.empty((nb, oh, ow, c_out), device=dev)
lib = _native.load()
NAMES = {0: "CTA start", 1: "set-up done (TMEM, barriers)", 2: "source table done", 3: "epilogue: first accumulators complete",
         4: "MMA thread: last MMA of tile set 0 issued", 5: "epilogue of tile set 0 done"}
NAMES.update({8 + i: f"loader: halo stage {i} landed" for i in range(8)})
NAMES.update({16 + i: f"MMA thread: halo stage {i} available" for i in range(8)})
rows = []
for rep in range(6):
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    lib.pillars_set_debug_times(buf.data_ptr())
    if gather:
        B.conv_forward(l, y, c_out, 0, False, nb, h, w, rows=rows_t, cell_row=idx)
    else:
        B.conv_forward(l, y, c_out, 0, False, nb, h, w, x_nhwc=x)
    torch.cuda.synchronize()
    lib.pillars_set_debug_times(None)
    v = buf.cpu().numpy()
    if rep >= 2:
        rows.append({k_: int(v[k_]) - int(v[0]) for k_ in NAMES if v[k_] != 0})
print(f"conv {c_in}->{c_out} k{k} s{stride} on 16 x {h} x {w}: microseconds since the CTA's start (CTA 100, median of {len(rows)})")
for k_ in sorted(rows[0], key=lambda q: np.median([r[q] for r in rows if q in r])):
    print(f"  {NAMES[k_]:40s} {np.median([r[k_] for r in rows if k_ in r]) / 1e3:8.2f}")
