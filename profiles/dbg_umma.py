import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from lidar_vision_vqa_b200 import tokens as T, ops
from oracle import tokens_oracle as to
dev = torch.device("cuda:0")
c, d, h, w, b = 32, 128, 16, 16, 2
sd = to.random_token_params(c, d, seed=1)
rng = np.random.default_rng(0)
occ = rng.random((b, 1, h, w)) < 0.3
bev = np.where(occ, np.maximum(rng.standard_normal((b, c, h, w)), 0), 0).astype(np.float32)
ref = to.bev_tokens(bev, sd)
for proj in ("fma", "umma"):
    tk = T.VATLiDARTokenizer(c, d, projection=proj)
    tk.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tk = tk.eval().to(dev)
    out = tk(torch.from_numpy(bev).to(dev))
    torch.cuda.synchronize()
    ws = ops.workspace(0, dev, slot=7)
    err = np.abs(out.cpu().numpy() - ref)
    print(proj, "max err", err.max(), "bad cells", int((err.max(-1) > 1e-3).sum()), "of", b * h * w, flush=True)
    if proj == "umma":
        import ctypes
        from lidar_vision_vqa_b200 import _native
        lib = _native.load()
        mapb = (4 * b * h * w + 255) // 256 * 256
        words = ws[mapb:mapb + 16].view(torch.int32).cpu().numpy()
        print("count/debug words", [hex(int(x) & 0xffffffff) for x in words], "active pairs expected", int((err.max(-1) >= 0).sum()))
        bad = np.argwhere(err.max(-1) > 1e-3)
        print("first bad", bad[:10].tolist())
        if len(bad):
            bb, cc = bad[0]
            print("got", out[bb, cc, :8].cpu().numpy(), "\nref", ref[bb, cc, :8])
