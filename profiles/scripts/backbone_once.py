"""One forward of BaseBEVBackbone on a 16 x 512^2 pillar canvas given as rows + index map (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidar_vision_vqa_b200 import backbone as B
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
occ = torch.rand((nb, 512, 512), generator=g) < 0.05
pos = occ.nonzero()
idx = torch.full((nb, 512, 512), -1, dtype=torch.int32)
idx[pos[:, 0], pos[:, 1], pos[:, 2]] = torch.arange(len(pos), dtype=torch.int32)
rows = torch.rand((len(pos), 64), generator=g).to(dev)
m = B.BaseBEVBackbone(dict(LAYER_NUMS=[3, 5, 5], LAYER_STRIDES=[2, 2, 2], NUM_FILTERS=[64, 128, 256],
                           UPSAMPLE_STRIDES=[0.5, 1, 2], NUM_UPSAMPLE_FILTERS=[128, 128, 128]), 64).eval().to(dev)
with torch.inference_mode():
    out = m({"pillar_features": rows, "bev_index_map": idx.to(dev)})
torch.cuda.synchronize()
print(out["spatial_features_2d"].shape, int(out["_conv_error_word"].item()))
