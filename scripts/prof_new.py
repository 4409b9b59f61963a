"""Runs each of the kernels added / reworked late in round 1 a few times at bench sizes, for an `ncu --set full` capture:
sem_ce_up8 (fused upsample + CE), desc_pos_fwd_planes, bits GEMM with plane-sourced epilogue, combine_heatmap, NMS round."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ssp_b200 as S
from ssp_b200 import synth

dev = torch.device("cuda")
B, Hc, Wc = 32, 30, 40
g = torch.Generator(device=dev); g.manual_seed(0)
rng = np.random.default_rng(0)
Hs = np.stack([np.linalg.inv(synth.sample_homography(rng)) for _ in range(B)]).astype(np.float32)
ex = {"semi": torch.randn((B, 65, Hc, Wc), device=dev, generator=g), "semi_warp": torch.randn((B, 65, Hc, Wc), device=dev, generator=g),
      "desc": torch.nn.functional.normalize(torch.randn((B, 256, Hc, Wc), device=dev, generator=g), dim=1),
      "desc_warp": torch.nn.functional.normalize(torch.randn((B, 256, Hc, Wc), device=dev, generator=g), dim=1),
      "labels_2D": (torch.rand((B, 1, 240, 320), device=dev, generator=g) < 0.005).float(),
      "warped_labels": (torch.rand((B, 1, 240, 320), device=dev, generator=g) < 0.005).float(),
      "mask_2D": torch.ones((B, 1, 240, 320), device=dev), "mask_warp_2D": torch.ones((B, 1, 240, 320), device=dev),
      "mat_H": torch.from_numpy(Hs).to(dev)}
sem = torch.randn((B, 133, Hc, Wc), device=dev, generator=g) * 2
lab = torch.randint(0, 134, (B, 240, 320), device=dev, generator=g)
for _ in range(2):
    leaves = [ex[k].clone().requires_grad_(True) for k in ("semi", "semi_warp", "desc", "desc_warp")]
    out = S.step.loss_step(leaves[0], leaves[1], leaves[2], leaves[3], ex["labels_2D"], ex["warped_labels"], ex["mask_2D"],
                           ex["mask_warp_2D"], ex["mat_H"])
    out["loss"].backward()
    x = sem.clone().requires_grad_(True)
    S.utils.sem_loss(x, lab).backward()
    I, N = 4, 100
    Hw = torch.from_numpy(np.stack([[np.linalg.inv(synth.sample_homography(rng, max_angle=1.57)) for _ in range(N)] for _ in range(I)]).astype(np.float32)).to(dev)
    semi = torch.randn((I, N, 65, Hc, Wc), device=dev, generator=g) * 3
    mask = torch.ones((I, N, 240, 320), device=dev)
    S.step.adaptation_step(semi, Hw, mask)
torch.cuda.synchronize()
print("done")
