"""Per-row kernel table: device time (CUDA events), algorithmic bytes / flops, fraction of the measured
roofline, and the CPU oracle timed beside it on a bounded sample.  Writes gpurun_out/kernels.json + kernels.md.

    python scripts/bench_kernels.py            # on a B200
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import ssp_b200 as S
from oracle import ssp_oracle as O
from ssp_b200 import synth

dev = "cuda"
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
rows = []


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def gpu_time(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def cpu_time(fn, n=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def add(row, what, us, bytes_=None, flops=None, cpu_us=None, cpu_scale=1.0, note=""):
    r = {"row": row, "what": what, "gpu_us": round(us, 1)}
    if bytes_:
        r["GBps"] = round(bytes_ / us / 1e3, 1)
        r["frac_hbm"] = round(bytes_ / us / 1e3 / PK["hbm_gbs"], 3)
    if flops:
        r["TFLOPs"] = round(flops / us / 1e6, 1)
        r["frac_bf16"] = round(flops / us / 1e6 / PK["bf16_tflops"], 3)
    if cpu_us is not None:
        r["cpu_us_same_work"] = round(cpu_us * cpu_scale, 0)
        r["speedup_vs_cpu_oracle"] = round(cpu_us * cpu_scale / us, 0)
    r["note"] = note
    rows.append(r)
    print(r, flush=True)


def homs(n, seed, identity_first=False):
    rng = np.random.default_rng(seed)
    Hs = np.stack([np.linalg.inv(synth.sample_homography(rng)) for _ in range(n)])
    if identity_first:
        Hs[0] = np.eye(3)
    return Hs.astype(np.float32), np.linalg.inv(Hs).astype(np.float32)


H, W, Hc, Wc = 240, 320, 30, 40
# a1
Hs, Hinv = homs(32, 1)
pts = (synth.uniform((76800, 2), 3) * 2 - 1).astype(np.float32)
p_d, H_d = cu(pts), cu(Hs)
us = gpu_time(lambda: S.warp_points(p_d, H_d, device=dev))
add("a1", "warp_points 76800 pts x 32 H", us, 76800 * 8 + 32 * 76800 * 8, cpu_us=cpu_time(lambda: O.warp_points(pts, Hs[:4])), cpu_scale=8)
# a2 / a3
Hs100, Hinv100 = homs(100, 2, True)
img = cu(synth.uniform((100, 1, H, W), 5))
Hinv_d = cu(Hinv100)
img4 = synth.uniform((4, 1, H, W), 5)
for mode in ("bilinear", "nearest"):
    us = gpu_time(lambda: S.inv_warp_image_batch(img, Hinv_d, device=dev, mode=mode))
    add("a2", "inv_warp_image_batch %s 100x240x320" % mode, us, 100 * 2 * H * W * 4,
        cpu_us=cpu_time(lambda: O.inv_warp_image_batch(img4, Hinv100[:4], mode)), cpu_scale=25)
for r in (0, 3):
    us = gpu_time(lambda: S.compute_valid_mask(torch.tensor([H, W]), Hinv_d, device=dev, erosion_radius=r))
    add("a3", "compute_valid_mask r=%d 100x240x320" % r, us, 100 * H * W * 4,
        cpu_us=cpu_time(lambda: O.compute_valid_mask((H, W), Hinv100[:4], r)), cpu_scale=25)
# a4
B = 32
lab = synth.keypoint_labels(B, H, W, 7)
lab_d = cu(lab)
mask_d = S.compute_valid_mask(torch.tensor([H, W]), cu(Hinv[:B]), device=dev, erosion_radius=3).unsqueeze(1)
us = gpu_time(lambda: S.labels2Dto3D(lab_d, 8))
add("a4", "labels2Dto3D B=32", us, B * (H * W * 4 + 65 * 1200 * 4), cpu_us=cpu_time(lambda: O.labels2Dto3D(lab[:4])), cpu_scale=8)
us = gpu_time(lambda: S.getMasks(mask_d, 8, device=dev))
add("a4", "getMasks B=32", us, B * (H * W * 4 + 1200 * 4))
semi = synth.pseudo_normal((B, 65, Hc, Wc), 9)
semi_d = cu(semi).requires_grad_(True)
semi2_d = cu(semi).requires_grad_(True)
mask_np = mask_d.cpu().numpy()


def det_fwd():
    return S.detector_loss_pair_2d(semi_d, lab_d, mask_d, semi2_d, lab_d, mask_d)


us_f = gpu_time(det_fwd)
l3, m3 = O.labels2Dto3D(lab[:4]), O.getMasks(mask_np[:4])
add("a4", "detector loss x2 fwd (labels2Dto3D+getMasks fused) B=32", us_f, 2 * B * (65 * 1200 * 4 + 2 * H * W * 4),
    cpu_us=cpu_time(lambda: O.detector_loss(semi[:4], O.labels2Dto3D(lab[:4]), O.getMasks(mask_np[:4]))), cpu_scale=16)


def det_fb():
    a, b, _ = det_fwd()
    (a + b).backward()


us_fb = gpu_time(det_fb)
add("a4", "detector loss x2 fwd+bwd B=32", us_fb, 2 * B * (3 * 65 * 1200 * 4 + 4 * H * W * 4),
    cpu_us=cpu_time(lambda: O.detector_loss(semi[:4], l3, m3, grad=True)), cpu_scale=16)
# a6 / a7
N = 100
semis = cu(synth.pseudo_normal((N, 65, Hc, Wc), 11) * 3)
us = gpu_time(lambda: S.flattenDetection(semis))
semis_np = semis.cpu().numpy()
add("a6", "flattenDetection N=100", us, N * (65 * 1200 * 4 + H * W * 4), cpu_us=cpu_time(lambda: O.flattenDetection(semis_np[:10])), cpu_scale=10)
heat = S.flattenDetection(semis)
masks = S.compute_valid_mask(torch.tensor([H, W]), Hinv_d, device=dev).unsqueeze(1)
Hw = cu(Hs100[None])
us = gpu_time(lambda: S.combine_heatmap(heat, Hw, masks, device=dev))
heat_np, masks_np = heat.cpu().numpy(), masks.cpu().numpy()
add("a7", "combine_heatmap N=100 240x320", us, (2 * N + 1) * H * W * 4,
    cpu_us=cpu_time(lambda: O.combine_heatmap(heat_np[:10], Hs100[None, :10], masks_np[:10])), cpu_scale=10)
# a8 / a9
agg = S.combine_heatmap(heat, Hw, masks, device=dev)[0]
agg_np = agg.cpu().numpy()
us = gpu_time(lambda: S.getPtsFromHeatmap(agg, 0.015, 4), n=10)
add("a8", "getPtsFromHeatmap 240x320 aggregated heatmap (incl. D2H + sync)", us, H * W * 4,
    cpu_us=cpu_time(lambda: O.getPtsFromHeatmap(agg_np, 0.015, 4), n=1), note="latency-bound rounds")
big = cu(synth.unique_heatmap(480, 640, 77))
big_np = big.cpu().numpy()
us = gpu_time(lambda: S.getPtsFromHeatmap(big, 0.015, 4), n=10)
add("a8", "getPtsFromHeatmap 480x640 dense unique heatmap", us, 480 * 640 * 4, cpu_us=cpu_time(lambda: O.getPtsFromHeatmap(big_np, 0.015, 4), n=1))
prob = cu((synth.unique_heatmap(H, W, 5, hi=1.0) * (synth.uniform((H, W), 6) < 0.2)).astype(np.float32))
prob_np = prob.cpu().numpy()
us = gpu_time(lambda: S.box_nms(prob, 4, keep_top_k=1), n=10)
add("a9", "box_nms 240x320, 20% candidates", us, 2 * H * W * 4, cpu_us=cpu_time(lambda: O.box_nms(prob_np, 4), n=1))
# a10
kp = np.stack([synth.uniform((1000,), 1) * 640, synth.uniform((1000,), 2) * 480], 1).astype(np.float64)
Hpix = np.array([[0.9, 0.05, 12.0], [-0.04, 1.1, -7.0], [1e-4, -2e-4, 1.0]])
us = gpu_time(lambda: S.warp_keypoints(kp, Hpix, shape=(480, 640)), n=10)
add("a10", "warp_keypoints f64 + in-bounds, 1000 pts (incl. H2D/D2H)", us, cpu_us=cpu_time(lambda: O.keep_in_bounds_f64(O.warp_keypoints_f64(kp, Hpix), (480, 640))))


# a5
def desc_case(tag, B, Hc, Wc, engines, cpu_pairs):
    Nc = Hc * Wc
    D = synth.unit_descriptors(B, 256, Hc, Wc, 21, smooth=0.3)
    Dw = synth.unit_descriptors(B, 256, Hc, Wc, 22, smooth=0.3)
    Hm, _ = homs(B, 23)
    mv = (synth.uniform((B, 1, Hc, Wc), 24) < 0.9).astype(np.float32)
    Dd, Dwd, Hd, mvd = cu(D), cu(Dw), cu(Hm), cu(mv)
    flops = 2.0 * Nc * Nc * 256 * B
    cpu_f = cpu_time(lambda: O.descriptor_loss(D[:cpu_pairs], Dw[:cpu_pairs], Hm[:cpu_pairs], mv[:cpu_pairs]), n=1)
    cpu_fb = cpu_time(lambda: O.descriptor_loss(D[:cpu_pairs], Dw[:cpu_pairs], Hm[:cpu_pairs], mv[:cpu_pairs], grad=(1, 0, 0)), n=1)
    for e in engines:
        us = gpu_time(lambda: S.descriptor_loss(Dd, Dwd, Hd, mask_valid=mvd, device=dev, engine=e), n=10)
        add("a5", "descriptor_loss fwd %s B=%d %dx%d cells [%s]" % (tag, B, Hc, Wc, e), us, flops=flops, cpu_us=cpu_f, cpu_scale=B / cpu_pairs)
        Dg, Dwg = cu(D).requires_grad_(True), cu(Dw).requires_grad_(True)

        def fb():
            l, _, p, n = S.descriptor_loss(Dg, Dwg, Hd, mask_valid=mvd, device=dev, engine=e)
            l.backward()

        us = gpu_time(fb, n=10)
        add("a5", "descriptor_loss fwd+bwd %s B=%d %dx%d cells [%s]" % (tag, B, Hc, Wc, e), us, flops=3 * flops, cpu_us=cpu_fb, cpu_scale=B / cpu_pairs,
            note="3 GEMMs of 2*Nc^2*256 credited (fwd + 2 bwd)")


desc_case("240x320", 32, 30, 40, ("bf16x3", "bf16", "fp32"), 2)
desc_case("480x640", 4, 60, 80, ("bf16x3", "bf16"), 1)
desc_case("376x1240", 4, 47, 155, ("bf16x3", "bf16"), 1)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"peaks": PK, "cpu_cores": os.cpu_count(), "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "kernels.md"), "w") as f:
    f.write("| row | case | GPU us | GB/s | frac HBM | TFLOP/s | frac bf16 | CPU oracle us (same work, %d cores) | x |\n|---|---|---|---|---|---|---|---|---|\n" % os.cpu_count())
    for r in rows:
        f.write("| %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (r["row"], r["what"], r["gpu_us"], r.get("GBps", ""), r.get("frac_hbm", ""), r.get("TFLOPs", ""),
                                                                   r.get("frac_bf16", ""), r.get("cpu_us_same_work", ""), r.get("speedup_vs_cpu_oracle", "")))
print("done")
