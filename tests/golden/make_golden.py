"""Generates tests/golden/*.npz by running the UNMODIFIED reference (Gabriel-SGama/Semantic-SuperPoint) on
seeded inputs.  Run in the authoring container only (it imports /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

Versions at generation time are recorded in tests/golden/VERSIONS.json.  The reference ships no fixtures of
its own for this path, so these files ARE the pin of oracle/ssp_oracle.py.
"""
import collections
import collections.abc
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import cv2  # noqa: E402
import torchvision  # noqa: E402

import ssp_b200  # noqa: E402,F401  (only for the synth generators; no kernels run here)
from ssp_b200 import synth  # noqa: E402

import utils.utils as RU  # noqa: E402  reference
from utils.homographies import sample_homography_np  # noqa: E402


def ref_trainer_class():
    """Import Train_model_heatmap_all with the two missing third-party modules stubbed."""
    collections.Mapping = collections.abc.Mapping
    for name in ("utils.loss_functions.min_norm_solvers", "torch_poly_lr_decay", "tensorboardX"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.MinNormSolver = object
            m.gradient_normalizers = lambda *a, **k: None
            m.PolynomialLRDecay = object
            m.SummaryWriter = object
            sys.modules[name] = m
    import Train_model_heatmap_all as T
    return T.Train_model_heatmap_all


WARP_PARAMS = dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.2,
                   perspective_amplitude_x=0.2, perspective_amplitude_y=0.2, patch_ratio=0.85, max_angle=1.57,
                   allow_artifacts=True)


def ref_homographies(n, seed, identity_first=False):
    np.random.seed(seed)
    Hs = np.stack([sample_homography_np(np.array([2, 2]), shift=-1, **WARP_PARAMS) for _ in range(n)])
    Hs = np.linalg.inv(Hs)  # datasets/Coco.py:345
    if identity_first:
        Hs[0] = np.eye(3)
    return Hs.astype(np.float32), np.linalg.inv(Hs).astype(np.float32)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def gen_semantic():
    """semantic head.  The low-res logits are captured at convSout of the UNMODIFIED reference model
    (models/SuperPointNet_gauss2_ssmall.py:86-91), its own forward upsamples them, and the trainer's sem_loss
    (Train_model_heatmap_all.py:181-193) is the loss; a second case feeds larger synthetic logits through the same
    two reference lines (F.interpolate + sem_loss)."""
    import torch.nn.functional as F
    from models.SuperPointNet_gauss2_ssmall import SuperPointNet_gauss2_ssmall
    t = torch.from_numpy
    Tcls = ref_trainer_class()
    torch.manual_seed(3)
    net = SuperPointNet_gauss2_ssmall()
    net.train()
    grabbed = {}

    def hook(_m, _i, out):
        out.retain_grad()
        grabbed["lr"] = out

    net.convSout.register_forward_hook(hook)
    img = t(synth.uniform((2, 1, 32, 48), 101))
    out = net(img)
    sem = out["sem"]
    sem.retain_grad()
    lab = (synth.uniform((2, 32, 48), 102) * 134).astype(np.int64)      # 0..133, 133 = ignore_index
    lab[0, :5, :] = 133
    loss = Tcls.sem_loss(None, sem, t(lab), "cpu")
    loss.backward()
    # larger logits, odd borders exercised by a 3 x 5 cell map
    lr2 = t(synth.pseudo_normal((1, 133, 3, 5), 103) * 4).requires_grad_(True)
    full2 = F.interpolate(lr2, (24, 40), mode="bilinear", align_corners=False)
    full2.retain_grad()
    lab2 = (synth.uniform((1, 24, 40), 104) * 140).astype(np.int64).clip(0, 133)
    loss2 = Tcls.sem_loss(None, full2, t(lab2), "cpu")
    (loss2 * 0.7).backward()
    save("semantic", lr=grabbed["lr"].detach().numpy(), full_sample=sem.detach().numpy()[:, ::7], label=lab, loss=loss.detach().numpy(),
         dlr=grabbed["lr"].grad.numpy(), dfull_sample=sem.grad.numpy()[:, ::9, ::5, ::7],
         lr2=lr2.detach().numpy(), label2=lab2, loss2=loss2.detach().numpy(), dlr2=lr2.grad.numpy(), g2=np.float32(0.7),
         dfull2_sample=full2.grad.numpy()[:, ::9, ::3, ::5])


def gen_matching():
    """sample_desc_from_points / nn_match_two_way of the unmodified reference (models/model_wrap.py)."""
    import models.model_wrap as MW
    t = torch.from_numpy
    me = types.SimpleNamespace(cell=8, device="cpu")
    coarse = synth.unit_descriptors(1, 256, 15, 20, 111, smooth=0.5)            # 120 x 160 image
    coarse2 = (coarse + 0.35 * synth.unit_descriptors(1, 256, 15, 20, 112, smooth=0.5)).astype(np.float32)
    n = 90
    pts = np.stack([synth.uniform((n,), 113) * 159, synth.uniform((n,), 114) * 119, synth.uniform((n,), 115)]).astype(np.float64)
    pts[:2] = np.round(pts[:2])                                                 # keypoints are integer pixels
    pts[:2, :4] = [[0, 159, 0, 159], [0, 0, 119, 119]]                          # image corners
    pts2 = pts.copy()
    pts2[:2] = np.clip(pts2[:2] + np.round((synth.uniform((2, n), 116) - 0.5) * 3), 0, [[159], [119]])
    pts2 = pts2[:, 17:]                                                         # different counts on the two sides
    d1 = MW.SuperPointFrontend_torch.sample_desc_from_points(me, t(coarse), pts)
    d2 = MW.SuperPointFrontend_torch.sample_desc_from_points(me, t(coarse2), pts2)
    out = {}
    for thr in (0.36, 0.7):
        out["matches_%d" % int(round(thr * 100))] = MW.PointTracker.nn_match_two_way(types.SimpleNamespace(), d1, d2, thr)
    save("matching", pts=pts, pts2=pts2, desc1=d1, desc2=d2, **out)   # coarse maps: synth seeds 111 / 112, see the tests


def gen_warp_labels():
    """datasets/data_tools.warpLabels of the unmodified reference (oracle pinned ahead of the kernels)."""
    from datasets.data_tools import warpLabels
    t = torch.from_numpy
    Hs, _ = ref_homographies(2, 9)
    H, W = 48, 64
    pts = np.stack([np.floor(synth.uniform((120,), 131) * W), np.floor(synth.uniform((120,), 132) * H)], axis=1)
    out = {}
    for i in range(2):
        o = warpLabels(t(pts.copy()), H, W, t(Hs[i]), bilinear=True)
        for k, v in o.items():
            out["%s_%d" % (k, i)] = v.numpy()
    save("warp_labels", pts=pts, H=Hs, **out)


def gen_eval_keypoints():
    """SURVEY 8a-a10, evaluation side: evaluations/detector_evaluation.py warp_keypoints (:139-150) and the repeatability
    masks keep_true_keypoints / filter_keypoints (nested in compute_repeatability, :152-191), pinned through the module
    function and through compute_repeatability's own return values on synthetic detections."""
    import evaluations.detector_evaluation as DE
    Hh, W = 480, 640
    K = 1000
    kp = np.stack([synth.uniform((K,), 141) * (W + 40) - 20, synth.uniform((K,), 142) * (Hh + 40) - 20,
                   synth.uniform((K,), 143)], axis=1).astype(np.float64)
    Hpix = np.array([[0.92, 0.06, 14.0], [-0.05, 1.07, -9.0], [1.1e-4, -1.7e-4, 1.0]])
    warped = DE.warp_keypoints(kp[:, :2], Hpix)
    # detections in the warped image: the true warps of 700 of the points plus noise, and 300 unrelated points
    wk = np.concatenate([warped[:700] + (synth.uniform((700, 2), 144) - 0.5) * 4.0,
                         np.stack([synth.uniform((300,), 145) * W, synth.uniform((300,), 146) * Hh], 1)], axis=0)
    wk = np.concatenate([wk, synth.uniform((K, 1), 147)], axis=1).astype(np.float64)
    data = {"image": np.zeros((Hh, W)), "homography": Hpix, "prob": kp.copy(), "warped_prob": wk.copy()}
    rep, loc = DE.compute_repeatability(data, keep_k_points=300, distance_thresh=3)
    data = {"image": np.zeros((Hh, W)), "homography": Hpix, "prob": kp.copy(), "warped_prob": wk.copy()}
    rep1k, loc1k = DE.compute_repeatability(data, keep_k_points=1000, distance_thresh=3)
    save("eval_keypoints", kp=kp, H=Hpix, warped=warped, warped_prob=wk, shape=np.array([Hh, W]),
         repeatability=np.float64(rep), loc_err=np.float64(loc), repeatability_1000=np.float64(rep1k), loc_err_1000=np.float64(loc1k))


def gen_sparse():
    """utils/loss_functions/sparse_loss.batch_descriptor_loss_sparse of the unmodified reference on CPU.  The
    index lists it samples internally are captured by wrapping the two PixelwiseContrastiveLoss static methods (they are called
    with the lists); losses and gradients are the reference's own autograd results.  Image 1 uses a strongly shrinking
    homography so that fewer than num_matching_attempts cells match (the padding branch of crop_or_pad_choice)."""
    import utils.loss_functions.sparse_loss as SL
    from utils.loss_functions.pixelwise_contrastive_loss import PixelwiseContrastiveLoss as P
    cap = {"m": [], "n": []}
    om, on = P.match_loss, P.non_match_descriptor_loss

    def wm(a, b, ma, mb, **kw):
        cap["m"].append((ma.clone(), mb.clone()))
        return om(a, b, ma, mb, **kw)

    def wn(a, b, na, nb, **kw):
        cap["n"].append((na.clone(), nb.clone()))
        return on(a, b, na, nb, **kw)

    P.match_loss, P.non_match_descriptor_loss = staticmethod(wm), staticmethod(wn)
    B, Hc, Wc, Dch = 3, 30, 40, 256
    Hs, _ = ref_homographies(B, 17)
    Hs[1] = np.array([[2.2, 0.1, 0.05], [-0.1, 2.4, -0.02], [0.0, 0.0, 1.0]], np.float32)  # maps ~1/5 of the cells into the image
    D = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 151, smooth=0.3)).requires_grad_(True)
    Dw = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 152, smooth=0.3)).requires_grad_(True)
    torch.manual_seed(123)
    np.random.seed(321)
    loss, _none, pos, neg = SL.batch_descriptor_loss_sparse(D, Dw, torch.from_numpy(Hs), device="cpu", lamda_d=250,
                                                             num_matching_attempts=1000, num_masked_non_matches_per_match=10)
    g = np.array([1.0, 0.5, 0.25], np.float32)
    (g[0] * loss + g[1] * pos + g[2] * neg).backward()
    save("sparse_loss", H=Hs, seed_torch=np.int64(123), seed_numpy=np.int64(321), g=g,
         matches_a=torch.stack([m[0] for m in cap["m"]]).numpy(), matches_b=torch.stack([m[1] for m in cap["m"]]).numpy(),
         non_a=torch.stack([n[0] for n in cap["n"]]).numpy().astype(np.int32), non_b=torch.stack([n[1] for n in cap["n"]]).numpy().astype(np.int32),
         loss=loss.detach().numpy(), pos=pos.detach().numpy(), neg=neg.detach().numpy(),
         dD_sample=D.grad.numpy()[:, :, ::3, ::4].copy(), dDw_sample=Dw.grad.numpy()[:, :, ::3, ::4].copy(),
         dD_abs_sum=np.float64(np.abs(D.grad.numpy()).sum()), dDw_abs_sum=np.float64(np.abs(Dw.grad.numpy()).sum()))
    P.match_loss, P.non_match_descriptor_loss = staticmethod(om), staticmethod(on)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sparse":
        return gen_sparse()
    if len(sys.argv) > 1 and sys.argv[1] == "eval_keypoints":
        return gen_eval_keypoints()
    if len(sys.argv) > 1 and sys.argv[1] == "warp_labels":
        return gen_warp_labels()
    if len(sys.argv) > 1 and sys.argv[1] == "semantic":
        return gen_semantic()
    if len(sys.argv) > 1 and sys.argv[1] == "matching":
        return gen_matching()
    torch.manual_seed(0)
    t = torch.from_numpy

    # ---- a1 / a10 warp_points, filter_points
    H3, _ = ref_homographies(3, 1)
    pts = (synth.uniform((57, 2), 11) * 2 - 1).astype(np.float32)
    wp_b = RU.warp_points(t(pts), t(H3)).numpy()
    wp_1 = RU.warp_points(t(pts), t(H3[1])).numpy()
    pix = (synth.uniform((200, 2), 12) * np.array([80, 60]) - 8).astype(np.float32)
    fp, fm = RU.filter_points(t(pix), torch.tensor([64, 48]), return_mask=True)
    save("warp_points", H=H3, pts=pts, out_batched=wp_b, out_single=wp_1, pix=pix, filt_pts=fp.numpy(), filt_mask=fm.numpy())

    # ---- a2 inv_warp_image_batch
    Hs, Hinv = ref_homographies(3, 2)
    img = synth.uniform((3, 1, 48, 64), 21)
    ob = RU.inv_warp_image_batch(t(img), t(Hinv), mode="bilinear").numpy()
    on = RU.inv_warp_image_batch(t(img), t(Hinv), mode="nearest").numpy()
    o1 = RU.inv_warp_image(t(img[0, 0]), t(Hinv[0]), mode="bilinear").numpy()
    oi = RU.inv_warp_image_batch(t(img[:1]), torch.eye(3)).numpy()
    save("inv_warp", img=img, Hinv=Hinv, out_bilinear=ob, out_nearest=on, out_single=o1, out_identity=oi)

    # ---- a3 compute_valid_mask (+ structuring elements of the container's OpenCV)
    _, Hinv5 = ref_homographies(5, 3)
    vm = {}
    for r in (0, 1, 3):
        vm["mask_r%d" % r] = RU.compute_valid_mask(torch.tensor([48, 64]), t(Hinv5), erosion_radius=r).numpy().astype(np.uint8)
    _, Hinv2 = ref_homographies(2, 4)
    vm["mask_240_r3"] = RU.compute_valid_mask(torch.tensor([240, 320]), t(Hinv2), erosion_radius=3).numpy().astype(np.uint8)
    vm["mask_identity_r3"] = RU.compute_valid_mask(torch.tensor([48, 64]), torch.eye(3), erosion_radius=3).numpy().astype(np.uint8)
    for r in range(1, 9):
        vm["ellipse_%d" % r] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * r, 2 * r))
    save("valid_mask", Hinv5=Hinv5, Hinv2=Hinv2, **vm)

    # ---- a4 labels2Dto3D, getMasks, detector_loss (+ gradient)
    lab_bin = synth.keypoint_labels(2, 48, 64, 31, p=0.02)
    lab_soft = np.clip(cv2.GaussianBlur(lab_bin[0, 0], (0, 0), 0.2), 0, 1)[None, None]  # sigma 0.2 leak across cells
    lab_soft = np.concatenate([lab_soft, synth.uniform((1, 1, 48, 64), 32) * (synth.uniform((1, 1, 48, 64), 33) < 0.03)], 0)
    lab_soft = lab_soft.astype(np.float32)
    tiny = lab_bin.copy() * np.float32(1e-9)  # sub-rounding cell sums: dustbin stays 1
    l3_bin = RU.labels2Dto3D(t(lab_bin), 8, add_dustbin=True).numpy()
    l3_soft = RU.labels2Dto3D(t(lab_soft), 8, add_dustbin=True).numpy()
    l3_tiny = RU.labels2Dto3D(t(tiny), 8, add_dustbin=True).numpy()
    l3_nodust = RU.labels2Dto3D(t(lab_bin), 8, add_dustbin=False).numpy()
    mask2d = RU.compute_valid_mask(torch.tensor([48, 64]), t(Hinv5[:2]), erosion_radius=3).unsqueeze(1)
    Tcls = ref_trainer_class()
    m3 = Tcls.getMasks(None, mask2d, 8).numpy()
    semi = torch.from_numpy(synth.pseudo_normal((2, 65, 6, 8), 34) * 2).requires_grad_(True)
    loss = Tcls.detector_loss(None, semi, t(l3_bin), t(m3), loss_type="softmax")
    loss.backward()
    semi2 = torch.from_numpy(synth.pseudo_normal((2, 65, 6, 8), 35) * 3).requires_grad_(True)
    loss2 = Tcls.detector_loss(None, semi2, t(l3_soft), t(m3), loss_type="softmax")
    (loss2 * 2.5).backward()
    save("detector", lab_bin=lab_bin, lab_soft=lab_soft, lab_tiny=tiny, l3_bin=l3_bin, l3_soft=l3_soft, l3_tiny=l3_tiny,
         l3_nodust=l3_nodust, mask2d=mask2d.numpy(), mask3d=m3, semi=semi.detach().numpy(), loss=loss.detach().numpy(),
         dsemi=semi.grad.numpy(), semi2=semi2.detach().numpy(), loss2=loss2.detach().numpy(), dsemi2=semi2.grad.numpy(),
         g2=np.float32(2.5))

    # ---- a6 flattenDetection
    semif = synth.pseudo_normal((3, 65, 6, 8), 41) * 2
    save("flatten", semi=semif, heat=RU.flattenDetection(t(semif)).numpy(), heat3d=RU.flattenDetection(t(semif[0])).numpy())

    # ---- a7 combine_heatmap (restated 6 lines of export.py:49-60 over the reference's inv_warp_image_batch)
    N = 7
    Hs7, Hinv7 = ref_homographies(N, 5, identity_first=True)
    heat = synth.uniform((N, 1, 48, 64), 51) * 0.3
    m2d = RU.compute_valid_mask(torch.tensor([48, 64]), t(Hinv7), erosion_radius=0).unsqueeze(1)
    hh = t(heat) * m2d
    hh = RU.inv_warp_image_batch(hh, t(Hs7), mode="bilinear")   # export.py:284-287 swaps the names on purpose
    mm = RU.inv_warp_image_batch(m2d, t(Hs7), mode="bilinear")
    comb = (torch.sum(hh, dim=0) / torch.sum(mm, dim=0)).numpy()
    save("combine", heat=heat, mask=m2d.numpy(), Hwarp=Hs7, out=comb)

    # ---- a8 getPtsFromHeatmap / nms_fast on tie-free heatmaps (inputs regenerated from the seed)
    pts_a = RU.getPtsFromHeatmap(synth.unique_heatmap(120, 160, 61), 0.015, 4)
    pts_b = RU.getPtsFromHeatmap(synth.unique_heatmap(240, 320, 62), 0.015, 4)
    pts_c = RU.getPtsFromHeatmap(synth.unique_heatmap(64, 96, 63), 0.03, 2)
    sparse = synth.unique_heatmap(48, 64, 64) * (synth.uniform((48, 64), 65) < 0.05)
    pts_d = RU.getPtsFromHeatmap(sparse.astype(np.float32), 0.015, 4)
    one = np.zeros((48, 64), np.float32); one[20, 30] = 0.5
    pts_e = RU.getPtsFromHeatmap(one, 0.015, 4)
    corners = np.stack([synth.uniform((300,), 66) * 63, synth.uniform((300,), 67) * 47, synth.unique_heatmap(1, 300, 68)[0]])
    corners[:2] = np.round(corners[:2])
    _, uniq = np.unique(corners[1] * 64 + corners[0], return_index=True)
    corners = corners[:, np.sort(uniq)].astype(np.float64)
    nf_out, nf_inds = RU.nms_fast(corners, 48, 64, 4)
    save("nms", pts_120=pts_a, pts_240=pts_b, pts_64=pts_c, pts_sparse=pts_d, pts_one=pts_e, sparse=sparse.astype(np.float32),
         corners=corners, nms_fast_out=nf_out, nms_fast_inds=nf_inds)

    # ---- a9 box_nms: reference lines 629-649 need CUDA; replayed with torchvision.ops.nms on CPU
    prob = synth.unique_heatmap(48, 64, 71, hi=1.0) * (synth.uniform((48, 64), 72) < 0.3)
    prob = prob.astype(np.float32)
    p_t = t(prob)
    pts_t = torch.nonzero(p_t > 0.01).float()
    sz = torch.tensor(4 / 2.0)
    boxes = torch.cat([pts_t - sz, pts_t + sz], dim=1)
    scores = p_t[pts_t[:, 0].long(), pts_t[:, 1].long()]
    idx = torchvision.ops.nms(boxes, scores, 0.1)
    out = torch.zeros_like(p_t)
    out[pts_t[idx, 0].long(), pts_t[idx, 1].long()] = scores[idx]
    save("box_nms", prob=prob, out=out.numpy())

    # ---- a5 descriptor_loss forward + backward, small (inputs stored) and 30x40 (inputs from seed)
    def run_desc(D, Dw, Hm, mv, g):
        Dt, Dwt = t(D).requires_grad_(True), t(Dw).requires_grad_(True)
        loss, mask, pos, neg = RU.descriptor_loss(Dt, Dwt, t(Hm), mask_valid=t(mv), lamda_d=250, descriptor_dist=4, lambda_d=800)
        (g[0] * loss + g[1] * pos + g[2] * neg).backward()
        return dict(loss=loss.detach().numpy(), pos=pos.detach().numpy(), neg=neg.detach().numpy(),
                    mask=mask.numpy().astype(np.uint8), dD=Dt.grad.numpy(), dDw=Dwt.grad.numpy())

    Hd, _ = ref_homographies(2, 6)
    D = synth.unit_descriptors(2, 256, 10, 12, 81, smooth=0.35)
    Dw = synth.unit_descriptors(2, 256, 10, 12, 82, smooth=0.35)
    mv = (synth.uniform((2, 1, 10, 12), 83) < 0.8).astype(np.float32)
    r = run_desc(D, Dw, Hd, mv, (1.0, 0.0, 0.0))
    r2 = run_desc(D, Dw, Hd, mv, (0.3, 1.7, 0.9))
    save("desc_small", H=Hd, D=D, Dw=Dw, mv=mv, g_b=np.array([0.3, 1.7, 0.9], np.float32),
         **{k + "_a": v for k, v in r.items()}, **{k + "_b": v for k, v in r2.items()})

    Hd1, _ = ref_homographies(1, 7)
    D = synth.unit_descriptors(1, 256, 30, 40, 91, smooth=0.3)
    Dw = synth.unit_descriptors(1, 256, 30, 40, 92, smooth=0.3)
    mv = (synth.uniform((1, 1, 30, 40), 93) < 0.85).astype(np.float32)
    r = run_desc(D, Dw, Hd1, mv, (1.0, 1.0, 1.0))
    save("desc_30x40", H=Hd1, mv=mv, loss=r["loss"], pos=r["pos"], neg=r["neg"], mask_rowsum=r["mask"].reshape(1, 1200, 1200).sum(-1),
         dD_sample=r["dD"][0, :, ::7, ::9], dDw_sample=r["dDw"][0, :, ::7, ::9])
    # identity homography, identical descriptors: positive term must vanish (the reference's informal KAT,
    # utils/loss_functions/sparse_loss.py:345 "pos should be 0")
    rI = run_desc(D, D.copy(), np.eye(3, dtype=np.float32)[None], np.ones((1, 1, 30, 40), np.float32), (1.0, 0.0, 0.0))
    save("desc_identity", loss=rI["loss"], pos=rI["pos"], neg=rI["neg"])

    gen_semantic()
    gen_matching()
    gen_warp_labels()

    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump({"torch": torch.__version__, "numpy": np.__version__, "cv2": cv2.__version__,
                   "torchvision": torchvision.__version__, "reference": "Gabriel-SGama/Semantic-SuperPoint @ /root/reference"}, f, indent=1)


if __name__ == "__main__":
    main()
