"""Host-side logic that needs no GPU: structuring elements, NMS stencils, input generators, drop-in binding,
loud failure without CUDA."""
import hashlib
import types

import numpy as np
import pytest
import torch

import ssp_b200 as S
from ssp_b200 import synth


def test_ellipse_kernel_matches_opencv(golden):
    g = golden("valid_mask")
    for r in range(1, 9):
        assert np.array_equal(S.ellipse_kernel(r), g["ellipse_%d" % r])
    assert S.ellipse_kernel(3).tolist() == [[0, 0, 0, 1, 0, 0], [0, 1, 1, 1, 1, 1], [1] * 6, [1] * 6, [1] * 6, [0, 1, 1, 1, 1, 1]]


def test_box_nms_stencil_size4():
    """with size=4, IoU>0.1 suppresses exactly these |dx|,|dy| offsets."""
    from ssp_b200.utils import _iou_stencil
    st, R = _iou_stencil(4, 0.1, "cpu")
    assert R == 3
    st = st.numpy().reshape(7, 7)
    want = {(0, 1), (0, 2), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (0, 0)}
    got = {(abs(dx), abs(dy)) for dy in range(-3, 4) for dx in range(-3, 4) if st[dy + 3, dx + 3]}
    assert got == want


def test_synth_is_bit_reproducible():
    h = hashlib.sha256(synth.uniform((1000,), 7).tobytes()).hexdigest()
    assert h == hashlib.sha256(synth.uniform((1000,), 7).tobytes()).hexdigest()
    u = synth.uniform((4096,), 1)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    assert float(u[0]) == pytest.approx(float(synth.uniform((1,), 1)[0]))
    d = synth.unit_descriptors(2, 256, 3, 4, 5)
    np.testing.assert_allclose((d.astype(np.float64) ** 2).sum(1), 1.0, atol=1e-6)
    hm = synth.unique_heatmap(24, 32, 3)
    assert len(np.unique(hm)) == 24 * 32
    Hm = synth.sample_homography(np.random.default_rng(0))
    assert Hm.shape == (3, 3) and abs(Hm[2, 2] - 1.0) < 0.2 and abs(np.linalg.det(Hm)) > 0.1


def test_dropin_binds_and_restores():
    fake = types.ModuleType("fake_utils")
    fake.descriptor_loss = lambda *a, **k: "ref"
    fake.warp_points = lambda *a, **k: "ref"

    class Trainer:
        def detector_loss(self, *a, **k):
            return "ref"

    bound = S.dropin.install(utils_module=fake, trainer_class=Trainer)
    assert fake.descriptor_loss is S.utils.descriptor_loss and fake.labels2Dto3D is S.utils.labels2Dto3D
    assert any(b.endswith("detector_loss") for b in bound)
    # dataset-side names: ours in the main process, the reference's own function inside a DataLoader worker (no CUDA there)
    assert fake.warp_points.__wrapped__ is S.utils.warp_points
    import torch.utils.data as tud
    orig_info = tud.get_worker_info
    tud.get_worker_info = lambda: object()
    try:
        assert fake.warp_points(torch.zeros(4, 2), torch.eye(3)) == "ref"
    finally:
        tud.get_worker_info = orig_info
    sparse_mod = types.ModuleType("fake_sparse")
    sparse_mod.batch_descriptor_loss_sparse = lambda *a, **k: "ref"
    bound = S.dropin.install(utils_module=fake, sparse_module=sparse_mod)
    assert sparse_mod.batch_descriptor_loss_sparse is S.sparse.batch_descriptor_loss_sparse
    S.dropin.uninstall()
    assert fake.descriptor_loss() == "ref" and Trainer().detector_loss() == "ref" and fake.warp_points() == "ref"
    assert sparse_mod.batch_descriptor_loss_sparse() == "ref"


def test_signatures_mirror_reference():
    import inspect
    sig = inspect.signature(S.descriptor_loss)
    assert list(sig.parameters)[:8] == ["descriptors", "descriptors_warped", "homographies", "mask_valid", "cell_size",
                                        "lamda_d", "device", "descriptor_dist"]
    assert sig.parameters["lamda_d"].default == 250 and sig.parameters["descriptor_dist"].default == 4
    assert any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())  # **config swallows lambda_d=800
    assert list(inspect.signature(S.inv_warp_image_batch).parameters)[:4] == ["img", "mat_homo_inv", "device", "mode"]
    assert list(inspect.signature(S.warp_points).parameters) == ["points", "homographies", "device"]
    assert list(inspect.signature(S.getPtsFromHeatmap).parameters) == ["heatmap", "conf_thresh", "nms_dist"]
    assert list(inspect.signature(S.box_nms).parameters) == ["prob", "size", "iou", "min_prob", "keep_top_k"]
    assert list(inspect.signature(S.compute_valid_mask).parameters)[:4] == ["image_shape", "inv_homography", "device", "erosion_radius"]
    # Train_model_heatmap_all.sem_loss(self, pred, label, device="cpu")
    sem = inspect.signature(S.utils.sem_loss)
    assert list(sem.parameters)[:3] == ["pred", "label", "device"] and sem.parameters["device"].default == "cpu"
    assert sem.parameters["ignore_index"].default == 133


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.warp_points(torch.zeros(4, 2), torch.eye(3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.descriptor_loss(torch.zeros(1, 256, 4, 4), torch.zeros(1, 256, 4, 4), torch.eye(3)[None])
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.detector_loss(torch.zeros(1, 65, 4, 4), torch.zeros(1, 65, 4, 4), torch.ones(1, 4, 4))
    with pytest.raises(NotImplementedError):
        S.box_nms(torch.zeros(8, 8), 4)


def test_shard_range_covers_everything():
    from ssp_b200.dist import shard_range
    for n in (0, 1, 7, 8, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_orchestration_call_sequences(monkeypatch):
    """The host side of the loss step with the C library faked out (every entry point recorded, nothing computed): the
    Python orchestration -- buffer shapes, autograd plumbing, engine dispatch, fused / unfused / semantic variants -- runs
    on a CPU-only box and issues the documented kernel sequence."""
    from ssp_b200 import _lib, losses, step, utils
    calls = []
    for m in (_lib, losses, utils):
        monkeypatch.setattr(m, "call", lambda name, *a: calls.append(name))
        monkeypatch.setattr(m, "stream_of", lambda t: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(utils, "_cuda_device", lambda device, *t: torch.device("cpu"))

    class NoFork(object):
        def __init__(self, dev):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def join(self):
            pass

    monkeypatch.setattr(losses, "_Fork", NoFork)
    monkeypatch.setattr(losses, "CHECK_LIST_OVERFLOW", False)  # reads a device counter the fake library never writes
    B, Hc, Wc = 2, 6, 8
    leaf = lambda *shape: torch.randn(*shape).requires_grad_()
    semi, semi_w, D, Dw = leaf(B, 65, Hc, Wc), leaf(B, 65, Hc, Wc), leaf(B, 256, Hc, Wc), leaf(B, 256, Hc, Wc)
    lab, m, H = torch.zeros(B, 1, Hc * 8, Wc * 8), torch.ones(B, 1, Hc * 8, Wc * 8), torch.eye(3).repeat(B, 1, 1)

    def run(**kw):
        del calls[:]
        out = step.loss_step(semi, semi_w, D, Dw, lab, lab, m, m, H, **kw)
        out["loss"].backward()
        return list(calls)

    fwd = ["ssp_detector_loss_fwd_pair", "ssp_desc_pack2_geometry", "ssp_desc_pos_fwd_planes",
           "ssp_desc_dense_fwd_tc", "ssp_desc_finalize"]
    # one-node fused step: the mask is folded into the indicator words, no pack pass in the backward, both GEMMs in one launch
    # backward of the one-node fused step: detector backward + coefficient / alpha / transpose blocks in ONE launch, then both GEMMs
    assert run() == fwd + ["ssp_step_bwd_prologue", "ssp_desc_bits_gemm_tc_pair"]
    # separately differentiable components (reference multi_task_loss weighting): general mask path with the pack pass
    bwd_desc = ["ssp_desc_alpha", "ssp_desc_pos_coef", "ssp_desc_pack", "ssp_desc_bits_gemm_tc_pair"]
    unfused = run(fused=False)
    assert unfused[:5] == fwd and sorted(unfused[5:]) == sorted(bwd_desc + ["ssp_detector_loss_bwd_pair"])
    assert [c for c in run(engine="fp32") if "_desc_" in c] == [
        "ssp_desc_geometry", "ssp_desc_pos_fwd", "ssp_desc_dense_fwd_simt", "ssp_desc_finalize", "ssp_desc_alpha",
        "ssp_desc_pos_coef", "ssp_desc_bits_gemm_simt", "ssp_desc_bits_gemm_simt", "ssp_desc_pos_apply"]
    assert "ssp_desc_pos_fwd" in run(engine="bf16") and "ssp_desc_pos_fwd_planes" not in run(engine="bf16")
    sp, sl = leaf(B, 133, Hc, Wc), torch.randint(0, 134, (B, Hc * 8, Wc * 8))
    sem = run(sem_pred=sp, sem=sl, sem_warp_pred=sp, warped_sem=sl)
    assert sem.count("ssp_sem_ce_up8") == 2 and sem.count("ssp_sem_ce_up8_bwd") == 2      # 1/8-resolution logits: fused upsample
    full = leaf(B, 133, Hc * 8, Wc * 8)
    sem = run(sem_pred=full, sem=sl, sem_warp_pred=full, warped_sem=sl)
    assert sem.count("ssp_sem_ce_fwd") == 2 and sem.count("ssp_sem_ce_bwd") == 2          # full-resolution logits
    with pytest.raises(RuntimeError, match="1/8"):
        utils.sem_loss(leaf(B, 133, Hc * 4, Wc * 4), sl)


REF = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REF), reason="reference checkout only exists in the authoring container")
def test_signatures_match_live_reference_and_dropin_binds():
    """Where the reference is available: same parameter names / defaults as its own callables, and install() binds over
    the real `utils.utils` module (then restores it)."""
    import importlib
    import inspect
    import sys
    sys.path.insert(0, REF)
    try:
        RU = importlib.import_module("utils.utils")
        for name in S.dropin.UTILS_NAMES:
            ref_sig, our_sig = inspect.signature(getattr(RU, name)), inspect.signature(getattr(S.utils, name))
            ref_p = [(p.name, p.default) for p in ref_sig.parameters.values() if p.kind != p.VAR_KEYWORD]
            our_p = [(p.name, p.default) for p in our_sig.parameters.values() if p.kind != p.VAR_KEYWORD]
            assert our_p[:len(ref_p)] == ref_p, name  # ours may only append optional extras
            assert all(d is not inspect.Parameter.empty for _, d in our_p[len(ref_p):]), name
        orig = RU.descriptor_loss
        bound = S.dropin.install(utils_module=RU)
        assert RU.descriptor_loss is S.utils.descriptor_loss and len(bound) >= len(S.dropin.UTILS_NAMES)
        S.dropin.uninstall()
        assert RU.descriptor_loss is orig
        # method twins in models/model_wrap.py (8a8 + 8f rank 3): same parameters after `self`, bound over the classes
        import collections
        import collections.abc
        collections.Mapping = collections.abc.Mapping
        MW = importlib.import_module("models.model_wrap")
        for cls, meth, ours in ((MW.SuperPointFrontend_torch, "sample_desc_from_points", S.utils.sample_desc_from_points),
                                (MW.PointTracker, "nn_match_two_way", S.utils.nn_match_two_way),
                                (MW.SuperPointFrontend_torch, "nms_fast", S.utils.nms_fast)):
            ref_p = [p.name for p in inspect.signature(getattr(cls, meth)).parameters.values()][1:]
            our_p = [p.name for p in inspect.signature(ours).parameters.values()]
            assert our_p[:len(ref_p)] == ref_p, meth
        saved = (MW.SuperPointFrontend_torch.sample_desc_from_points, MW.PointTracker.nn_match_two_way)
        bound = S.dropin.install(utils_module=RU, frontend_class=MW.SuperPointFrontend_torch, tracker_class=MW.PointTracker)
        assert any(b.endswith("nn_match_two_way") for b in bound) and any(b.endswith("sample_desc_from_points") for b in bound)
        assert MW.PointTracker.nn_match_two_way is not saved[1]
        S.dropin.uninstall()
        assert (MW.SuperPointFrontend_torch.sample_desc_from_points, MW.PointTracker.nn_match_two_way) == saved
    finally:
        sys.path.remove(REF)
        for m in [k for k in sys.modules if k in ("utils", "models") or k.startswith(("utils.", "models."))]:
            del sys.modules[m]
