"""Multi-GPU path on real devices: the peer-memory exchange kernel and the sharded loss step.

* protocol test, one process: two "ranks" = two exchange buffers on the same device driven from two streams.
* parity test, two processes (one per GPU when the box has >= 2, otherwise both on cuda:0 -- the exchange buffers are
  cudaIpc-mapped either way; rendezvous over gloo): the sharded `loss_step` (CUDA kernels + exchange kernel, eager and
  as a captured CUDA graph) must equal the single-GPU loss step on the concatenated batch: same global-batch
  normalisers as the reference's single-GPU formulas (utils/utils.py:886-887, Train_model_heatmap_all.py:178).
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import ssp_b200 as S
from ssp_b200 import _lib, synth

pytestmark = pytest.mark.gpu
HC, WC = 30, 40


@pytest.mark.parametrize("world", [2, 8])
def test_exchange_kernel_ranks_in_one_process(world):
    """`world` "ranks" = `world` exchange buffers on one device, each driven from its own stream (the kernels spin on each
    other's flags, so they must be co-resident: one small CTA each)."""
    lib = _lib.load()
    dev = torch.device("cuda")
    bufs = []
    for _ in range(world):
        p = ctypes.c_void_p()
        _lib.check(lib.ssp_xchg_alloc(ctypes.byref(p), None), "ssp_xchg_alloc")
        bufs.append(p.value)
    table = (ctypes.c_void_p * world)(*bufs)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(0)
    try:
        for it in range(5):  # several exchanges: both parities of the double-buffered slots, flags reused
            loc = []
            for r in range(world):
                det0 = torch.tensor(rng.random(3) + 1.0, dtype=torch.float32, device=dev)
                det1 = torch.tensor(rng.random(3) + 1.0, dtype=torch.float32, device=dev)
                d8 = torch.tensor(rng.random(8) + 1.0, dtype=torch.float32, device=dev)
                sem = torch.tensor(rng.random(3) + 1.0, dtype=torch.float32, device=dev)
                loc.append((det0, det1, d8, sem))
            ref = [[t.clone().cpu().numpy().astype(np.float64) for t in l] for l in loc]
            Bl = [3 - (r % 2) for r in range(world)]  # uneven shards
            totals = [torch.zeros(1, device=dev) for _ in range(world)]
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    det0, det1, d8, sem = loc[r]
                    _lib.call("ssp_loss_exchange", table, r, world, _lib.ptr(det0), _lib.ptr(det1), _lib.ptr(d8), _lib.ptr(sem), None,
                              Bl[r], HC, WC, 0.25, _lib.ptr(totals[r]), 10.0, ctypes.c_void_p(streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(world):
                _lib.check(lib.ssp_xchg_status(ctypes.c_void_p(bufs[r]), None), "exchange status")
            # expected global values
            for i in (0, 1):
                num = sum(ref[r][i][1] for r in range(world))
                den = sum(ref[r][i][2] - 1e-5 for r in range(world)) + 1e-5
                for r in range(world):
                    got = loc[r][i].cpu().numpy()
                    np.testing.assert_allclose(got, [num / den, num, den], rtol=3e-6)
            sums = sum(ref[r][2][4:8] for r in range(world))
            norm = float(sum(Bl)) * (sums[3] + 1.0) * HC * WC
            ssem = sum(ref[r][3][1:3] for r in range(world))
            for r in range(world):
                got = loc[r][2].cpu().numpy()
                np.testing.assert_allclose(got[:3], sums[:3] / norm, rtol=3e-6)
                np.testing.assert_allclose(got[3], norm, rtol=1e-6)
                np.testing.assert_allclose(got[4:], sums, rtol=2e-6)
                np.testing.assert_allclose(loc[r][3].cpu().numpy(), [ssem[0] / ssem[1], ssem[0], ssem[1]], rtol=3e-6)
                assert torch.equal(loc[0][2], loc[r][2])  # bit-identical on all ranks (same summation order)
            for r in range(world):  # weighted total of the fused step: det0 + det1 + lambda_loss * desc
                want = float(loc[r][0][0]) + float(loc[r][1][0]) + 0.25 * float(loc[r][2][0])
                np.testing.assert_allclose(float(totals[r]), want, rtol=1e-6)
        # world = 1 degenerates to the local fix-up
        d8 = torch.tensor([0, 0, 0, 0, 10.0, 4.0, 6.0, 7.0], dtype=torch.float32, device=dev)
        one = (ctypes.c_void_p * 1)(bufs[0])
        _lib.call("ssp_loss_exchange", one, 0, 1, None, None, _lib.ptr(d8), None, None, 2, 3, 4, 1.0, None, 10.0, None)
        norm = 2.0 * 8.0 * 12.0
        np.testing.assert_allclose(d8.cpu().numpy(), [10 / norm, 4 / norm, 6 / norm, norm, 10, 4, 6, 7], rtol=1e-6)
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            lib.ssp_xchg_free(ctypes.c_void_p(b))


def test_exchange_timeout_poisons_instead_of_hanging():
    """A peer that never arrives: the kernel gives up after timeout_s, writes NaN and raises the sticky error."""
    lib = _lib.load()
    bufs = []
    for _ in range(2):
        p = ctypes.c_void_p()
        _lib.check(lib.ssp_xchg_alloc(ctypes.byref(p), None), "ssp_xchg_alloc")
        bufs.append(p.value)
    table = (ctypes.c_void_p * 2)(*bufs)
    det0 = torch.ones(3, device="cuda")
    _lib.call("ssp_loss_exchange", table, 0, 2, _lib.ptr(det0), None, None, None, None, 1, HC, WC, 1.0, None, 0.05, None)
    torch.cuda.synchronize()
    assert torch.isnan(det0).all()
    assert lib.ssp_xchg_status(ctypes.c_void_p(bufs[0]), None) != 0
    assert b"timed out" in lib.ssp_last_error()
    for b in bufs:
        lib.ssp_xchg_free(ctypes.c_void_p(b))


# ------------------------------------------------------------------------------------------------
def _step_inputs(B, seed):
    rng = np.random.default_rng(seed)
    Hs = np.stack([np.linalg.inv(synth.sample_homography(rng)) for _ in range(B)]).astype(np.float32)
    return {
        "semi": synth.pseudo_normal((B, 65, HC, WC), seed + 1), "semi_warp": synth.pseudo_normal((B, 65, HC, WC), seed + 2),
        "desc": synth.unit_descriptors(B, 256, HC, WC, seed + 3, smooth=0.3),
        "desc_warp": synth.unit_descriptors(B, 256, HC, WC, seed + 4, smooth=0.3),
        "labels_2D": synth.keypoint_labels(B, HC * 8, WC * 8, seed + 5), "warped_labels": synth.keypoint_labels(B, HC * 8, WC * 8, seed + 6),
        "mat_H": Hs, "inv_H": np.linalg.inv(Hs).astype(np.float32),
    }


def _run_step(inp, lo, hi, dev, group, graph, fused=True):
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev)
    d = {k: cu(v) for k, v in inp.items()}
    d["mask_2D"] = torch.ones((hi - lo, 1, HC * 8, WC * 8), device=dev)
    d["mask_warp_2D"] = S.compute_valid_mask(torch.tensor([HC * 8, WC * 8]), d["inv_H"], device=dev, erosion_radius=3).unsqueeze(1)
    if graph:
        g = S.step.GraphedLossStep(d, dist_group=group, fused=fused)
        for _ in range(3):
            out = g.replay()
        torch.cuda.synchronize()
        grads = [x.clone() for x in out["grads"]]
    else:
        leaves = [d[k].detach().requires_grad_(True) for k in ("semi", "semi_warp", "desc", "desc_warp")]
        out = S.step.loss_step(leaves[0], leaves[1], leaves[2], leaves[3], d["labels_2D"], d["warped_labels"], d["mask_2D"],
                               d["mask_warp_2D"], d["mat_H"], dist_group=group, fused=fused)
        out["loss"].backward()
        grads = [l.grad for l in leaves]
    torch.cuda.synchronize()
    keys = ("loss", "loss_det", "loss_det_warp", "loss_desc", "positive_dist", "negative_dist")
    return {k: float(out[k]) for k in keys}, [g.cpu().numpy() for g in grads]


def _worker(rank, world, port, ngpu, B, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank % ngpu))
        import torch.distributed as tdist
        from ssp_b200 import dist as sdist
        torch.cuda.set_device(rank % ngpu)
        dev = torch.device("cuda", rank % ngpu)
        tdist.init_process_group(backend="gloo", rank=rank, world_size=world)
        ex = sdist.get_exchange(True)
        assert ex.backend == "p2p"
        inp = _step_inputs(B, 40)
        lo, hi = sdist.shard_range(B, rank, world)
        res = {}
        for name, graph, fused in (("eager", False, True), ("graph", True, True), ("unfused", False, False)):
            res[name] = _run_step(inp, lo, hi, dev, True, graph, fused)
        ex.check()
        if rank == 0:
            res["single"] = _run_step(inp, 0, B, dev, None, False)
        q.put((rank, lo, hi, res, None))
        tdist.barrier()
        sdist.close_exchanges()
        tdist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, 0, 0, None, traceback.format_exc()))
        raise e


@pytest.mark.parametrize("B", [4, 5])
def test_sharded_loss_step_equals_single_gpu(B):
    """2 ranks; B = 5 gives uneven shards (3 + 2 pairs)."""
    ngpu = torch.cuda.device_count()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, min(ngpu, 2), B, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in range(2)]
    [p.join(timeout=120) for p in procs]
    for g in got:
        assert g[4] is None, g[4]
    assert all(p.exitcode == 0 for p in procs)
    got.sort()
    single_vals, single_grads = got[0][3]["single"]
    for rank, lo, hi, res, _ in got:
        for name in ("eager", "graph", "unfused"):
            vals, grads = res[name]
            for k, v in vals.items():
                np.testing.assert_allclose(v, single_vals[k], rtol=2e-5, err_msg="%s %s rank %d" % (name, k, rank))
            for gsh, gfull in zip(grads, single_grads):
                ref = gfull[lo:hi]
                scale = np.abs(ref).max()
                np.testing.assert_allclose(gsh, ref, rtol=1e-4, atol=1e-5 * scale, err_msg="%s rank %d" % (name, rank))
