"""Timeline trace of the two tcgen05 kernels (forward dense GEMM, backward indicator GEMM pair) at bench sizes.

    SSP_TRACE=1 python scripts/trace_desc.py gpurun_out/trace.npz

Uses the instrumented library (build.py, SSP_TRACE=1): lane 0 of one warp per role records (clock64 << 8 | tag) at the
waits of its pipeline; `scripts/trace_report.py` turns the dump into per-role wait / work cycle tables.
"""
import os
import sys

assert os.environ.get("SSP_TRACE") == "1", "run with SSP_TRACE=1"

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ssp_b200 import _lib, synth
from ssp_b200._lib import call, ptr, stream_of

B, Hc, Wc, Dch = 32, 30, 40, 256
Nc, Ncp = Hc * Wc, 1280
dev = "cuda"
lib = _lib.load()
D = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 1, smooth=0.3)).to(dev)
Dw = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 2, smooth=0.3)).to(dev)
mv = torch.ones((B, Ncp), device=dev)
PA = torch.empty((2, B, Ncp, Dch), dtype=torch.bfloat16, device=dev)  # hi, lo planes of D (lo above hi)
PB = torch.empty_like(PA)                                             # hi, lo planes of Dw
st = stream_of(D)
call("ssp_desc_pack2", ptr(D), ptr(Dw), None, B, Dch, Nc, ptr(PA[0]), ptr(PA[1]), ptr(PB[0]), ptr(PB[1]), st)
nneg = lib.ssp_desc_dense_tc_nblocks(B, Nc)
part = torch.empty((nneg, 2), dtype=torch.float64, device=dev)
bitsR = torch.empty((B, Ncp // 32, Ncp), dtype=torch.int32, device=dev)
bitsC = torch.empty_like(bitsR)
out = torch.empty((B, Dch, Nc), device=dev)
out2 = torch.empty((B, Dch, Nc), device=dev)
plist = torch.full((B, Ncp, 16), -1, dtype=torch.int32, device=dev)
plist[:, :Nc, 0] = torch.arange(Nc, device=dev, dtype=torch.int32)[None]
pcoef = torch.ones((B, Ncp, 16), device=dev)


def fwd(bits=True, split=True, transpose=True):
    call("ssp_desc_dense_fwd_tc", ptr(PA[0]), ptr(PA[1]) if split else None, ptr(PB[0]), ptr(PB[1]) if split else None, ptr(mv), None,
         B, Hc, Wc, 0.2, ptr(part), ptr(bitsR) if bits else None, ptr(bitsC) if (bits and transpose) else None, None, st)


def bwd1(pos=True, split=True):
    call("ssp_desc_bits_gemm_tc_planes", ptr(bitsR), ptr(PB[0]), ptr(PB[1]) if split else None, None, ptr(plist) if pos else None,
         ptr(pcoef) if pos else None, ptr(PB[0]), ptr(PB[1]) if split else None, B, Nc, ptr(out), st)


def bwd2(pos=True, split=True):
    lo = (lambda t: ptr(t[1])) if split else (lambda t: None)
    pl, pc = (ptr(plist), ptr(pcoef)) if pos else (None, None)
    call("ssp_desc_bits_gemm_tc_pair", ptr(bitsR), ptr(PB[0]), lo(PB), None, pl, pc, ptr(PB[0]), lo(PB), ptr(out),
         ptr(bitsC), ptr(PA[0]), lo(PA), ptr(mv), pl, pc, ptr(PA[0]), lo(PA), ptr(out2), B, Nc, st)

CAP = lib.ssp_debug_trace_cap()
flush = torch.empty((64 << 20,), dtype=torch.float32, device=dev)  # 256 MB > L2
res = {}
for name, fn in (("fwd", fwd), ("bwd", bwd2)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tr = torch.zeros((148, 4, CAP), dtype=torch.int64, device=dev)
    flush.zero_()
    call("ssp_debug_trace", ptr(tr))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    call("ssp_debug_trace", None)
    res[name] = tr.cpu().numpy()
    res[name + "_us"] = np.float64(1e3 * e0.elapsed_time(e1))
    print(name, "traced launch: %.1f us" % res[name + "_us"], flush=True)
np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.npz", **res)
