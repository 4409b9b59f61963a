"""Micro-benchmark of the tcgen05 kernels in isolation (CUDA events, 20 launches each)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssp_b200 as S
from ssp_b200 import _lib, synth
from ssp_b200._lib import call, ptr, stream_of
B, Hc, Wc, Dch = 32, 30, 40, 256
Nc, Ncp = Hc * Wc, 1280
dev = "cuda"
lib = _lib.load()
D = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 1, smooth=0.3)).to(dev)
Dw = torch.from_numpy(synth.unit_descriptors(B, Dch, Hc, Wc, 2, smooth=0.3)).to(dev)
mv = torch.ones((B, Ncp), device=dev)
planes = [torch.empty((B, Ncp, Dch), dtype=torch.bfloat16, device=dev) for _ in range(4)]
st = stream_of(D)
call("ssp_desc_pack2", ptr(D), ptr(Dw), None, B, Dch, Nc, ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(planes[3]), st)
nneg = lib.ssp_desc_dense_tc_nblocks(B, Nc)
part = torch.empty((nneg, 2), dtype=torch.float64, device=dev)
bitsR = torch.empty((B, Ncp // 32, Ncp), dtype=torch.int32, device=dev)
bitsC = torch.empty_like(bitsR)
out = torch.empty((B, Dch, Nc), device=dev)
plist = torch.full((B, Ncp, 16), -1, dtype=torch.int32, device=dev)
plist[:, :Nc, 0] = torch.arange(Nc, device=dev, dtype=torch.int32)[None]
pcoef = torch.ones((B, Ncp, 16), device=dev)

def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-44s %8.1f us" % (name, 1e3 * e0.elapsed_time(e1) / n), flush=True)

timeit("fwd x3 bits", lambda: call("ssp_desc_dense_fwd_tc", ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(planes[3]), ptr(mv), B, Hc, Wc, 0.2, ptr(part), ptr(bitsR), ptr(bitsC), None, st))
timeit("fwd x3 nobits", lambda: call("ssp_desc_dense_fwd_tc", ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(planes[3]), ptr(mv), B, Hc, Wc, 0.2, ptr(part), None, None, None, st))
timeit("fwd x1 bits", lambda: call("ssp_desc_dense_fwd_tc", ptr(planes[0]), None, ptr(planes[2]), None, ptr(mv), B, Hc, Wc, 0.2, ptr(part), ptr(bitsR), ptr(bitsC), None, st))
timeit("bwd gemm x3 + pos", lambda: call("ssp_desc_bits_gemm_tc", ptr(bitsR), ptr(planes[2]), ptr(planes[3]), None, ptr(plist), ptr(pcoef), ptr(Dw), B, Nc, ptr(out), st))
timeit("bwd gemm x3 no pos", lambda: call("ssp_desc_bits_gemm_tc", ptr(bitsR), ptr(planes[2]), ptr(planes[3]), None, None, None, None, B, Nc, ptr(out), st))
timeit("bwd gemm x1 + pos", lambda: call("ssp_desc_bits_gemm_tc", ptr(bitsR), ptr(planes[2]), None, None, ptr(plist), ptr(pcoef), ptr(Dw), B, Nc, ptr(out), st))
timeit("bwd gemm x1 no pos", lambda: call("ssp_desc_bits_gemm_tc", ptr(bitsR), ptr(planes[2]), None, None, None, None, None, B, Nc, ptr(out), st))
bitsR.zero_()
timeit("bwd gemm x3 no pos, all-zero bits", lambda: call("ssp_desc_bits_gemm_tc", ptr(bitsR), ptr(planes[2]), ptr(planes[3]), None, None, None, None, B, Nc, ptr(out), st))
