// Backward coefficients of the sparse positive pairs (+ alpha / srow vectors, + the bit-matrix transposes) as a device
// function over an argument block, so that the same body can run as its own launch (desc_common.cu) or as some of the blocks
// of the fused step's backward prologue next to the detector-loss backward (detector.cu).
#pragma once
#include "desc_common.cuh"

struct PosCoefArgs {
  const int* rowcol; const float* rowdot; const int* colcnt; const int* colrow; const float* coldot;
  const uint32_t* bitsR; const float* mv_pad; const float* g3; float gscale; int gmode; const float* out8;
  int Nc_pad; float lamda, mpos;
  float* rowcoef; int* colrow_sorted; float* colcoef; float* alpha_out; float* srow_out; uint32_t* bitsC_out; int NWv;
};

struct PosG {  // upstream gradients of (loss, pos, neg) and the normaliser, loaded once per thread
  float gl, gp, gn, norm;
};
__device__ __forceinline__ float pos_coef(float dot, float mv, const PosG& G, float lamda, float mpos) {
  float x = mpos - dot;
  float ind = x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f);
  return -lamda * ind * (G.gl * mv + G.gp) / G.norm;
}
// coefficient of the negative hinge of a column with mask value mv (what desc_alpha_kernel writes)
__device__ __forceinline__ float neg_alpha(float mv, const PosG& G) { return (G.gl * mv + G.gn) / G.norm; }

// N = number of list slots handled (4: vector loads, the common case; DESC_MAXP: any list).  Lists are filled front to
// back, so a row list with slot 3 empty / a column count <= 4 is complete within the first 4 slots.
template <int N>
__device__ __forceinline__ void pos_coef_rows(const int (&c4)[4], const float (&d4)[4], const int* __restrict__ rowcol,
                                              const float* __restrict__ rowdot, size_t base, int b, int cell, int Nc_pad, int NW,
                                              const uint32_t* __restrict__ bitsR, const float* __restrict__ mv_pad,
                                              const PosG& G, float lamda, float mpos, float* __restrict__ rowcoef) {
  int cc[N];
  float dd[N], cf[N];
#pragma unroll
  for (int n = 0; n < 4; ++n) { cc[n] = c4[n]; dd[n] = d4[n]; }
#pragma unroll
  for (int n = 4; n < N; n += 4) {
    int4 c = *reinterpret_cast<const int4*>(rowcol + base + n);
    float4 d = *reinterpret_cast<const float4*>(rowdot + base + n);
    cc[n] = c.x; cc[n + 1] = c.y; cc[n + 2] = c.z; cc[n + 3] = c.w;
    dd[n] = d.x; dd[n + 1] = d.y; dd[n + 2] = d.z; dd[n + 3] = d.w;
  }
  // the dependent gathers (mask value and indicator word of every partner) are issued together, then consumed
  float mvc[N];
  uint32_t wc[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const int c = cc[n];
    mvc[n] = c >= 0 ? __ldg(mv_pad + (size_t)b * Nc_pad + c) : 0.f;
    wc[n] = c >= 0 ? __ldg(bitsR + ((size_t)b * NW + (c >> 5)) * Nc_pad + cell) : 0u;
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float coef = 0.f;
    const int c = cc[n];
    if (c >= 0) {
      coef = pos_coef(dd[n], mvc[n], G, lamda, mpos);
      if ((wc[n] >> DESC_BITPOS(c & 31)) & 1u) coef -= neg_alpha(mvc[n], G);
    }
    cf[n] = coef;
  }
#pragma unroll
  for (int n = 0; n < N; n += 4)
    *reinterpret_cast<float4*>(rowcoef + base + n) = make_float4(cf[n], cf[n + 1], cf[n + 2], cf[n + 3]);
}

template <int N>
__device__ __forceinline__ void pos_coef_cols(int cnt, const int (&r4)[4], const float (&d4)[4], const int* __restrict__ colrow,
                                              const float* __restrict__ coldot, size_t base, int b, int cell, int Nc_pad, int NW,
                                              const uint32_t* __restrict__ bitsR, float mv, const PosG& G, float lamda,
                                              float mpos, int* __restrict__ colrow_sorted, float* __restrict__ colcoef) {
  int rr[N];
  float dd[N];
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    int rv[4];
    float dv[4];
    if (n == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { rv[u] = r4[u]; dv[u] = d4[u]; }
    } else {
      int4 r = *reinterpret_cast<const int4*>(colrow + base + n);
      float4 d = *reinterpret_cast<const float4*>(coldot + base + n);
      rv[0] = r.x; rv[1] = r.y; rv[2] = r.z; rv[3] = r.w;
      dv[0] = d.x; dv[1] = d.y; dv[2] = d.z; dv[3] = d.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      rr[n + u] = n + u < cnt ? rv[u] : 0x7fffffff;  // slots past the count were never written
      dd[n + u] = n + u < cnt ? dv[u] : 0.f;
    }
  }
  // sort by row index: deterministic summation order whatever order the forward's atomics handed out the slots in
#pragma unroll
  for (int i = 1; i < N; ++i)
#pragma unroll
    for (int j = N - 1; j >= 1; --j)
      if (j >= i && rr[j] < rr[j - 1]) {
        int t = rr[j]; rr[j] = rr[j - 1]; rr[j - 1] = t;
        float u = dd[j]; dd[j] = dd[j - 1]; dd[j - 1] = u;
      }
  uint32_t wr[N];
#pragma unroll
  for (int n = 0; n < N; ++n)
    wr[n] = n < cnt ? __ldg(bitsR + ((size_t)b * NW + (cell >> 5)) * Nc_pad + rr[n]) : 0u;
  const float al = neg_alpha(mv, G);
  float cf[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    cf[n] = 0.f;
    if (n < cnt) {
      float coef = pos_coef(dd[n], mv, G, lamda, mpos);
      if ((wr[n] >> DESC_BITPOS(cell & 31)) & 1u) coef -= al;
      cf[n] = coef;
    } else {
      rr[n] = -1;
    }
  }
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    *reinterpret_cast<int4*>(colrow_sorted + base + n) = make_int4(rr[n], rr[n + 1], rr[n + 2], rr[n + 3]);
    *reinterpret_cast<float4*>(colcoef + base + n) = make_float4(cf[n], cf[n + 1], cf[n + 2], cf[n + 3]);
  }
  // the consumers scan all DESC_MAXP slots of a list for entries >= 0
#pragma unroll
  for (int n = N; n < DESC_MAXP; n += 4) *reinterpret_cast<int4*>(colrow_sorted + base + n) = make_int4(-1, -1, -1, -1);
}

// bz = 0: row lists (+ the alpha / srow vectors of the backward GEMMs, when asked for); 1: column lists.  One thread
// per cell; the kernel is a chain of dependent gathers, so the first four slots of a list, the column count and the mask
// value are loaded before anything is looked at, and the second-level gathers of a list are issued together.
// gmode 0: g3 = {dL/dloss, dL/dpos, dL/dneg};  1: g3[0] = dL/dloss only (the fused step).  Everything is scaled by gscale.
// (bx, by, bz) = the block's coordinates in a (Nc_pad / 128, B, 2 or 6) grid of 128-thread blocks, gridx = Nc_pad / 128
__device__ __forceinline__ void desc_pos_coef_block(const PosCoefArgs& A, int bx, int by, int bz, int gridx) {
  const int* __restrict__ rowcol = A.rowcol; const float* __restrict__ rowdot = A.rowdot;
  const int* __restrict__ colcnt = A.colcnt; const int* __restrict__ colrow = A.colrow;
  const float* __restrict__ coldot = A.coldot; const uint32_t* __restrict__ bitsR = A.bitsR;
  const float* __restrict__ mv_pad = A.mv_pad; const float* __restrict__ g3 = A.g3; const float* __restrict__ out8 = A.out8;
  const float gscale = A.gscale, lamda = A.lamda, mpos = A.mpos;
  const int gmode = A.gmode, Nc_pad = A.Nc_pad, NWv = A.NWv;
  float* __restrict__ rowcoef = A.rowcoef; int* __restrict__ colrow_sorted = A.colrow_sorted;
  float* __restrict__ colcoef = A.colcoef; float* __restrict__ alpha_out = A.alpha_out; float* __restrict__ srow_out = A.srow_out;
  uint32_t* __restrict__ bitsC_out = A.bitsC_out;
  const int b = by;
  const int NW = Nc_pad / 32;
  if (bz >= 2) {
    // ---- blocks of bz = 2..5: bitsC = transpose of bitsR (only the backward GEMM reads it, so it is made here, inside a launch
    // whose other blocks are chains of dependent gathers: the streaming transposes fill the memory pipe they leave idle).
    //   bitsR[b][cw][r]: bit DESC_BITPOS(j) = indicator(row r, column 32 cw + j);  bitsC[b][rw][c]: bit DESC_BITPOS(i) =
    //   indicator(row 32 rw + i, column c).  One warp per 32 x 32 tile, lane L holds the word of row 32 rw + inv(L).
    const int rw = (bz - 2) * gridx + bx;  // gridx = Nc_pad / 128, four z slices: rw < NW
    if (rw >= NW) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int rl = ((lane & 15) << 1) | (lane >> 4);
    for (int cw0 = warp; cw0 < NW; cw0 += 4 * nwarp) {
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cw = cw0 + nwarp * u;
        // words beyond the valid range were never written by the forward: they transpose to zero
        w[u] = (cw < NWv && rw < NWv) ? __ldg(bitsR + ((size_t)b * NW + cw) * Nc_pad + rw * 32 + rl) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cw = cw0 + nwarp * u;
        if (cw >= NW) break;
        bitsC_out[((size_t)b * NW + rw) * Nc_pad + cw * 32 + rl] = warp_transpose32(w[u], lane);
      }
    }
    return;
  }
  const int cell = bx * blockDim.x + threadIdx.x;
  if (cell >= Nc_pad) return;
  const size_t base = ((size_t)b * Nc_pad + cell) * DESC_MAXP;
  const bool cols = bz != 0;
  // first-level loads, all independent
  const int4 l4 = *reinterpret_cast<const int4*>((cols ? colrow : rowcol) + base);
  const float4 f4 = *reinterpret_cast<const float4*>((cols ? coldot : rowdot) + base);
  const int cnt_raw = cols ? colcnt[(size_t)b * Nc_pad + cell] : 0;
  const float mv = mv_pad[(size_t)b * Nc_pad + cell];
  PosG G;
  G.gl = g3[0] * gscale;
  G.gp = gmode ? 0.f : g3[1] * gscale;
  G.gn = gmode ? 0.f : g3[2] * gscale;
  G.norm = out8[3];
  const int li[4] = {l4.x, l4.y, l4.z, l4.w};
  const float lf[4] = {f4.x, f4.y, f4.z, f4.w};
  if (!cols) {
    if (alpha_out) alpha_out[(size_t)b * Nc_pad + cell] = neg_alpha(mv, G);
    if (srow_out) srow_out[(size_t)b * Nc_pad + cell] = neg_alpha(1.f, G);
    // A cell has 0.8 partners on average (descriptor_dist 4 on an 8-pixel grid): the 4-slot path is the one that runs;
    // longer lists (descriptor_dist close to the cell size, strong minification) take the full one.
    if (li[3] < 0)
      pos_coef_rows<4>(li, lf, rowcol, rowdot, base, b, cell, Nc_pad, NW, bitsR, mv_pad, G, lamda, mpos, rowcoef);
    else
      pos_coef_rows<DESC_MAXP>(li, lf, rowcol, rowdot, base, b, cell, Nc_pad, NW, bitsR, mv_pad, G, lamda, mpos, rowcoef);
  } else {
    const int cnt = min(cnt_raw, DESC_MAXP);
    if (cnt <= 4)
      pos_coef_cols<4>(cnt, li, lf, colrow, coldot, base, b, cell, Nc_pad, NW, bitsR, mv, G, lamda, mpos, colrow_sorted, colcoef);
    else
      pos_coef_cols<DESC_MAXP>(cnt, li, lf, colrow, coldot, base, b, cell, Nc_pad, NW, bitsR, mv, G, lamda, mpos, colrow_sorted, colcoef);
  }
}

