// Descriptor-loss kernels shared by both GEMM engines: geometry, exact positive pairs (fwd/bwd),
// finalisation, the on-request 5-D pair mask, the bf16 hi/lo operand pack and the backward scales.
// Reference: utils/utils.py:779-893 (descriptor_loss), :745-768 (normPts/denormPts), :315-343.
#include "desc_common.cuh"
#include <cuda_bf16.h>

// ----------------------------------------------------------------------------------------------
// geometry: warped cell centres  w = denorm(swap(warp(swap(norm(c)))))   [utils/utils.py:829-851]
// ----------------------------------------------------------------------------------------------
__global__ void desc_geometry_kernel(const float* __restrict__ Hm, const float* __restrict__ mask_valid, int B,
                                     int Hc, int Wc, int cell, int Nc_pad, float2* __restrict__ wpts,
                                     float* __restrict__ mv_pad) {
  int b = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nc_pad) return;
  int Nc = Hc * Wc;
  float2 w = make_float2(SSP_FAR, SSP_FAR);
  float mv = 0.f;
  if (c < Nc) {
    float cx, cy;
    cell_center(c, Wc, cell, cx, cy);
    float Hpx = (float)(Hc * cell), Wpx = (float)(Wc * cell);
    // normPts divides by (H, W), not (H-1, W-1)   [utils/utils.py:745-755, :838]
    float ny = cy / Hpx * 2.f - 1.f;
    float nx = cx / Wpx * 2.f - 1.f;
    float ox, oy;
    homography_apply(Hm + b * 9, nx, ny, ox, oy);
    // denormPts: (p + 1) * shape / 2   [utils/utils.py:758-768]
    w.x = (ox + 1.f) * Wpx / 2.f;
    w.y = (oy + 1.f) * Hpx / 2.f;
    mv = mask_valid ? mask_valid[(size_t)b * Nc + c] : 1.f;
  }
  wpts[(size_t)b * Nc_pad + c] = w;
  mv_pad[(size_t)b * Nc_pad + c] = mv;
}

extern "C" int ssp_desc_geometry(const float* Hm, const float* mask_valid, int B, int Hc, int Wc, int cell,
                                 float* wpts, float* mv_pad, void* stream) {
  SSP_REQUIRE(Hm && wpts && mv_pad, "ssp_desc_geometry: null pointer");
  SSP_REQUIRE(B > 0 && B <= 65535 && Hc > 0 && Wc > 0 && cell > 0, "ssp_desc_geometry: bad sizes");
  int Nc_pad = desc_nc_pad(Hc * Wc);
  dim3 grid(ssp_ceil_div(Nc_pad, 128), B);
  desc_geometry_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Hm, mask_valid, B, Hc, Wc, cell, Nc_pad,
                                                                reinterpret_cast<float2*>(wpts), mv_pad);
  SSP_CUDA_CHECK_LAUNCH("desc_geometry_kernel");
  return SSP_OK;
}

// candidate window of cell indices whose centre can be within `dist` of (wx, wy); +-1 cell of slack,
// the exact predicate decides.
__device__ __forceinline__ void pos_window(float wx, float wy, float dist, int Hc, int Wc, int cell, int& k0,
                                           int& k1, int& l0, int& l1) {
  float half = (float)(cell / 2), fc = (float)cell;
  float a = floorf((wy - dist - half) / fc) - 1.f, b = ceilf((wy + dist - half) / fc) + 1.f;
  float c = floorf((wx - dist - half) / fc) - 1.f, d = ceilf((wx + dist - half) / fc) + 1.f;
  // clamp in float first: far-away / non-finite points give an empty window
  k0 = (int)fmaxf(a, 0.f);
  k1 = (int)fminf(b, (float)(Hc - 1));
  l0 = (int)fmaxf(c, 0.f);
  l1 = (int)fminf(d, (float)(Wc - 1));
  if (!(a <= (float)Hc && b >= -1.f && c <= (float)Wc && d >= -1.f)) { k0 = 1; k1 = 0; }
}

__device__ __forceinline__ float dot_exact(const float* __restrict__ a, const float* __restrict__ b, int Dch,
                                           size_t stride) {
  float acc = 0.f;
  for (int d = 0; d < Dch; ++d) acc = fmaf(__ldg(a + d * stride), __ldg(b + d * stride), acc);
  return acc;
}

// ----------------------------------------------------------------------------------------------
// positive pairs, forward:  partial sums of lamda * max(mpos - dot, 0)  (unweighted, mv-weighted)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
desc_pos_fwd_kernel(const float* __restrict__ D, const float* __restrict__ Dw, const float2* __restrict__ wpts,
                    const float* __restrict__ mv_pad, DescGeom g, double* __restrict__ partials) {
  __shared__ double sh[32];
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  double acc_u = 0.0, acc_w = 0.0;
  if (row < g.B * g.Nc) {
    int b = row / g.Nc, ij = row - b * g.Nc;
    float2 w = wpts[(size_t)b * g.Nc_pad + ij];
    int k0, k1, l0, l1;
    pos_window(w.x, w.y, g.dist, g.Hc, g.Wc, g.cell, k0, k1, l0, l1);
    const float* Db = D + (size_t)b * g.Dch * g.Nc + ij;
    const float* Dwb = Dw + (size_t)b * g.Dch * g.Nc;
    for (int k = k0; k <= k1; ++k)
      for (int l = l0; l <= l1; ++l) {
        int c = k * g.Wc + l;
        float cx, cy;
        cell_center(c, g.Wc, g.cell, cx, cy);
        if (!pair_positive(w.x, w.y, cx, cy, g.dist)) continue;
        float dot = dot_exact(Db, Dwb + c, g.Dch, g.Nc);
        float pos = g.lamda * fmaxf(g.mpos - dot, 0.f);
        acc_u += (double)pos;
        acc_w += (double)(pos * mv_pad[(size_t)b * g.Nc_pad + c]);
      }
  }
  double ru = block_sum_d(acc_u, sh);
  double rw = block_sum_d(acc_w, sh);
  if (threadIdx.x == 0) {
    partials[2 * (size_t)blockIdx.x] = ru;
    partials[2 * (size_t)blockIdx.x + 1] = rw;
  }
}

extern "C" int ssp_desc_pos_nblocks(int B, int Nc) { return ssp_ceil_div(B * Nc, 128); }

static int fill_geom(DescGeom& g, int B, int Hc, int Wc, int Dch, int cell, float dist, float lamda, float mpos,
                     float mneg) {
  g.B = B; g.Hc = Hc; g.Wc = Wc; g.Nc = Hc * Wc; g.Nc_pad = desc_nc_pad(Hc * Wc); g.Dch = Dch;
  g.cell = cell; g.dist = dist; g.lamda = lamda; g.mpos = mpos; g.mneg = mneg;
  return (B > 0 && Hc > 0 && Wc > 0 && Dch > 0 && cell > 0) ? 0 : -1;
}

extern "C" int ssp_desc_pos_fwd(const float* D, const float* Dw, const float* wpts, const float* mv_pad, int B,
                                int Hc, int Wc, int Dch, int cell, float dist, float lamda, float mpos,
                                double* partials, void* stream) {
  SSP_REQUIRE(D && Dw && wpts && mv_pad && partials, "ssp_desc_pos_fwd: null pointer");
  DescGeom g;
  SSP_REQUIRE(fill_geom(g, B, Hc, Wc, Dch, cell, dist, lamda, mpos, 0.f) == 0, "ssp_desc_pos_fwd: bad sizes");
  desc_pos_fwd_kernel<<<ssp_desc_pos_nblocks(B, g.Nc), 128, 0, (cudaStream_t)stream>>>(
      D, Dw, reinterpret_cast<const float2*>(wpts), mv_pad, g, partials);
  SSP_CUDA_CHECK_LAUNCH("desc_pos_fwd_kernel");
  return SSP_OK;
}

// ----------------------------------------------------------------------------------------------
// finalize: fixed-order sums of the partials, global-batch normaliser  [utils/utils.py:883-890]
//   out8 = { loss_desc, pos_sum, neg_sum, normalization, num_loss, num_pos, num_neg, sum(mask_valid) }
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
desc_finalize_kernel(const double* __restrict__ pos_part, int npos, const double* __restrict__ neg_part, int nneg,
                     const float* __restrict__ mv_pad, size_t nmv, int B, int Hc, int Wc, float* __restrict__ out4) {
  __shared__ double sh[32];
  double pu = 0, pw = 0, nu = 0, nw = 0, sm = 0;
  for (int i = threadIdx.x; i < npos; i += blockDim.x) { pu += pos_part[2 * i]; pw += pos_part[2 * i + 1]; }
  for (int i = threadIdx.x; i < nneg; i += blockDim.x) { nu += neg_part[2 * i]; nw += neg_part[2 * i + 1]; }
  for (size_t i = threadIdx.x; i < nmv; i += blockDim.x) sm += (double)mv_pad[i];
  pu = block_sum_d(pu, sh);
  pw = block_sum_d(pw, sh);
  nu = block_sum_d(nu, sh);
  nw = block_sum_d(nw, sh);
  sm = block_sum_d(sm, sh);
  if (threadIdx.x == 0) {
    float norm = (float)B * ((float)sm + 1.f) * (float)Hc * (float)Wc;
    out4[0] = (float)((pw + nw) / (double)norm);
    out4[1] = (float)(pu / (double)norm);
    out4[2] = (float)(nu / (double)norm);
    out4[3] = norm;
    // raw sums for the multi-GPU exchange (global-batch normaliser)
    out4[4] = (float)(pw + nw);
    out4[5] = (float)pu;
    out4[6] = (float)nu;
    out4[7] = (float)sm;
  }
}

extern "C" int ssp_desc_finalize(const double* pos_part, int npos, const double* neg_part, int nneg,
                                 const float* mv_pad, int B, int Hc, int Wc, float* out4, void* stream) {
  SSP_REQUIRE(pos_part && neg_part && mv_pad && out4, "ssp_desc_finalize: null pointer");
  SSP_REQUIRE(npos >= 0 && nneg >= 0 && B > 0 && Hc > 0 && Wc > 0, "ssp_desc_finalize: bad sizes");
  size_t nmv = (size_t)B * desc_nc_pad(Hc * Wc);
  desc_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pos_part, npos, neg_part, nneg, mv_pad, nmv, B, Hc, Wc, out4);
  SSP_CUDA_CHECK_LAUNCH("desc_finalize_kernel");
  return SSP_OK;
}

// ----------------------------------------------------------------------------------------------
// the 5-D correspondence mask [B,Hc,Wc,Hc,Wc] (float 0/1) -- returned by the reference but never read
// by its caller; materialised only on request.
// ----------------------------------------------------------------------------------------------
__global__ void desc_pair_mask_kernel(const float2* __restrict__ wpts, DescGeom g, float* __restrict__ mask) {
  int b = blockIdx.z, r = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.Nc) return;
  float2 w = wpts[(size_t)b * g.Nc_pad + r];
  float cx, cy;
  cell_center(c, g.Wc, g.cell, cx, cy);
  mask[((size_t)b * g.Nc + r) * g.Nc + c] = pair_positive(w.x, w.y, cx, cy, g.dist) ? 1.f : 0.f;
}

extern "C" int ssp_desc_pair_mask(const float* wpts, int B, int Hc, int Wc, int cell, float dist, float* mask,
                                  void* stream) {
  SSP_REQUIRE(wpts && mask, "ssp_desc_pair_mask: null pointer");
  DescGeom g;
  SSP_REQUIRE(fill_geom(g, B, Hc, Wc, 1, cell, dist, 0.f, 0.f, 0.f) == 0 && B <= 65535 && g.Nc <= 65535,
              "ssp_desc_pair_mask: bad sizes");
  dim3 grid(ssp_ceil_div(g.Nc, 256), g.Nc, B);
  desc_pair_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(wpts), g, mask);
  SSP_CUDA_CHECK_LAUNCH("desc_pair_mask_kernel");
  return SSP_OK;
}

// ----------------------------------------------------------------------------------------------
// backward scales: alpha[b,c] = (g_loss * mv[b,c] + g_neg) / norm   (coefficient of the negative hinge)
// ----------------------------------------------------------------------------------------------
__global__ void desc_alpha_kernel(const float* __restrict__ mv_pad, const float* __restrict__ g3,
                                  const float* __restrict__ out4, size_t n, float* __restrict__ alpha) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) alpha[i] = (g3[0] * mv_pad[i] + g3[2]) / out4[3];
}

extern "C" int ssp_desc_alpha(const float* mv_pad, const float* g3, const float* out4, int B, int Nc_pad,
                              float* alpha, void* stream) {
  SSP_REQUIRE(mv_pad && g3 && out4 && alpha, "ssp_desc_alpha: null pointer");
  size_t n = (size_t)B * Nc_pad;
  desc_alpha_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mv_pad, g3, out4, n, alpha);
  SSP_CUDA_CHECK_LAUNCH("desc_alpha_kernel");
  return SSP_OK;
}

// ----------------------------------------------------------------------------------------------
// positive pairs, backward.  Runs after the dense indicator-GEMMs stored dD / dDw.
//   coef = -lamda * ind(mpos - dot) * (g_loss * mv[c] + g_pos) / norm,  ind = 1 (x>0), 0.5 (x==0), 0
//   rows  (thread per ij): dD [b,:,ij] += coef * Dw[b,:,c]
//   cols  (thread per c, brute-force scan of all rows): dDw[b,:,c] += coef * D[b,:,ij]
// Each output column has exactly one writer, so the result is deterministic (no atomics).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float pos_coef(float dot, float mv, const float* __restrict__ g3, float norm,
                                          const DescGeom& g) {
  float x = g.mpos - dot;
  float ind = x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f);
  return -g.lamda * ind * (g3[0] * mv + g3[1]) / norm;
}

__global__ void __launch_bounds__(128)
desc_pos_bwd_rows_kernel(const float* __restrict__ D, const float* __restrict__ Dw,
                         const float2* __restrict__ wpts, const float* __restrict__ mv_pad,
                         const float* __restrict__ g3, const float* __restrict__ out4, DescGeom g,
                         float* __restrict__ dD) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= g.B * g.Nc) return;
  int b = row / g.Nc, ij = row - b * g.Nc;
  float2 w = wpts[(size_t)b * g.Nc_pad + ij];
  int k0, k1, l0, l1;
  pos_window(w.x, w.y, g.dist, g.Hc, g.Wc, g.cell, k0, k1, l0, l1);
  const float* Db = D + (size_t)b * g.Dch * g.Nc + ij;
  const float* Dwb = Dw + (size_t)b * g.Dch * g.Nc;
  float* dDb = dD + (size_t)b * g.Dch * g.Nc + ij;
  float norm = out4[3];
  for (int k = k0; k <= k1; ++k)
    for (int l = l0; l <= l1; ++l) {
      int c = k * g.Wc + l;
      float cx, cy;
      cell_center(c, g.Wc, g.cell, cx, cy);
      if (!pair_positive(w.x, w.y, cx, cy, g.dist)) continue;
      float dot = dot_exact(Db, Dwb + c, g.Dch, g.Nc);
      float coef = pos_coef(dot, mv_pad[(size_t)b * g.Nc_pad + c], g3, norm, g);
      if (coef == 0.f) continue;
      for (int d = 0; d < g.Dch; ++d) dDb[(size_t)d * g.Nc] += coef * __ldg(Dwb + c + (size_t)d * g.Nc);
    }
}

__global__ void __launch_bounds__(128)
desc_pos_bwd_cols_kernel(const float* __restrict__ D, const float* __restrict__ Dw,
                         const float2* __restrict__ wpts, const float* __restrict__ mv_pad,
                         const float* __restrict__ g3, const float* __restrict__ out4, DescGeom g,
                         float* __restrict__ dDw) {
  __shared__ float2 sw[128];
  int b = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = c < g.Nc;
  float cx = 0.f, cy = 0.f, mv = 0.f;
  if (active) {
    cell_center(c, g.Wc, g.cell, cx, cy);
    mv = mv_pad[(size_t)b * g.Nc_pad + c];
  }
  const float* Db = D + (size_t)b * g.Dch * g.Nc;
  const float* Dwb = Dw + (size_t)b * g.Dch * g.Nc + c;
  float* dDwb = dDw + (size_t)b * g.Dch * g.Nc + c;
  float norm = out4[3];
  for (int r0 = 0; r0 < g.Nc; r0 += 128) {
    __syncthreads();
    int r = r0 + threadIdx.x;
    sw[threadIdx.x] = r < g.Nc ? wpts[(size_t)b * g.Nc_pad + r] : make_float2(SSP_FAR, SSP_FAR);
    __syncthreads();
    if (!active) continue;
    int lim = min(128, g.Nc - r0);
    for (int q = 0; q < lim; ++q) {
      float2 w = sw[q];
      if (!pair_positive(w.x, w.y, cx, cy, g.dist)) continue;
      int ij = r0 + q;
      float dot = dot_exact(Db + ij, Dwb, g.Dch, g.Nc);
      float coef = pos_coef(dot, mv, g3, norm, g);
      if (coef == 0.f) continue;
      for (int d = 0; d < g.Dch; ++d) dDwb[(size_t)d * g.Nc] += coef * __ldg(Db + ij + (size_t)d * g.Nc);
    }
  }
}

extern "C" int ssp_desc_pos_bwd(const float* D, const float* Dw, const float* wpts, const float* mv_pad,
                                const float* g3, const float* out4, int B, int Hc, int Wc, int Dch, int cell,
                                float dist, float lamda, float mpos, float* dD, float* dDw, void* stream) {
  SSP_REQUIRE(D && Dw && wpts && mv_pad && g3 && out4 && dD && dDw, "ssp_desc_pos_bwd: null pointer");
  DescGeom g;
  SSP_REQUIRE(fill_geom(g, B, Hc, Wc, Dch, cell, dist, lamda, mpos, 0.f) == 0 && B <= 65535, "ssp_desc_pos_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  desc_pos_bwd_rows_kernel<<<ssp_ceil_div(B * g.Nc, 128), 128, 0, st>>>(
      D, Dw, reinterpret_cast<const float2*>(wpts), mv_pad, g3, out4, g, dD);
  SSP_CUDA_CHECK_LAUNCH("desc_pos_bwd_rows_kernel");
  dim3 grid(ssp_ceil_div(g.Nc, 128), B);
  desc_pos_bwd_cols_kernel<<<grid, 128, 0, st>>>(D, Dw, reinterpret_cast<const float2*>(wpts), mv_pad, g3, out4, g, dDw);
  SSP_CUDA_CHECK_LAUNCH("desc_pos_bwd_cols_kernel");
  return SSP_OK;
}

// ----------------------------------------------------------------------------------------------
// operand pack for the tcgen05 engine: [B, Dch, Nc] fp32 (NCHW) -> K-major bf16 planes
//   hi[b, c, d] = bf16(s * x),  lo[b, c, d] = bf16(s * x - hi)   (lo optional), rows c >= Nc zeroed
//   s = scale[b, c] if given (backward: alpha), else 1.
// 32 cells x Dch tile transposed through shared memory; reads are 128 B per warp per channel, writes
// are contiguous bf16 rows.
// ----------------------------------------------------------------------------------------------
#define PK_CELLS 32
__global__ void __launch_bounds__(256)
desc_pack_kernel(const float* __restrict__ src, const float* __restrict__ scale, int Dch, int Nc, int Nc_pad,
                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  extern __shared__ float tile[];  // [PK_CELLS][Dch + 1]
  int b = blockIdx.y;
  int c0 = blockIdx.x * PK_CELLS;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int ld = Dch + 1;
  int c = c0 + lane;
  float s = 1.f;
  if (scale && c < Nc) s = scale[(size_t)b * Nc_pad + c];
  for (int d = w; d < Dch; d += 8) {
    float v = 0.f;
    if (c < Nc) v = __ldg(src + ((size_t)b * Dch + d) * Nc + c) * s;
    tile[lane * ld + d] = v;
  }
  __syncthreads();
  for (int cc = w; cc < PK_CELLS; cc += 8) {
    size_t o = ((size_t)b * Nc_pad + c0 + cc) * Dch;
    for (int d = lane; d < Dch; d += 32) {
      float v = tile[cc * ld + d];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[o + d] = h;
      if (lo) lo[o + d] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

extern "C" int ssp_desc_pack(const float* src, const float* scale, int B, int Dch, int Nc, void* hi, void* lo,
                             void* stream) {
  SSP_REQUIRE(src && hi, "ssp_desc_pack: null pointer");
  SSP_REQUIRE(B > 0 && B <= 65535 && Dch > 0 && Nc > 0, "ssp_desc_pack: bad sizes");
  int Nc_pad = desc_nc_pad(Nc);
  size_t smem = (size_t)PK_CELLS * (Dch + 1) * sizeof(float);
  SSP_REQUIRE(smem <= 48 * 1024, "ssp_desc_pack: descriptor dim %d too large", Dch);
  dim3 grid(Nc_pad / PK_CELLS, B);
  desc_pack_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(src, scale, Dch, Nc, Nc_pad, (__nv_bfloat16*)hi,
                                                              (__nv_bfloat16*)lo);
  SSP_CUDA_CHECK_LAUNCH("desc_pack_kernel");
  return SSP_OK;
}
