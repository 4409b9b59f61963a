// Semantic-head cross entropy: the SSp training step's third loss.
// Reference (Gabriel-SGama/Semantic-SuperPoint):
//   models/SuperPointNet_gauss2_ssmall.py:86-91   sem = F.interpolate(convSout(..), x_hw, mode="bilinear", align_corners=False)
//   Train_model_heatmap_all.py:181-193            nn.CrossEntropyLoss(ignore_index=133)(pred [B,133,H,W], label [B,H,W] int64)
//
// Two paths behind the same loss:
//   * ssp_sem_ce_fwd / _bwd      pred is already full resolution (unmodified model): streaming softmax-CE over NCHW,
//                                one thread per pixel, channel-strided coalesced reads.  HBM-bound: C*H*W*4 bytes read
//                                forward, the same read plus the same written backward.
//   * ssp_sem_ce_up8 / _up8_bwd  pred is the LOW-RES head output [B,C,H/8,W/8]: the x8 bilinear upsample is fused
//                                into the loss, so the [B,133,H,W] logits (40.9 MB per 240x320 image, forward and
//                                backward) never exist.  One exponential per (pixel, channel) serves both the loss and
//                                the gradient, which is accumulated straight into the low-res layout.
// fp32 throughout; ex2/lg2.approx (2 ulp) for the exponentials.
#include "common.cuh"

#define SEM_LOG2E 1.4426950408889634f
#define SEM_LN2 0.6931471805599453f

__device__ __forceinline__ float sem_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// per-block {loss sum, valid count} -> out3 = {sum / count, sum, count}; 0/0 = NaN like the reference's mean over nothing
__global__ void __launch_bounds__(256) sem_finalize_kernel(const double* __restrict__ partials, int nblk, float* __restrict__ out3) {
  __shared__ double sh[32];
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    s += partials[2 * (size_t)i];
    c += partials[2 * (size_t)i + 1];
  }
  s = block_sum_d(s, sh);
  c = block_sum_d(c, sh);
  if (threadIdx.x == 0) {
    out3[0] = (float)(s / c);
    out3[1] = (float)s;
    out3[2] = (float)c;
  }
}

// ----------------------------------------------------------------------------------------------
// full-resolution path
// ----------------------------------------------------------------------------------------------
#define SEMF_THREADS 256
#define SEMF_U 8

__global__ void __launch_bounds__(SEMF_THREADS)
sem_ce_full_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int C, int HW,
                       long long npix, int ignore, float* __restrict__ lse2, double* __restrict__ partials) {
  long long pix = (long long)blockIdx.x * SEMF_THREADS + threadIdx.x;
  float loss = 0.f, cnt = 0.f;
  if (pix < npix) {
    long long b = pix / HW;
    int p = (int)(pix - b * HW);
    const float* __restrict__ x = logits + (size_t)b * C * HW + p;
    long long lab = labels[pix];
    // online softmax in the log2 domain, SEMF_U channels (independent loads) per step
    float m = -INFINITY, s = 0.f;
    for (int c0 = 0; c0 < C; c0 += SEMF_U) {
      float v[SEMF_U];
#pragma unroll
      for (int i = 0; i < SEMF_U; ++i) v[i] = (c0 + i < C) ? __ldg(x + (size_t)(c0 + i) * HW) * SEM_LOG2E : -INFINITY;
      float mm = v[0];
#pragma unroll
      for (int i = 1; i < SEMF_U; ++i) mm = fmaxf(mm, v[i]);
      float mn = fmaxf(m, mm);
      s *= sem_ex2(m - mn);  // first step: 0 * 2^-inf = 0
#pragma unroll
      for (int i = 0; i < SEMF_U; ++i) s += sem_ex2(v[i] - mn);
      m = mn;
    }
    float l2 = m + log2f(s);
    lse2[pix] = l2;  // log2-domain log-sum-exp, consumed by the backward
    if (lab != (long long)ignore && lab >= 0 && lab < C) {
      loss = (l2 - __ldg(x + (size_t)lab * HW) * SEM_LOG2E) * SEM_LN2;
      cnt = 1.f;
    }
  }
  __shared__ double sh[32];
  double r0 = block_sum_d((double)loss, sh);
  double r1 = block_sum_d((double)cnt, sh);
  if (threadIdx.x == 0) {
    partials[2 * (size_t)blockIdx.x] = r0;
    partials[2 * (size_t)blockIdx.x + 1] = r1;
  }
}

// d logits[c] = gout / count * (softmax_c - [c == label]) on counted pixels, 0 elsewhere
__global__ void __launch_bounds__(SEMF_THREADS)
sem_ce_full_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                       const float* __restrict__ lse2, int C, int HW, long long npix, int ignore,
                       const float* __restrict__ out3, const float* __restrict__ gout, float* __restrict__ dlogits) {
  long long pix = (long long)blockIdx.x * SEMF_THREADS + threadIdx.x;
  if (pix >= npix) return;
  long long b = pix / HW;
  int p = (int)(pix - b * HW);
  const float* __restrict__ x = logits + (size_t)b * C * HW + p;
  float* __restrict__ o = dlogits + (size_t)b * C * HW + p;
  long long lab = labels[pix];
  bool valid = lab != (long long)ignore && lab >= 0 && lab < C;
  float cnt = __ldg(out3 + 2);
  float scale = cnt > 0.f ? __ldg(gout) / cnt : 0.f;
  if (!valid) {
    for (int c = 0; c < C; ++c) o[(size_t)c * HW] = 0.f;
    return;
  }
  float l2 = lse2[pix];
  int il = (int)lab;
  for (int c0 = 0; c0 < C; c0 += SEMF_U) {
    float v[SEMF_U];
#pragma unroll
    for (int i = 0; i < SEMF_U; ++i) v[i] = (c0 + i < C) ? __ldg(x + (size_t)(c0 + i) * HW) : 0.f;
#pragma unroll
    for (int i = 0; i < SEMF_U; ++i)
      if (c0 + i < C) {
        float pr = sem_ex2(fmaf(v[i], SEM_LOG2E, -l2));
        o[(size_t)(c0 + i) * HW] = (pr - (c0 + i == il ? 1.f : 0.f)) * scale;
      }
  }
}

// ----------------------------------------------------------------------------------------------
// fused x8 bilinear upsample + cross entropy
//
// F.interpolate(align_corners=False) with an exact x8 factor: source = (dst + 0.5) / 8 - 0.5, clamped below at 0, right /
// bottom neighbour clamped to the last cell.  Hence the 8x8 pixel block whose top-left pixel is (8i-4, 8j-4) -- a
// "phase block" (i, j), i in [0,Hc], j in [0,Wc] -- reads exactly the four cells (i-1|i, j-1|j) (clamped at the
// borders, where the block is cut by the image edge) with the SAME 8 horizontal and 8 vertical weights
// lambda = (2k+1)/16, k = 0..7.  A thread block walks a strip of phase blocks along x:
//   phase A  thread = channel c: e[p][c] = 2^(logit_p[c] - shift_p) for the 64 pixels (interpolation = 2 FMAs per row
//            + 1 per pixel, weights are immediates); shift_p = the same interpolation of the per-cell channel maxima,
//            an upper bound of every logit of the pixel, so no per-pixel max pass is needed
//   phase B  thread = pixel p: S_p = sum_c e[p][c] (column sums of the transposed tile in shared memory), the loss
//            term log S_p - (logit_p[label] - shift_p), q_p = 1/S_p (0 on ignored pixels), e[p][label] -= S_p
//   phase C  thread = channel c: grad_c(cell) += w_cell(p) * e[p][c] * q_p, separable again: row sums, then 4 FMAs per
//            row into rolling per-cell accumulators that are flushed with one float RED per finished cell
// If the shift bound is so loose that S_p underflows (logit spread > ~87 across the four cells) the pixel is redone
// against its own maximum (sem_pixel_exact).
// ----------------------------------------------------------------------------------------------
#define SEMU_PBMAX 11              // phase blocks per strip
#define SEMU_NCOL (SEMU_PBMAX + 1)  // cell columns per strip
#define SEMU_EP 65                  // E row stride (floats): bank = (c + p) mod 32, conflict-free both ways

struct SemUpParams {
  const float* logits;      // [B,C,Hc,Wc]
  const long long* labels;  // [B,8Hc,8Wc]
  float* gsum;              // [B,C,Hc,Wc] un-normalised gradient (zeroed by the host entry), or nullptr
  double* partials;         // [grid][2]
  int B, C, Hc, Wc, ignore, pbx, nstrips, cs;
};

__device__ __forceinline__ float sem_interp(float tl, float tr, float bl, float br, float ly, float lx) {
  float hl = fmaf(ly, bl - tl, tl), hr = fmaf(ly, br - tr, tr);
  return fmaf(lx, hr - hl, hl);
}

// rare path: recompute e[.][p] and S for pixel p against the pixel's own maximum; returns S, sets v_lab
__device__ __noinline__ float sem_pixel_exact(const float* __restrict__ L0, const float* __restrict__ L1, int cs, int C,
                                              float ly, float lx, int lab, int p, float* __restrict__ E, float& v_lab) {
  float vmax = -INFINITY;
  for (int c = 0; c < C; ++c) vmax = fmaxf(vmax, sem_interp(L0[c], L0[cs + c], L1[c], L1[cs + c], ly, lx));
  float S = 0.f;
  for (int c = 0; c < C; ++c) {
    float e = sem_ex2(sem_interp(L0[c], L0[cs + c], L1[c], L1[cs + c], ly, lx) - vmax);
    E[c * SEMU_EP + p] = e;
    S += e;
  }
  v_lab = sem_interp(L0[lab], L0[cs + lab], L1[lab], L1[cs + lab], ly, lx) - vmax;
  return S;
}

template <bool GRAD>
__global__ void __launch_bounds__(256)
sem_ce_up8_kernel(const __grid_constant__ SemUpParams P) {
  extern __shared__ __align__(16) float smem[];
  const int C = P.C, cs = P.cs, Hc = P.Hc, Wc = P.Wc;
  float* Q = smem;                            // [64]                1/S_p, 0 on ignored pixels (16-byte aligned)
  float* red = Q + 64;                        // [8]
  float* Mx = red + 8;                        // [2][SEMU_NCOL]      per-cell channel maximum
  float* Pm = Mx + 2 * SEMU_NCOL;             // [16][2*SEMU_NCOL]   partial maxima
  float* E = Pm + 16 * 2 * SEMU_NCOL;         // [C][65]             2^(logit - shift), transposed tile
  float* Ls = E + C * SEMU_EP;                // [2][SEMU_NCOL][cs]  cell logits * log2(e)
  const int tid = threadIdx.x, nthr = blockDim.x;

  int blk = blockIdx.x;
  const int strip = blk % P.nstrips;
  blk /= P.nstrips;
  const int ib = blk % (Hc + 1);
  const int b = blk / (Hc + 1);
  const int jb0 = strip * P.pbx;
  const int npb = min(P.pbx, Wc + 1 - jb0);
  const int ncolv = npb + 1;
  const int itop = max(ib - 1, 0), ibot = min(ib, Hc - 1);
  const int H = Hc * 8, W = Wc * 8;

  // ---- stage the 2 x (npb+1) cells of the strip, all channels, scaled to the log2 domain
  {
    const float* __restrict__ Lg = P.logits + (size_t)b * C * Hc * Wc;
    const int per_c = 2 * ncolv;
    for (int idx = tid; idx < per_c * C; idx += nthr) {
      int c = idx / per_c, rem = idx - c * per_c;
      int row = rem / ncolv, k = rem - row * ncolv;
      int j = min(max(jb0 - 1 + k, 0), Wc - 1);
      int i = row ? ibot : itop;
      Ls[(row * SEMU_NCOL + k) * cs + c] = __ldg(Lg + ((size_t)c * Hc + i) * Wc + j) * SEM_LOG2E;
    }
  }
  __syncthreads();
  // ---- per-cell maximum over channels
  {
    const int ncell = 2 * ncolv;
    int nsub = min(nthr / ncell, 16);
    int cell = tid % ncell, sub = tid / ncell;
    if (sub < nsub) {
      int row = cell / ncolv, k = cell - row * ncolv;
      const float* src = Ls + (row * SEMU_NCOL + k) * cs;
      int chunk = (C + nsub - 1) / nsub;
      float m = -INFINITY;
      for (int c = sub * chunk; c < min(C, (sub + 1) * chunk); ++c) m = fmaxf(m, src[c]);
      Pm[sub * 2 * SEMU_NCOL + row * SEMU_NCOL + k] = m;
    }
    __syncthreads();
    if (tid < ncell) {
      int row = tid / ncolv, k = tid - row * ncolv;
      float m = -INFINITY;
      for (int s = 0; s < nsub; ++s) m = fmaxf(m, Pm[s * 2 * SEMU_NCOL + row * SEMU_NCOL + k]);
      Mx[row * SEMU_NCOL + k] = m;
    }
  }
  __syncthreads();

  const bool chan = tid < C;
  const bool pixt = tid < 64;
  float a_tl = 0.f, a_bl = 0.f;
  if (chan) {
    a_tl = Ls[tid];
    a_bl = Ls[SEMU_NCOL * cs + tid];
  }
  float acc_tl = 0.f, acc_bl = 0.f, acc_tr = 0.f, acc_br = 0.f;
  float loss_acc = 0.f, cnt_acc = 0.f;
  const int pr = tid >> 3, pdx = tid & 7;  // pixel thread: row / column inside the phase block
  const int py = 8 * ib - 4 + pr;
  float* __restrict__ Ec = E + tid * SEMU_EP;

  for (int pb = 0; pb < npb; ++pb) {
    // label of this thread's pixel, in flight during phase A
    long long lab = P.ignore;
    if (pixt) {
      int px = 8 * (jb0 + pb) - 4 + pdx;
      if (py >= 0 && py < H && px >= 0 && px < W) lab = P.labels[((size_t)b * H + py) * W + px];
    }
    const float m_tl = Mx[pb], m_tr = Mx[pb + 1], m_bl = Mx[SEMU_NCOL + pb], m_br = Mx[SEMU_NCOL + pb + 1];
    float a_tr = 0.f, a_br = 0.f;
    // ---- phase A
    if (chan) {
      a_tr = Ls[(pb + 1) * cs + tid];
      a_br = Ls[(SEMU_NCOL + pb + 1) * cs + tid];
      const float dl = a_bl - a_tl, dr = a_br - a_tr, dml = m_bl - m_tl, dmr = m_br - m_tr;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float ly = (2 * r + 1) * 0.0625f;
        float hl = fmaf(ly, dl, a_tl) - fmaf(ly, dml, m_tl);
        float hr = fmaf(ly, dr, a_tr) - fmaf(ly, dmr, m_tr);
        float dh = hr - hl;
#pragma unroll
        for (int dx = 0; dx < 8; ++dx) Ec[r * 8 + dx] = sem_ex2(fmaf((2 * dx + 1) * 0.0625f, dh, hl));
      }
    }
    __syncthreads();
    // ---- phase B
    if (pixt) {
      float S0 = 0.f, S1 = 0.f, S2 = 0.f, S3 = 0.f;
      const float* __restrict__ Ep = E + tid;
      int c = 0;
      for (; c + 4 <= C; c += 4) {
        S0 += Ep[(c + 0) * SEMU_EP];
        S1 += Ep[(c + 1) * SEMU_EP];
        S2 += Ep[(c + 2) * SEMU_EP];
        S3 += Ep[(c + 3) * SEMU_EP];
      }
      for (; c < C; ++c) S0 += Ep[c * SEMU_EP];
      float S = (S0 + S1) + (S2 + S3);
      float q = 0.f;
      if (lab != (long long)P.ignore && lab >= 0 && lab < C) {
        const int il = (int)lab;
        const float ly = (2 * pr + 1) * 0.0625f, lx = (2 * pdx + 1) * 0.0625f;
        const float* __restrict__ L0 = Ls + pb * cs;
        const float* __restrict__ L1 = Ls + (SEMU_NCOL + pb) * cs;
        float v;
        if (S > 1e-30f) {
          v = sem_interp(L0[il], L0[cs + il], L1[il], L1[cs + il], ly, lx) - sem_interp(m_tl, m_tr, m_bl, m_br, ly, lx);
        } else {
          S = sem_pixel_exact(L0, L1, cs, C, ly, lx, il, tid, E, v);
        }
        loss_acc += (log2f(S) - v) * SEM_LN2;
        cnt_acc += 1.f;
        q = 1.f / S;
        if (GRAD) E[il * SEMU_EP + tid] -= S;  // (e - S) q = softmax - 1 for the label channel
      }
      Q[tid] = q;
    }
    __syncthreads();
    // ---- phase C
    if (GRAD && chan) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 q0 = *reinterpret_cast<const float4*>(Q + r * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(Q + r * 8 + 4);
        const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        float rs = 0.f, rr = 0.f;
#pragma unroll
        for (int dx = 0; dx < 8; ++dx) {
          float t = Ec[r * 8 + dx] * qq[dx];
          rs += t;
          rr = fmaf((2 * dx + 1) * 0.0625f, t, rr);
        }
        const float rl = rs - rr;
        const float ly = (2 * r + 1) * 0.0625f;
        acc_tl = fmaf(1.f - ly, rl, acc_tl);
        acc_bl = fmaf(ly, rl, acc_bl);
        acc_tr = fmaf(1.f - ly, rr, acc_tr);
        acc_br = fmaf(ly, rr, acc_br);
      }
      // the left cell column is complete for this strip row: one RED per cell, then roll
      float* __restrict__ G = P.gsum + ((size_t)b * C + tid) * Hc * Wc;
      const int jl = min(max(jb0 + pb - 1, 0), Wc - 1);
      atomicAdd(G + itop * Wc + jl, acc_tl);
      atomicAdd(G + ibot * Wc + jl, acc_bl);
      acc_tl = acc_tr;
      acc_bl = acc_br;
      acc_tr = 0.f;
      acc_br = 0.f;
    }
    a_tl = a_tr;
    a_bl = a_br;
    __syncthreads();  // E and Q are rewritten by the next phase block
  }
  if (GRAD && chan) {
    float* __restrict__ G = P.gsum + ((size_t)b * C + tid) * Hc * Wc;
    const int jl = min(max(jb0 + npb - 1, 0), Wc - 1);
    atomicAdd(G + itop * Wc + jl, acc_tl);
    atomicAdd(G + ibot * Wc + jl, acc_bl);
  }
  // ---- block partials: pixel threads are warps 0 and 1
  if (tid < 64) {
    float s = warp_sum(loss_acc), n = warp_sum(cnt_acc);
    if ((tid & 31) == 0) {
      red[(tid >> 5) * 2] = s;
      red[(tid >> 5) * 2 + 1] = n;
    }
  }
  __syncthreads();
  if (tid == 0) {
    P.partials[2 * (size_t)blockIdx.x] = (double)red[0] + (double)red[2];
    P.partials[2 * (size_t)blockIdx.x + 1] = (double)red[1] + (double)red[3];
  }
}

__global__ void sem_scale_kernel(const float* __restrict__ gsum, const float* __restrict__ out3,
                                 const float* __restrict__ gout, size_t n, float* __restrict__ dlogits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float cnt = __ldg(out3 + 2);
  float scale = cnt > 0.f ? __ldg(gout) / cnt : 0.f;
  if (i < n) dlogits[i] = gsum[i] * scale;
}

// ----------------------------------------------------------------------------------------------
// host entry points
// ----------------------------------------------------------------------------------------------
static int sem_up_plan(int C, int Wc, int& pbx, int& nstrips, int& cs, int& nthr, size_t& smem) {
  nstrips = ssp_ceil_div(Wc + 1, SEMU_PBMAX);
  pbx = ssp_ceil_div(Wc + 1, nstrips);
  cs = C | 1;
  nthr = ((C > 64 ? C : 64) + 31) / 32 * 32;
  smem = (64 + 8 + 2 * SEMU_NCOL + 16 * 2 * SEMU_NCOL + (size_t)C * SEMU_EP + 2 * SEMU_NCOL * cs) * sizeof(float);
  return 0;
}

extern "C" size_t ssp_sem_ce_ws_bytes(int B, int C, int H, int W, int upsampled) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  size_t nblk;
  if (upsampled) {
    int pbx, nstrips, cs, nthr;
    size_t smem;
    sem_up_plan(C, W / 8, pbx, nstrips, cs, nthr, smem);
    nblk = (size_t)B * (H / 8 + 1) * nstrips;
  } else {
    nblk = ((size_t)B * H * W + SEMF_THREADS - 1) / SEMF_THREADS;
  }
  return nblk * 2 * sizeof(double);
}

extern "C" int ssp_sem_ce_fwd(const float* logits, const long long* labels, int B, int C, int H, int W,
                              int ignore_index, float* lse2, float* out3, void* ws, size_t ws_bytes, void* stream) {
  SSP_REQUIRE(logits && labels && lse2 && out3 && ws, "ssp_sem_ce_fwd: null pointer");
  SSP_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), "ssp_sem_ce_fwd: bad sizes B=%d C=%d H=%d W=%d", B, C, H, W);
  SSP_REQUIRE(ws_bytes >= ssp_sem_ce_ws_bytes(B, C, H, W, 0) && ((uintptr_t)ws & 7) == 0, "ssp_sem_ce_fwd: workspace too small or misaligned");
  long long npix = (long long)B * H * W;
  int nblk = (int)((npix + SEMF_THREADS - 1) / SEMF_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  sem_ce_full_fwd_kernel<<<nblk, SEMF_THREADS, 0, st>>>(logits, labels, C, H * W, npix, ignore_index, lse2, (double*)ws);
  SSP_CUDA_CHECK_LAUNCH("sem_ce_full_fwd_kernel");
  sem_finalize_kernel<<<1, 256, 0, st>>>((const double*)ws, nblk, out3);
  SSP_CUDA_CHECK_LAUNCH("sem_finalize_kernel");
  return SSP_OK;
}

extern "C" int ssp_sem_ce_bwd(const float* logits, const long long* labels, const float* lse2, int B, int C, int H,
                              int W, int ignore_index, const float* out3, const float* gout, float* dlogits,
                              void* stream) {
  SSP_REQUIRE(logits && labels && lse2 && out3 && gout && dlogits, "ssp_sem_ce_bwd: null pointer");
  SSP_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), "ssp_sem_ce_bwd: bad sizes B=%d C=%d H=%d W=%d", B, C, H, W);
  long long npix = (long long)B * H * W;
  int nblk = (int)((npix + SEMF_THREADS - 1) / SEMF_THREADS);
  sem_ce_full_bwd_kernel<<<nblk, SEMF_THREADS, 0, (cudaStream_t)stream>>>(logits, labels, lse2, C, H * W, npix,
                                                                         ignore_index, out3, gout, dlogits);
  SSP_CUDA_CHECK_LAUNCH("sem_ce_full_bwd_kernel");
  return SSP_OK;
}

extern "C" int ssp_sem_ce_up8(const float* logits_lr, const long long* labels, int B, int C, int Hc, int Wc,
                              int ignore_index, float* gsum, float* out3, void* ws, size_t ws_bytes, void* stream) {
  SSP_REQUIRE(logits_lr && labels && out3 && ws, "ssp_sem_ce_up8: null pointer");
  SSP_REQUIRE(B > 0 && Hc > 0 && Wc > 0 && C >= 2 && C <= 256, "ssp_sem_ce_up8: bad sizes B=%d C=%d (2..256) Hc=%d Wc=%d", B, C, Hc, Wc);
  SSP_REQUIRE(ws_bytes >= ssp_sem_ce_ws_bytes(B, C, Hc * 8, Wc * 8, 1) && ((uintptr_t)ws & 7) == 0, "ssp_sem_ce_up8: workspace too small or misaligned");
  SemUpParams P;
  int nthr;
  size_t smem;
  sem_up_plan(C, Wc, P.pbx, P.nstrips, P.cs, nthr, smem);
  long long nblk = (long long)B * (Hc + 1) * P.nstrips;
  SSP_REQUIRE(nblk < (1ll << 31), "ssp_sem_ce_up8: grid too large");
  P.logits = logits_lr; P.labels = labels; P.gsum = gsum; P.partials = (double*)ws;
  P.B = B; P.C = C; P.Hc = Hc; P.Wc = Wc; P.ignore = ignore_index;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done[2] = {false, false};
  if (gsum) {
    SSP_CUDA_CALL(cudaMemsetAsync(gsum, 0, (size_t)B * C * Hc * Wc * sizeof(float), st));
    if (!attr_done[1]) {
      SSP_CUDA_CALL(cudaFuncSetAttribute(sem_ce_up8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done[1] = true;
    }
    sem_ce_up8_kernel<true><<<(unsigned)nblk, nthr, smem, st>>>(P);
  } else {
    if (!attr_done[0]) {
      SSP_CUDA_CALL(cudaFuncSetAttribute(sem_ce_up8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done[0] = true;
    }
    sem_ce_up8_kernel<false><<<(unsigned)nblk, nthr, smem, st>>>(P);
  }
  SSP_CUDA_CHECK_LAUNCH("sem_ce_up8_kernel");
  sem_finalize_kernel<<<1, 256, 0, st>>>((const double*)ws, (int)nblk, out3);
  SSP_CUDA_CHECK_LAUNCH("sem_finalize_kernel");
  return SSP_OK;
}

extern "C" int ssp_sem_ce_up8_bwd(const float* gsum, const float* out3, const float* gout, int B, int C, int Hc,
                                  int Wc, float* dlogits, void* stream) {
  SSP_REQUIRE(gsum && out3 && gout && dlogits, "ssp_sem_ce_up8_bwd: null pointer");
  SSP_REQUIRE(B > 0 && C > 0 && Hc > 0 && Wc > 0, "ssp_sem_ce_up8_bwd: bad sizes");
  size_t n = (size_t)B * C * Hc * Wc;
  sem_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gsum, out3, gout, n, dlogits);
  SSP_CUDA_CHECK_LAUNCH("sem_scale_kernel");
  return SSP_OK;
}
