// Sparse descriptor loss: the loss every shipped training config uses.
// Reference: utils/loss_functions/sparse_loss.py:65-284 (descriptor_loss_sparse, batch_descriptor_loss_sparse) with
//            utils/loss_functions/pixelwise_contrastive_loss.py:140-265 (match_loss / non_match_descriptor_loss, dist="cos").
//
// Per image the reference samples K = num_matching_attempts cell correspondences (a, b) through the homography and
// Kn = K * num_masked_non_matches_per_match random non-matches on the HOST (numpy / torch CPU RNG) -- the sampled index lists
// are inputs here (semantic-superpoint_b200/sparse.py reproduces the sampling call for call) -- and then evaluates
//   match_b    = 1/K  sum_k max(1 - <D[:,a_k], Dw[:,b_k]>, 0)
//   nonmatch_b = sum_k max(<D[:,a'_k], Dw[:,b'_k]> - 0.2, 0) / (#{k: term > 0} + 1)
//   loss_b     = lamda_d * match_b + nonmatch_b;      outputs = means over the batch of (loss, match, nonmatch).
// Kernels: NCHW -> cell-major transpose (a gathered descriptor becomes one contiguous 1 KB row instead of 256 words 4*Nc bytes
// apart), one warp per sampled pair for the dot products, a fixed-order per-image reduction, and for the backward a
// warp-per-pair scatter (fp32 atomics into cell-major gradient buffers) followed by the transpose back.
#include "common.cuh"

// [B, Dch, Nc] <-> [B, Nc, Dch] through a 32 x 32 shared tile (both directions coalesced)
__global__ void __launch_bounds__(256)
sparse_transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  src += (size_t)b * rows * cols;
  dst += (size_t)b * rows * cols;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < rows && c < cols) ? __ldg(src + (size_t)r * cols + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[tx][ty + 8 * k];
  }
}

extern "C" int ssp_transpose_batched(const float* src, int B, int rows, int cols, float* dst, void* stream) {
  SSP_REQUIRE(src && dst, "ssp_transpose_batched: null pointer");
  SSP_REQUIRE(B > 0 && B <= 65535 && rows > 0 && cols > 0, "ssp_transpose_batched: bad sizes");
  dim3 grid(ssp_ceil_div(cols, 32), ssp_ceil_div(rows, 32), B);
  SSP_REQUIRE(grid.y <= 65535, "ssp_transpose_batched: too many rows");
  sparse_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, dst);
  SSP_CUDA_CHECK_LAUNCH("sparse_transpose_kernel");
  return SSP_OK;
}

// dots[b, k] = <Dt[b, ia[b,k], :], Dwt[b, ib[b,k], :]>   (cell-major descriptors [B, Nc, Dch]); one warp per pair
__global__ void __launch_bounds__(256)
sparse_dots_kernel(const float* __restrict__ Dt, const float* __restrict__ Dwt, const int* __restrict__ ia,
                   const int* __restrict__ ib, int Kt, int Nc, int Dch, float* __restrict__ dots) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= Kt) return;
  const int a = ia[(size_t)b * Kt + k], c = ib[(size_t)b * Kt + k];
  float acc = 0.f;
  if (a >= 0 && a < Nc && c >= 0 && c < Nc) {
    const float* pa = Dt + ((size_t)b * Nc + a) * Dch;
    const float* pc = Dwt + ((size_t)b * Nc + c) * Dch;
    for (int d = lane; d < Dch; d += 32) acc = fmaf(__ldg(pa + d), __ldg(pc + d), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) dots[(size_t)b * Kt + k] = acc;
}

// per image, fixed summation order: stats[b] = { match_b, nonmatch_b, loss_b, hard count }
__global__ void __launch_bounds__(256)
sparse_reduce_kernel(const float* __restrict__ dots, int K, int Kn, float lamda, float mpos, float mneg,
                     float* __restrict__ stats) {
  __shared__ double shd[32];
  const int b = blockIdx.x;
  const float* dm = dots + (size_t)b * (K + Kn);
  const float* dn = dm + K;
  double sm = 0.0, sn = 0.0, hc = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sm += (double)fmaxf(mpos - dm[k], 0.f);
  for (int k = threadIdx.x; k < Kn; k += blockDim.x) {
    const float t = fmaxf(dn[k] - mneg, 0.f);
    sn += (double)t;
    hc += t != 0.f ? 1.0 : 0.0;  // torch.nonzero(non_match_loss): hard negatives
  }
  sm = block_sum_d(sm, shd);
  sn = block_sum_d(sn, shd);
  hc = block_sum_d(hc, shd);
  if (threadIdx.x == 0) {
    const float match = (float)(sm / (double)K);                    // 1/num_matches * sum
    const float nonmatch = (float)sn / ((float)hc + 1.f);           // sum / (num_hard_negatives + 1)
    stats[4 * b] = match;
    stats[4 * b + 1] = nonmatch;
    stats[4 * b + 2] = lamda * match + nonmatch;
    stats[4 * b + 3] = (float)hc;
  }
}

// batch means: out3 = { mean loss, mean match, mean nonmatch }   (torch.stack(...).mean(), sparse_loss.py:283-284)
__global__ void sparse_mean_kernel(const float* __restrict__ stats, int B, float* __restrict__ out3) {
  if (threadIdx.x == 0) {
    float l = 0.f, m = 0.f, n = 0.f;
    for (int b = 0; b < B; ++b) { m += stats[4 * b]; n += stats[4 * b + 1]; l += stats[4 * b + 2]; }
    out3[0] = l / (float)B;
    out3[1] = m / (float)B;
    out3[2] = n / (float)B;
  }
}

// Dt / Dwt: cell-major descriptors (ssp_transpose_batched of the NCHW tensors); ia / ib [B, K + Kn]: the K matches followed by the
// Kn non-matches (cell indices a in `descriptors`, b in `descriptors_warped`); dots [B, K + Kn] and stats [B, 4] are kept for
// the backward; out3 = the three returned scalars.
extern "C" int ssp_sparse_desc_loss_fwd(const float* Dt, const float* Dwt, const int* ia, const int* ib, int B, int Nc, int Dch,
                                        int K, int Kn, float lamda, float mpos, float mneg, float* dots, float* stats,
                                        float* out3, void* stream) {
  SSP_REQUIRE(Dt && Dwt && ia && ib && dots && stats && out3, "ssp_sparse_desc_loss_fwd: null pointer");
  SSP_REQUIRE(B > 0 && B <= 65535 && Nc > 0 && Dch > 0 && K > 0 && Kn >= 0, "ssp_sparse_desc_loss_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int Kt = K + Kn;
  dim3 grid(ssp_ceil_div(Kt, 8), B);
  sparse_dots_kernel<<<grid, 256, 0, st>>>(Dt, Dwt, ia, ib, Kt, Nc, Dch, dots);
  SSP_CUDA_CHECK_LAUNCH("sparse_dots_kernel");
  sparse_reduce_kernel<<<B, 256, 0, st>>>(dots, K, Kn, lamda, mpos, mneg, stats);
  SSP_CUDA_CHECK_LAUNCH("sparse_reduce_kernel");
  sparse_mean_kernel<<<1, 32, 0, st>>>(stats, B, out3);
  SSP_CUDA_CHECK_LAUNCH("sparse_mean_kernel");
  return SSP_OK;
}

// backward: coefficient of every sampled pair, scattered into the cell-major gradients
//   match k:     -1/K * [mpos - dot >= 0] * (g_loss * lamda + g_match) / B        (torch.clamp passes the gradient at the bound)
//   non-match k:  1/(hard_b + 1) * [dot - mneg >= 0] * (g_loss + g_nonmatch) / B
//   dDt[b, a, :] += coef * Dwt[b, c, :],   dDwt[b, c, :] += coef * Dt[b, a, :]
__global__ void __launch_bounds__(256)
sparse_bwd_kernel(const float* __restrict__ Dt, const float* __restrict__ Dwt, const int* __restrict__ ia,
                  const int* __restrict__ ib, const float* __restrict__ dots, const float* __restrict__ stats,
                  const float* __restrict__ g3, int B, int Nc, int Dch, int K, int Kn, float lamda, float mpos, float mneg,
                  float* __restrict__ dDt, float* __restrict__ dDwt) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int Kt = K + Kn;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= Kt) return;
  const float dot = dots[(size_t)b * Kt + k];
  float coef;
  if (k < K) coef = (mpos - dot >= 0.f) ? -(g3[0] * lamda + g3[1]) / ((float)K * (float)B) : 0.f;
  else       coef = (dot - mneg >= 0.f) ? (g3[0] + g3[2]) / ((stats[4 * b + 3] + 1.f) * (float)B) : 0.f;
  if (coef == 0.f) return;
  const int a = ia[(size_t)b * Kt + k], c = ib[(size_t)b * Kt + k];
  if (a < 0 || a >= Nc || c < 0 || c >= Nc) return;
  const float* pa = Dt + ((size_t)b * Nc + a) * Dch;
  const float* pc = Dwt + ((size_t)b * Nc + c) * Dch;
  float* ga = dDt + ((size_t)b * Nc + a) * Dch;
  float* gc = dDwt + ((size_t)b * Nc + c) * Dch;
  for (int d = lane; d < Dch; d += 32) {
    atomicAdd(ga + d, coef * __ldg(pc + d));
    atomicAdd(gc + d, coef * __ldg(pa + d));
  }
}

// dDt / dDwt [B, Nc, Dch] are zeroed here; g3 = device { dL/dloss, dL/dmatch, dL/dnonmatch }
extern "C" int ssp_sparse_desc_loss_bwd(const float* Dt, const float* Dwt, const int* ia, const int* ib, const float* dots,
                                        const float* stats, const float* g3, int B, int Nc, int Dch, int K, int Kn,
                                        float lamda, float mpos, float mneg, float* dDt, float* dDwt, void* stream) {
  SSP_REQUIRE(Dt && Dwt && ia && ib && dots && stats && g3 && dDt && dDwt, "ssp_sparse_desc_loss_bwd: null pointer");
  SSP_REQUIRE(B > 0 && B <= 65535 && Nc > 0 && Dch > 0 && K > 0 && Kn >= 0, "ssp_sparse_desc_loss_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * Nc * Dch * sizeof(float);
  SSP_CUDA_CALL(cudaMemsetAsync(dDt, 0, n, st));
  SSP_CUDA_CALL(cudaMemsetAsync(dDwt, 0, n, st));
  dim3 grid(ssp_ceil_div(K + Kn, 8), B);
  sparse_bwd_kernel<<<grid, 256, 0, st>>>(Dt, Dwt, ia, ib, dots, stats, g3, B, Nc, Dch, K, Kn, lamda, mpos, mneg, dDt, dDwt);
  SSP_CUDA_CHECK_LAUNCH("sparse_bwd_kernel");
  return SSP_OK;
}
