// Homography-adaptation heatmap aggregation.
// Reference: export.py:49-60 (combine_heatmap)  -- Gabriel-SGama/Semantic-SuperPoint
//   out = sum_n warp_n(heat_n * mask_n) / sum_n warp_n(mask_n), both warps bilinear with the same H_n.
// One kernel: the heat*mask product is formed at the gather, the N-sum stays in registers, the two
// intermediate [N,1,H,W] warped stacks of the reference never exist.  HBM traffic = the algorithmic
// (2N+1)*H*W*4 bytes.  A block owns an 8x8 output tile x 4 interleaved view groups.
#include "common.cuh"

#define CH_PIX 64
#define CH_GROUPS 4

__global__ void __launch_bounds__(CH_PIX * CH_GROUPS)
combine_heatmap_kernel(const float* __restrict__ heat, const float* __restrict__ mask,
                       const float* __restrict__ Hinv, int N, int H, int W, const float* __restrict__ xs,
                       const float* __restrict__ ys, float* __restrict__ out) {
  extern __shared__ float sh[];  // N*9 homographies, then 2*CH_PIX*CH_GROUPS partial sums
  float* hs = sh;
  float* part = sh + N * 9;
  // blockIdx.y = source image (batched export: heat/mask [I,N,H,W], Hinv [I,N,3,3], out [I,H,W])
  heat += (size_t)blockIdx.y * N * H * W;
  mask += (size_t)blockIdx.y * N * H * W;
  Hinv += (size_t)blockIdx.y * N * 9;
  out += (size_t)blockIdx.y * H * W;
  for (int i = threadIdx.x; i < N * 9; i += blockDim.x) hs[i] = Hinv[i];
  __syncthreads();
  int lp = threadIdx.x % CH_PIX, g = threadIdx.x / CH_PIX;
  // a block owns an 8x8 pixel tile, a warp an 8x4 patch: under any rotation the bilinear footprints of a warp
  // stay inside a compact source patch (few 32 B sectors per gather instead of one per lane)
  int tiles_x = (W + 7) / 8;
  int x = (blockIdx.x % tiles_x) * 8 + (lp & 7), y = (blockIdx.x / tiles_x) * 8 + (lp >> 3);
  bool inside = x < W && y < H;
  int pix = y * W + x;
  float sum_h = 0.f, sum_m = 0.f;
  if (inside) {
    float gx = __ldg(xs + x), gy = __ldg(ys + y);
    size_t plane = (size_t)H * W;
    for (int n = g; n < N; n += CH_GROUPS) {
      const float* h = hs + n * 9;
      float nx, ny;
      homography_apply(h, gx, gy, nx, ny);
      float ix = ((nx + 1.f) / 2.f) * (float)(W - 1);
      float iy = ((ny + 1.f) / 2.f) * (float)(H - 1);
      float fx = floorf(ix), fy = floorf(iy);
      // reject views whose 2x2 footprint is entirely outside (also keeps the int casts in range)
      if (!(fx >= -1.f && fx < (float)W && fy >= -1.f && fy < (float)H)) continue;
      int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
      float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
      float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
      bool xin0 = x0 >= 0, xin1 = x1 < W, yin0 = y0 >= 0, yin1 = y1 < H;
      const float* hp = heat + n * plane;
      const float* mp = mask + n * plane;
      float ah = 0.f, am = 0.f;
      if (yin0) {
        size_t r = (size_t)y0 * W;
        if (xin0) { float m = __ldg(mp + r + x0), w = wx0 * wy0; am += m * w; ah += (__ldg(hp + r + x0) * m) * w; }
        if (xin1) { float m = __ldg(mp + r + x1), w = wx1 * wy0; am += m * w; ah += (__ldg(hp + r + x1) * m) * w; }
      }
      if (yin1) {
        size_t r = (size_t)y1 * W;
        if (xin0) { float m = __ldg(mp + r + x0), w = wx0 * wy1; am += m * w; ah += (__ldg(hp + r + x0) * m) * w; }
        if (xin1) { float m = __ldg(mp + r + x1), w = wx1 * wy1; am += m * w; ah += (__ldg(hp + r + x1) * m) * w; }
      }
      sum_h += ah;
      sum_m += am;
    }
  }
  part[threadIdx.x] = sum_h;
  part[CH_PIX * CH_GROUPS + threadIdx.x] = sum_m;
  __syncthreads();
  if (g == 0 && inside) {
    float th = 0.f, tm = 0.f;
#pragma unroll
    for (int q = 0; q < CH_GROUPS; ++q) {
      th += part[q * CH_PIX + lp];
      tm += part[CH_PIX * CH_GROUPS + q * CH_PIX + lp];
    }
    out[pix] = th / tm;  // 0/0 -> NaN exactly like the reference when no view covers the pixel
  }
}

extern "C" int ssp_combine_heatmap(const float* heat, const float* mask, const float* Hinv, int I, int N, int H, int W,
                                   const float* xs, const float* ys, float* out, void* stream) {
  SSP_REQUIRE(heat && mask && Hinv && xs && ys && out, "ssp_combine_heatmap: null pointer");
  SSP_REQUIRE(I > 0 && I <= 65535 && N > 0 && H > 0 && W > 0, "ssp_combine_heatmap: bad sizes I=%d N=%d H=%d W=%d", I, N, H, W);
  size_t smem = ((size_t)N * 9 + 2 * CH_PIX * CH_GROUPS) * sizeof(float);
  SSP_REQUIRE(smem <= 48 * 1024, "ssp_combine_heatmap: N=%d views exceed the shared-memory table (max ~1100)", N);
  dim3 nblk(ssp_ceil_div(W, 8) * ssp_ceil_div(H, 8), I);
  combine_heatmap_kernel<<<nblk, CH_PIX * CH_GROUPS, smem, (cudaStream_t)stream>>>(heat, mask, Hinv, N, H, W, xs, ys, out);
  SSP_CUDA_CHECK_LAUNCH("combine_heatmap_kernel");
  return SSP_OK;
}
