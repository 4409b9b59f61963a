// Label warping for the warped training pair, on the device.
// Reference: datasets/data_tools.py:37-63 (warpLabels), :6-34 (extrapolate_points / scatter_points / get_labels_bi),
//            call sites datasets/Coco.py:330,367 -- Gabriel-SGama/Semantic-SuperPoint.
//
//   pnts (x, y) truncated to integers -> warp_points with the PIXEL homography (homography_scaling_torch, computed by the
//   host exactly like the reference: a 3x3 fp32 inverse and two products) -> filter 0 <= p <= shape-1 -> scatter at
//   round-half-even(p):  labels = 1,  res = p - round(p),  [labels_bi = the four bilinear weights at trunc(p) + {0,1}^2].
// The reference scatters with a sequential index_put on the CPU: when several points land on one pixel THE LAST ONE WINS.
// Here: pass A takes an atomicMax of the point index per target pixel, pass B lets only the winner write -- the same
// result, in any execution order.  One block per image (<= a few thousand points), batched over the images of a step.
#include "common.cuh"

#define WL_THREADS 256

// order-preserving compaction of the kept points of one image (block-wide, chunks of WL_THREADS)
__device__ __forceinline__ int wl_block_offset(bool keep, int* s_warp, int& running) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  __syncthreads();
  if (lane == 0) s_warp[w] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int i = 0; i < WL_THREADS / 32; ++i) {
    const int c = s_warp[i];
    if (i < w) before += c;
    total += c;
  }
  const int off = running + before + __popc(bal & ((1u << lane) - 1u));
  running += total;
  return off;
}

__global__ void __launch_bounds__(WL_THREADS)
warp_labels_kernel(const float* __restrict__ pnts, const int* __restrict__ counts, int Pmax, int H, int W,
                   const float* __restrict__ Hpix, int bilinear, float* __restrict__ labels, float* __restrict__ res,
                   float* __restrict__ labels_bi, float* __restrict__ warped, int* __restrict__ kept,
                   int* __restrict__ win, int* __restrict__ win_bi) {
  __shared__ float h[9];
  __shared__ int s_warp[WL_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int P = counts[b];
  if (tid < 9) h[tid] = Hpix[b * 9 + tid];
  __syncthreads();
  const size_t plane = (size_t)H * W;
  pnts += (size_t)b * Pmax * 2;
  warped += (size_t)b * Pmax * 2;
  win += b * plane;
  labels += b * plane;
  res += b * plane * 2;
  if (bilinear) { win_bi += b * plane; labels_bi += b * plane; }
  const float sx = (float)W, sy = (float)H;

  auto warp_pt = [&](int i, float& wx, float& wy) {
    // pnts.long(): truncation toward zero, then back to float for warp_points (utils/utils.py:315-343)
    const float px = truncf(pnts[2 * i]), py = truncf(pnts[2 * i + 1]);
    homography_apply(h, px, py, wx, wy);
  };
  auto inb = [&](float x, float y) { return x >= 0.f && x <= sx - 1.f && y >= 0.f && y <= sy - 1.f; };  // filter_points, inclusive

  // ---- pass A: winners (the highest point index per target pixel = the last write of the sequential reference)
  for (int i = tid; i < P; i += WL_THREADS) {
    float wx, wy;
    warp_pt(i, wx, wy);
    if (inb(wx, wy)) atomicMax(win + (int)rintf(wy) * W + (int)rintf(wx), i);
    if (bilinear) {
      // extrapolate_points: base = trunc(p); order of the concatenation: (0,0), (0,+1), (+1,0), (+1,+1), each block of P points
      const float bx = truncf(wx), by = truncf(wy);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float ex = bx + (float)(t >> 1), ey = by + (float)(t & 1);
        if (inb(ex, ey)) atomicMax(win_bi + (int)rintf(ey) * W + (int)rintf(ex), t * P + i);
      }
    }
  }
  __syncthreads();  // one block owns the image: a block barrier orders the atomics before the reads below
  // ---- pass B: the winners write; kept points are compacted in their original order
  int running = 0;
  for (int i0 = 0; i0 < P; i0 += WL_THREADS) {
    const int i = i0 + tid;
    bool keep = false;
    float wx = 0.f, wy = 0.f;
    if (i < P) {
      warp_pt(i, wx, wy);
      keep = inb(wx, wy);
      if (keep) {
        const float rx = rintf(wx), ry = rintf(wy);
        const int q = (int)ry * W + (int)rx;
        if (win[q] == i) {
          labels[q] = 1.f;
          res[2 * q] = wx - rx;
          res[2 * q + 1] = wy - ry;
        }
      }
      if (bilinear) {
        const float bx = truncf(wx), by = truncf(wy);
        const float fx = wx - bx, fy = wy - by;  // residuals (x, y)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float ex = bx + (float)(t >> 1), ey = by + (float)(t & 1);
          if (inb(ex, ey)) {
            const int q = (int)rintf(ey) * W + (int)rintf(ex);
            if (win_bi[q] == t * P + i) {
              // (1-x)(1-y), (1-x) y, x (1-y), x y  in the order of the concatenation
              const float wgt = ((t >> 1) ? fx : 1.f - fx) * ((t & 1) ? fy : 1.f - fy);
              labels_bi[q] = wgt;
            }
          }
        }
      }
    }
    const int off = wl_block_offset(keep, s_warp, running);
    if (keep) { warped[2 * off] = wx; warped[2 * off + 1] = wy; }
  }
  if (tid == 0) kept[b] = running;
}

// pnts [B,Pmax,2] (x, y) fp32 with counts[b] valid points per image; Hpix [B,3,3] PIXEL homographies
// (homography_scaling_torch of the normalised ones).  Outputs: labels [B,1,H,W], res [B,H,W,2], labels_bi [B,1,H,W] (when
// bilinear), warped [B,Pmax,2] with kept[b] in-bounds warped points in their original order.  ws: 2 * B*H*W ints.
extern "C" size_t ssp_warp_labels_ws_bytes(int B, int H, int W) { return (size_t)2 * B * H * W * sizeof(int); }

extern "C" int ssp_warp_labels(const float* pnts, const int* counts, int B, int Pmax, int H, int W, const float* Hpix,
                               int bilinear, float* labels, float* res, float* labels_bi, float* warped, int* kept, void* ws,
                               size_t ws_bytes, void* stream) {
  SSP_REQUIRE(pnts && counts && Hpix && labels && res && warped && kept && ws, "ssp_warp_labels: null pointer");
  SSP_REQUIRE(!bilinear || labels_bi, "ssp_warp_labels: bilinear needs the labels_bi output");
  SSP_REQUIRE(B > 0 && Pmax >= 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "ssp_warp_labels: bad sizes");
  SSP_REQUIRE(4ll * Pmax < (1ll << 30), "ssp_warp_labels: too many points per image");
  SSP_REQUIRE(ws_bytes >= ssp_warp_labels_ws_bytes(B, H, W), "ssp_warp_labels: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  int* win = reinterpret_cast<int*>(ws);
  SSP_CUDA_CALL(cudaMemsetAsync(win, 0xFF, (bilinear ? 2 : 1) * n * sizeof(int), st));  // -1: no point lands here
  SSP_CUDA_CALL(cudaMemsetAsync(labels, 0, n * sizeof(float), st));
  SSP_CUDA_CALL(cudaMemsetAsync(res, 0, 2 * n * sizeof(float), st));
  if (bilinear) SSP_CUDA_CALL(cudaMemsetAsync(labels_bi, 0, n * sizeof(float), st));
  warp_labels_kernel<<<B, WL_THREADS, 0, st>>>(pnts, counts, Pmax, H, W, Hpix, bilinear, labels, res, labels_bi, warped, kept,
                                               win, win + n);
  SSP_CUDA_CHECK_LAUNCH("warp_labels_kernel");
  return SSP_OK;
}
