// Shared definitions for the dense descriptor hinge loss (reference: utils/utils.py:779-893).
//
// Decomposition used by every engine (SIMT fp32 and tcgen05):
//   The GEMM kernel evaluates the negative hinge max(dot - margin_neg, 0) over ALL pairs (no geometry in
//   its epilogue) and emits the indicator bit-matrix I[r,c] = 1[dot > margin_neg] in two orientations, so
//   that the backward is two pure indicator-GEMMs and never recomputes S.  The sparse POSITIVE pairs
//   (mask == 1, <= DESC_MAXP per row) are owned by the pos kernels: exact fp32 dots, weight lamda_d, and the
//   removal of the negative term the dense part contributed for them (forward sums and backward GEMM).
//
// Layouts:
//   wpts   [B, Nc_pad] float2  warped cell centres (x, y) in pixels; padded rows hold SSP_FAR
//   mv_pad [B, Nc_pad] float   mask_valid per warped-image cell, zero padded
//   bitsR  [B, Nc_pad/32, Nc_pad] u32   word (cw, r): columns 32cw..32cw+31 of row r
//   bitsC  [B, Nc_pad/32, Nc_pad] u32   word (rw, c): rows 32rw..32rw+31 of column c
//          element j of a word lives at bit DESC_BITPOS(j): even elements in the low half, odd ones in the high half, so
//          that the backward's A-operand expansion (bit pair (2i, 2i+1) -> one 32-bit register of two bf16) is one
//          shift and one and per register
//   mvbits [B, Nc_pad/32] u32           mask_valid != 0 per cell, same bit order (optional, "fold" mode)
//   Nc_pad = ceil(Nc / 256) * 256
#pragma once
#include "common.cuh"

#define SSP_FAR 1.0e30f
#define DESC_PAD 256
#define DESC_BITPOS(j) ((((j) >> 1) & 15) | (((j) & 1) << 4))
#define DESC_MAXP 16  // positive pairs per row / column kept in the sparse lists (descriptor_dist <= cell);
                      // colcnt[B*Nc_pad] counts entries that did not fit (0 unless the warp shrinks by > 3x)

struct DescGeom {
  int B, Hc, Wc, Nc, Nc_pad, Dch;
  int cell;          // cell size (8)
  float dist;        // descriptor_dist
  float lamda;       // lamda_d
  float mpos, mneg;  // margins 1.0 / 0.2
};

static inline int desc_nc_pad(int Nc) { return (Nc + DESC_PAD - 1) / DESC_PAD * DESC_PAD; }

// centre of cell index c in pixels, (x, y)   [utils/utils.py:829-831]
__device__ __forceinline__ void cell_center(int c, int Wc, int cell, float& cx, float& cy) {
  int k = c / Wc, l = c - k * Wc;
  cy = (float)(k * cell + cell / 2);
  cx = (float)(l * cell + cell / 2);
}

// mask = ||centre - warped|| <= dist, evaluated in fp32 exactly in this order everywhere
// [utils/utils.py:854-857]
__device__ __forceinline__ bool pair_positive(float wx, float wy, float cx, float cy, float dist) {
  float dy = cy - wy, dx = cx - wx;
  float d2 = __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx));
  return __fsqrt_rn(d2) <= dist;
}

// 32x32 bit-matrix transpose across a warp (lane r holds row r as a word): five shuffle / mask stages (block-swap
// recursion).  out[lane c] bit r = in[lane r] bit c.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t a, int lane) {
  // out[lane c] bit r = in[lane r] bit c
#define SSP_TSTEP(J, M)                                                \
  {                                                                    \
    const uint32_t o = __shfl_xor_sync(0xffffffffu, a, J);             \
    if (lane & J) a ^= ((o >> J) ^ a) & (M);                            \
    else          a ^= (((a >> J) ^ o) & (M)) << J;                     \
  }
  SSP_TSTEP(16, 0x0000FFFFu)
  SSP_TSTEP(8, 0x00FF00FFu)
  SSP_TSTEP(4, 0x0F0F0F0Fu)
  SSP_TSTEP(2, 0x33333333u)
  SSP_TSTEP(1, 0x55555555u)
#undef SSP_TSTEP
  return a;
}

