// tcgen05 / TMEM / TMA engine for the dense descriptor loss (sm_100a).
// Reference semantics: utils/utils.py:863-890 (all-pairs dot product + hinge + masks + reductions).
//
// Operands are bf16 K-major planes [B, Nc_pad, 256] produced by desc_pack_kernel: plane "hi" alone
// (mode bf16) or hi + lo with three MMA chains hi*hi + lo*hi + hi*lo (mode bf16x3, ~2^-16 relative
// product error = fp32-grade dot products on the tensor pipe).
//
// Forward kernel, persistent 2-CTA clusters over (pair b, 128-row tile, 256-column tile) items:
//   warp 8      TMA producer: A rows (all 256 channels, resident) once, then the B ring of 32 KB
//               K-chunks (64 channels x 256 cells, SWIZZLE_128B) for every 256-column tile
//   warp 9      MMA issuer: tcgen05.mma kind::f16 M=128 N=256 K=16 into a double-buffered TMEM
//               accumulator (2 x 256 columns = all of TMEM); tcgen05.commit frees ring slots / publishes tiles
//   warps 0-7   epilogue: tcgen05.ld 32x32b (one row per thread, 128 columns per warp), hinge,
//               mask_valid weighting, running sums, indicator bit-matrix in both orientations.
//               No geometry here: the sparse positive pairs are corrected by the pos kernels.  The
//               pair matrix never reaches HBM.
//
// Backward kernel = indicator GEMM  out[b, d, r] = rowscale[r] * sum_k bit(r,k) * Bp[b, k, d], persistent over
// (pair, row-tile pair, channel half) items:
//   warps 0-3   expand 64 indicator bits per row into bf16 {0,1} and tcgen05.st them as the A
//               operand into TMEM (A never touches shared memory)
//   warps 4-11  epilogue of the previous item (double-buffered 128-column accumulators), incl. the sparse
//               positive-pair terms
//   warp 12     TMA producer of B tiles [64 cells x 128 channels] (MN-major, SWIZZLE_128B, multicast)
//   warp 13     MMA issuer, M=128 N=128 K=16, A from TMEM
#include "desc_common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstdlib>

// ---- optional timeline trace (build with -DSSP_TRACE, see scripts/trace_desc.py): lane 0 of one warp per role appends
// (clock64 << 8 | tag) records to a per-CTA, per-role slice of the buffer registered with ssp_debug_trace().
#define TRACE_CAP 4096
#ifdef SSP_TRACE
__device__ long long* g_trace_buf = nullptr;
#define TR_DECL(role) long long* tr_ = g_trace_buf ? g_trace_buf + ((size_t)blockIdx.x * 4 + (role)) * TRACE_CAP : nullptr; int trn_ = 0
#define TR(tag) do { if (tr_ && trn_ < TRACE_CAP) tr_[trn_++] = (clock64() << 8) | (long long)(tag); } while (0)
#else
#define TR_DECL(role) do { } while (0)
#define TR(tag) do { } while (0)
#endif

namespace {

constexpr int BM = 128;        // rows per CTA
constexpr int BN = 256;        // columns per accumulator tile (one MMA instruction = M128 x N256 x K16, 128 cycles)
constexpr int KD = 256;        // descriptor channels (GEMM K of the forward)
constexpr int KC = 64;         // channels per smem chunk = 128 B of bf16 = one swizzle row
constexpr int NKC = KD / KC;   // 4
constexpr int CHUNK_BYTES = BM * KC * 2;  // 16 KB: one A chunk (128 rows x 64 channels)
constexpr int BCHUNK_BYTES = BN * KC * 2; // 32 KB: one B chunk (256 cells x 64 channels)
constexpr int FWD_THREADS = 320;
constexpr int BAR_BYTES = 1536;  // mbarriers, TMEM pointer, warp partial sums, 8 x 32-word ballot scratch

template <int P> struct FwdCfg {
  static constexpr int NSTAGE = (P == 1) ? 4 : 3;
  static constexpr int A_BYTES = P * NKC * CHUNK_BYTES;
  static constexpr int B_BYTES = NSTAGE * BCHUNK_BYTES;
  static constexpr int SMEM = A_BYTES + B_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// One 32-column slice of the accumulator row owned by this thread: negative hinge over every pair (the sparse
// positive pairs are corrected by the pos kernels), mask_valid weighting, indicator bits in both orientations.
// Row-orientation bits are OR-ed into a register; column-orientation words are warp ballots (bit r = row r of this
// warp), parked in a 32-word shared scratch by lane 0 and picked up one per lane after the loop.
template <bool BITS>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], const float* __restrict__ mvp, float mneg, float& su,
                                          float& sw, uint32_t& rowword, uint32_t& colword, uint32_t* __restrict__ scratch,
                                          int lane) {
  rowword = 0;
  colword = 0;
  float su1 = 0.f, sw1 = 0.f;  // second accumulator pair: halves the dependent-add chains
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    float4 m4 = __ldg(reinterpret_cast<const float4*>(mvp) + j4);
    float mvv[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j4 * 4 + jj;
      float neg = fmaxf(__uint_as_float(v[j]) - mneg, 0.f);
      if (jj & 1) { su1 += neg; sw1 = fmaf(neg, mvv[jj], sw1); }
      else        { su += neg;  sw = fmaf(neg, mvv[jj], sw); }
      if (BITS) {
        bool p = neg > 0.f;
        rowword |= p ? (1u << j) : 0u;
        uint32_t bal = __ballot_sync(0xffffffffu, p);
        if (lane == 0) scratch[j] = bal;
      }
    }
  }
  if (BITS) {
    __syncwarp();
    colword = scratch[lane];
    __syncwarp();
  }
  su += su1;
  sw += sw1;
}

// Variant of epi_chunk with the 32 mask_valid values already in registers (EPI2 epilogue, see the kernel).
template <bool BITS>
__device__ __forceinline__ void epi_chunk_pre(const uint32_t (&v)[32], const float4 (&mq)[8], float mneg, float& su,
                                              float& sw, uint32_t& rowword, uint32_t& colword,
                                              uint32_t* __restrict__ scratch, int lane, bool maskbits) {
  rowword = 0;
  colword = 0;
  float su1 = 0.f, sw1 = 0.f;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float mvv[4] = {mq[j4].x, mq[j4].y, mq[j4].z, mq[j4].w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j4 * 4 + jj;
      float neg = fmaxf(__uint_as_float(v[j]) - mneg, 0.f);
      if (jj & 1) { su1 += neg; sw1 = fmaf(neg, mvv[jj], sw1); }
      else        { su += neg;  sw = fmaf(neg, mvv[jj], sw); }
      if (BITS) {
        bool p = neg > 0.f;
        // maskbits: the ROW-orientation indicator (consumed by the dD GEMM) drops columns whose mask_valid is 0, so that
        // GEMM can run on the unscaled forward planes of Dw with one scalar in its epilogue (alpha_c = s * mv_c for a
        // binary mask and g_neg = 0); the column orientation (dDw GEMM, row-scaled by alpha_c) stays complete
        rowword |= (p && (!maskbits || mvv[jj] != 0.f)) ? (1u << j) : 0u;
        uint32_t bal = __ballot_sync(0xffffffffu, p);
        if (lane == 0) scratch[j] = bal;
      }
    }
  }
  if (BITS) {
    __syncwarp();
    colword = scratch[lane];
    __syncwarp();
  }
  su += su1;
  sw += sw1;
}

// float -> double by bit manipulation (exact for normal numbers and zero; denormals flush to zero, inf / nan kept):
// keeps the epilogue off the FP64 pipe, whose conversions and adds took 28 % of the kernel's stall samples.
__device__ __forceinline__ double f32_to_f64_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t e = (u >> 23) & 0xffu;
  const unsigned long long sign = (unsigned long long)(u & 0x80000000u) << 32;
  unsigned long long d;
  if (e == 0u) d = sign;
  else if (e == 0xffu) d = sign | (0x7ffull << 52) | ((unsigned long long)(u & 0x7fffffu) << 29);
  else d = sign | ((unsigned long long)(e + 896u) << 52) | ((unsigned long long)(u & 0x7fffffu) << 29);
  return __longlong_as_double((long long)d);
}

// Persistent forward kernel.  Work item = (pair b, row-tile pair mp, column tile nt); the flattened item range is
// split evenly over the clusters (one 2-CTA cluster per SM pair), so all SMs finish together instead of running
// 2.16 waves of whole row tiles.  Within a cluster CTA rank r owns row tile 2*mp + r; A is reloaded only when
// (b, mp) changes (items are contiguous in nt).  Per-item, per-warp partial sums go to
// partials[((item*2 + rank)*8 + warp)*2 + {0,1}].
// EPI2 (opt-in, SSP_FWD_EPI=2; written from the round-1 stall analysis, to be validated on hardware before it becomes the
// default): tile sums stay in fp32 and are converted once per item without the FP64 pipe, and the mask_valid values of a
// chunk are fetched one chunk ahead (the first chunk's before the accumulator wait) instead of inside the chunk.
template <int P, bool BITS, bool EPI2>
__global__ void __launch_bounds__(FWD_THREADS, 1)
desc_dense_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                         const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                         const float* __restrict__ mv_pad, DescGeom g,
                         double* __restrict__ partials, uint32_t* __restrict__ bitsR, uint32_t* __restrict__ bitsC,
                         float* __restrict__ dbgS) {
  using Cfg = FwdCfg<P>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = sA + Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::B_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;
  uint64_t* b_empty = b_full + NSTAGE;
  uint64_t* t_full = b_empty + NSTAGE;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  uint32_t* ballot_scratch = tmem_ptr + 2;  // 8 warps x 32 words

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MP = g.Nc_pad / (2 * BM), NT = g.Nc_pad / BN;  // Nc_pad is a multiple of 256
  const uint32_t cta_rank = tc::cluster_ctarank();
  const int ncluster = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const long long T = (long long)g.B * MP * NT;
  const int it0 = (int)(T * cid / ncluster), it1 = (int)(T * (cid + 1) / ncluster);

  if (warp == 8 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi);
    tc::prefetch_tmap(&tmB_hi);
    if (P == 2) { tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmB_lo); }
  }
  if (warp == 9) {
    if (lane == 0) {
      tc::mbar_init(a_full, 1);
      tc::mbar_init(a_empty, 1);
      for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 2); }  // 2 = both CTAs' MMAs
      for (int s = 0; s < 2; ++s) { tc::mbar_init(t_full + s, 1); tc::mbar_init(t_empty + s, 8); }
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_ptr, 512);
  }
  tc::fence_before_sync();
  tc::cluster_sync_all();  // barriers of both CTAs are live before any remote arrive / multicast write
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      TR_DECL(1);
      int prev_key = -1, stage_it = 0;
      uint32_t a_empty_ph = 0;
      for (int it = it0; it < it1; ++it) {
        const int key = it / NT, nt = it - key * NT;  // key = b * MP + mp
        const int b = key / MP, mp = key - b * MP;
        const int row_base = b * g.Nc_pad;
        if (key != prev_key) {
          TR(10);
          if (prev_key >= 0) { tc::mbar_wait(a_empty, a_empty_ph); a_empty_ph ^= 1; }  // MMAs on the old A rows are done
          TR(11);
          tc::mbar_expect_tx(a_full, Cfg::A_BYTES);
          for (int p = 0; p < P; ++p)
            for (int kc = 0; kc < NKC; ++kc)
              tc::tma_load_2d(p == 0 ? &tmA_hi : &tmA_lo, a_full, sA + (p * NKC + kc) * CHUNK_BYTES, kc * KC,
                              row_base + (2 * mp + (int)cta_rank) * BM);
          prev_key = key;
        }
        for (int kc = 0; kc < NKC; ++kc)
          for (int p = 0; p < P; ++p, ++stage_it) {
            int s = stage_it % NSTAGE;
            uint32_t ph = (stage_it / NSTAGE) & 1;
            TR(12);
            tc::mbar_wait(b_empty + s, ph ^ 1);  // slot s is free in BOTH CTAs
            TR(13);
            tc::mbar_expect_tx(b_full + s, BCHUNK_BYTES);
            // my half (128 of the 256 cells) of the chunk, written into both CTAs' slot s
            tc::tma_load_2d_mc(p == 0 ? &tmB_hi : &tmB_lo, b_full + s, sB + s * BCHUNK_BYTES + cta_rank * (BCHUNK_BYTES / 2),
                               kc * KC, row_base + nt * BN + cta_rank * (BN / 2), (uint16_t)0x3);
          }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16_f32(BM, BN, 0, 0);
      const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
      TR_DECL(0);
      int prev_key = -1, stage_it = 0, tcount = 0;
      uint32_t a_full_ph = 0;
      for (int it = it0; it < it1; ++it, ++tcount) {
        const int key = it / NT;
        TR(1);
        if (key != prev_key) { tc::mbar_wait(a_full, a_full_ph); a_full_ph ^= 1; prev_key = key; }
        TR(2);
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        tc::mbar_wait(t_empty + as, aph ^ 1);
        TR(3);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        uint32_t first = 1;
        for (int kc = 0; kc < NKC; ++kc)
          for (int p = 0; p < P; ++p, ++stage_it) {
            int s = stage_it % NSTAGE;
            uint32_t ph = (stage_it / NSTAGE) & 1;
            TR(4);
            tc::mbar_wait(b_full + s, ph);
            TR(5);
            tc::fence_after_sync();
            // B plane p (0 = hi, 1 = lo) meets A hi; B hi additionally meets A lo (lo*lo is dropped)
            const int n_a = (P == 2 && p == 0) ? 2 : 1;
            const uint64_t db = tc::smem_desc_sw128(sB_u + s * BCHUNK_BYTES, 16, 1024);
            for (int pa = 0; pa < n_a; ++pa) {
              const uint64_t da = tc::smem_desc_sw128(sA_u + (pa * NKC + kc) * CHUNK_BYTES, 16, 1024);
              tc::mma_ss_x4(d_tmem, da, db, idesc, first ? 0u : 1u);  // 4 x (M128 N256 K16) over this 64-channel chunk
              first = 0;
            }
            tc::mma_commit_mc(b_empty + s, (uint16_t)0x3);  // tell both producers: this CTA is done with slot s
            TR(6);
          }
        tc::mma_commit(t_full + as);  // accumulator tile complete
        const int next_key = (it + 1 < it1) ? (it + 1) / NT : -2;
        if (next_key != key) tc::mma_commit(a_empty);  // the A rows may be replaced once these MMAs have drained
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue warps 0..7 ------------------------------
    const int q = warp & 3, half = warp >> 2;
    const int NW = g.Nc_pad / 32;
    int tcount = 0;
    TR_DECL(2 + (warp == 7 ? 1 : 0));
#ifdef SSP_TRACE
    if (!((warp == 0 || warp == 7) && lane == 0)) tr_ = nullptr;
#endif
    for (int it = it0; it < it1; ++it, ++tcount) {
      const int key = it / NT, nt = it - key * NT;
      const int b = key / MP, mp = key - b * MP;
      const int row_base = b * g.Nc_pad;
      const int m0 = (2 * mp + (int)cta_rank) * BM;
      const int row = m0 + q * 32 + lane;  // row inside the padded pair
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      float4 mq[8];
      if (EPI2) {
        const float4* mvq = reinterpret_cast<const float4*>(mv_pad + (size_t)row_base + nt * BN + half * 128);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) mq[j4] = __ldg(mvq + j4);
      }
      TR(30);
      tc::mbar_wait(t_full + as, aph);
      TR(31);
      tc::fence_after_sync();
      double su_d = 0.0, sw_d = 0.0;
      float su_t = 0.f, sw_t = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int cbase = nt * BN + half * 128 + ch * 32;
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * 128 + ch * 32, v);
        uint32_t rowword, colword;
        float su = 0.f, sw = 0.f;
        if (EPI2) {
          float4 mnext[8];
          if (ch < 3) {
            const float4* mvq = reinterpret_cast<const float4*>(mv_pad + (size_t)row_base + cbase + 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) mnext[j4] = __ldg(mvq + j4);
          }
          tc::tmem_ld_wait();
          epi_chunk_pre<BITS>(v, mq, g.mneg, su, sw, rowword, colword, ballot_scratch + warp * 32, lane, g.cell != 0);
          su_t += su;
          sw_t += sw;
          if (ch < 3) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) mq[j4] = mnext[j4];
          }
        } else {
          tc::tmem_ld_wait();
          const float* mvp = mv_pad + (size_t)row_base + cbase;
          epi_chunk<BITS>(v, mvp, g.mneg, su, sw, rowword, colword, ballot_scratch + warp * 32, lane);
          su_d += (double)su;
          sw_d += (double)sw;
        }
        if (BITS) {
          bitsR[((size_t)b * NW + cbase / 32) * g.Nc_pad + row] = rowword;
          bitsC[((size_t)b * NW + (m0 + q * 32) / 32) * g.Nc_pad + cbase + lane] = colword;
        }
        if (dbgS && row < g.Nc) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cbase + j < g.Nc) dbgS[((size_t)b * g.Nc + row) * g.Nc + cbase + j] = __uint_as_float(v[j]);
        }
      }
      // all TMEM reads of this stage are complete (wait::ld above): hand the stage back
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(t_empty + as);
      TR(32);
      if (EPI2) {
        su_t = warp_sum(su_t);
        sw_t = warp_sum(sw_t);
        if (lane == 0) {
          size_t slot = (((size_t)it * 2 + cta_rank) * 8 + warp) * 2;
          partials[slot] = f32_to_f64_bits(su_t);
          partials[slot + 1] = f32_to_f64_bits(sw_t);
        }
      } else {
        su_d = warp_sum_d(su_d);
        sw_d = warp_sum_d(sw_d);
        if (lane == 0) {
          size_t slot = (((size_t)it * 2 + cta_rank) * 8 + warp) * 2;
          partials[slot] = su_d;
          partials[slot + 1] = sw_d;
        }
      }
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync_all();  // nobody exits while the peer may still multicast into / arrive on this CTA
  if (warp == 9) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// indicator GEMM (backward)
// ------------------------------------------------------------------------------------------------
constexpr int KT = 64;                       // cells (GEMM K) per stage
constexpr int BG_BOX_BYTES = KT * 128;       // one [64 cells x 64 channels] box = 8 KB
constexpr int BG_N = 128;                    // channels per work item (half of the descriptor)
constexpr int BG_THREADS = 448;              // warps 0-3 expanders, 4-11 epilogue, 12 TMA producer, 13 MMA issuer

// NS_ = depth of the B (TMA) / A (TMEM) stage ring.  P = 2: 32 KB per stage, 4 stages = 128 KB (default) or 6 = 192 KB;
// P = 1: 16 KB per stage, 6 (default) or 8.  The TMEM A ring needs 256 + 32 * NS columns <= 512, i.e. NS <= 8.
template <int P, int NS_> struct BgCfg {
  static constexpr int NS = NS_;
  static constexpr int STAGE_BYTES = P * 2 * BG_BOX_BYTES;  // P planes x two 64-channel boxes = 16 KB per plane
  static constexpr int SMEM = NS * STAGE_BYTES + BAR_BYTES + 1024;
};

// Persistent indicator GEMM.  Work item = (pair b, row-tile pair mp, channel half dh); flattened item range split
// evenly over 2-CTA clusters.  The fp32 accumulator is double buffered in TMEM (2 x 128 columns), so the epilogue of
// item i (warps 4-7: TMEM -> registers -> coalesced NCHW stores) overlaps the main loop of item i+1 (warps 0-3 expand
// indicator bits into the TMEM A ring, warp 8 streams B through TMA multicast, warp 9 issues the MMAs).
// DEEPBITS (opt-in, SSP_BG_BITS=deep; from the round-1 stall analysis, to be validated on hardware): the expanders fetch
// the indicator words two 4-stage groups ahead over the FLAT stage sequence of all items of the cluster, so neither the
// L2 latency inside an item nor the cold start of every item (no prefetch across the item boundary today) is exposed.
// LATEPOS (opt-in, SSP_BG_POS=late; same status): the epilogue stores the scaled accumulator first, hands the TMEM stage
// back, and only then adds the sparse positive-pair terms as a read-modify-write of its own stores (bit-identical
// arithmetic), so the partner gathers (DRAM latency) no longer extend the time the accumulator stage is held.
template <int P, int NS_, bool DEEPBITS, bool LATEPOS>
__global__ void __launch_bounds__(BG_THREADS, 1)
desc_bits_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                         const uint32_t* __restrict__ bits, const float* __restrict__ rowscale,
                         const int* __restrict__ plist, const float* __restrict__ pcoef, const float* __restrict__ possrc,
                         const uint4* __restrict__ pos_hi, const uint4* __restrict__ pos_lo,
                         int B, int Nc, int Nc_pad, int sched, float* __restrict__ out) {
  using Cfg = BgCfg<P, NS_>;
  constexpr int NS = Cfg::NS;
  static_assert(256 + 32 * NS <= 512, "TMEM A ring does not fit");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * Cfg::STAGE_BYTES);
  uint64_t* b_full = bars;
  uint64_t* a_full = b_full + NS;
  uint64_t* s_free = a_full + NS;
  uint64_t* d_full = s_free + NS;
  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);
  // 4 indicator bits -> 4 bf16 {0,1} (two 32-bit TMEM columns); 16 entries x 8 B cover all 32 banks exactly once
  uint2* lut = reinterpret_cast<uint2*>(tmem_ptr + 2);
  if (threadIdx.x < 16) {
    uint32_t x = threadIdx.x;
    lut[x] = make_uint2((x & 1u) * 0x3F80u + ((x >> 1) & 1u) * 0x3F800000u, ((x >> 2) & 1u) * 0x3F80u + ((x >> 3) & 1u) * 0x3F800000u);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MP = Nc_pad / (2 * BM), NK = Nc_pad / KT, NW = Nc_pad / 32;
  const uint32_t cta_rank = tc::cluster_ctarank();
  const int ncluster = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const long long T = (long long)B * MP * 2;
  // item schedule.  0: contiguous ranges per cluster.  1: round-robin (cluster cid takes items cid, cid + ncluster, ...):
  // at any moment the 74 clusters work on ~7 consecutive pairs b, so the B planes of a pair (1.3 MB) are fetched from
  // HBM once and hit in L2 for the other row tiles, instead of all 32 pairs' planes (42 MB + outputs) cycling through L2.
  const int it0 = sched ? cid : (int)(T * cid / ncluster);
  const int it1 = sched ? (int)T : (int)(T * (cid + 1) / ncluster);
  const int itstep = sched ? ncluster : 1;
  constexpr uint32_t A_COL0 = 256;  // TMEM: accumulators at columns [0,128) and [128,256), then NS x 32 columns of A

  if (warp == 12 && lane == 0) {
    tc::prefetch_tmap(&tmB_hi);
    if (P == 2) tc::prefetch_tmap(&tmB_lo);
  }
  if (warp == 13) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) { tc::mbar_init(b_full + s, 1); tc::mbar_init(a_full + s, 128); tc::mbar_init(s_free + s, 2); }
      for (int s = 0; s < 2; ++s) { tc::mbar_init(d_full + s, 1); tc::mbar_init(d_empty + s, 8); }
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_ptr, 512);
  }
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 12) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      TR_DECL(1);
      int st_it = 0;
      for (int it = it0; it < it1; it += itstep) {
        const int key = it >> 1, dh = it & 1;
        const int b = key / MP;
        const int row_base = b * Nc_pad;
        for (int kc = 0; kc < NK; ++kc, ++st_it) {
          int s = st_it % NS;
          uint32_t ph = (st_it / NS) & 1;
          TR(12);
          tc::mbar_wait(s_free + s, ph ^ 1);
          TR(13);
          tc::mbar_expect_tx(b_full + s, Cfg::STAGE_BYTES);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          for (int p = 0; p < P; ++p)
            for (int dc = 0; dc < 2; ++dc)  // my 32 of the 64 cells of every box, multicast to both CTAs
              tc::tma_load_2d_mc(p == 0 ? &tmB_hi : &tmB_lo, b_full + s,
                                 st + (p * 2 + dc) * BG_BOX_BYTES + cta_rank * (BG_BOX_BYTES / 2), dh * BG_N + dc * 64,
                                 row_base + kc * KT + cta_rank * (KT / 2), (uint16_t)0x3);
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16_f32(BM, BG_N, 0, 1);  // A K-major (TMEM), B MN-major
      const uint32_t smem_u = tc::smem_u32(smem);
      TR_DECL(0);
      int st_it = 0, tcount = 0;
      for (int it = it0; it < it1; it += itstep, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        TR(2);
        tc::mbar_wait(d_empty + as, aph ^ 1);
        TR(3);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BG_N;
        uint32_t first = 1;
        for (int kc = 0; kc < NK; ++kc, ++st_it) {
          int s = st_it % NS;
          uint32_t ph = (st_it / NS) & 1;
          TR(4);
          tc::mbar_wait(b_full + s, ph);
          TR(5);
          tc::mbar_wait(a_full + s, ph);
          TR(7);
          tc::fence_after_sync();
          for (int p = 0; p < P; ++p) {
            // per K=16 step: 16 cells = two 8-row groups (SBO 1024 B, +2048 B per step); 128 channels = two
            // 64-wide blocks (LBO 8 KB); A advances 8 TMEM columns per step
            uint64_t db = tc::smem_desc_sw128(smem_u + s * Cfg::STAGE_BYTES + p * 2 * BG_BOX_BYTES, BG_BOX_BYTES, 1024);
            tc::mma_ts_x4(d_tmem, tmem_base + A_COL0 + s * 32, db, idesc, first ? 0u : 1u);
            first = 0;
          }
          tc::mma_commit_mc(s_free + s, (uint16_t)0x3);
          TR(6);
        }
        tc::mma_commit(d_full + as);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ------------------------------ expanders: indicator bits -> bf16 A operand in TMEM ------------------------------
    const int q = warp;
    int st_it = 0;
    TR_DECL(2);
#ifdef SSP_TRACE
    if (!(warp == 0 && lane == 0)) tr_ = nullptr;
#endif
    // one ring stage: expand the 64 indicator bits (w0, w1) of this row into 32 TMEM columns of bf16 {0,1}
    auto expand_stage = [&](uint32_t w0, uint32_t w1) {
      int s = st_it % NS;
      uint32_t ph = (st_it / NS) & 1;
      TR(20);
      tc::mbar_wait(s_free + s, ph ^ 1);
      TR(21);
      tc::fence_after_sync();
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint2 e0 = lut[(w0 >> (4 * i)) & 15u], e1 = lut[(w1 >> (4 * i)) & 15u];
        r[2 * i] = e0.x; r[2 * i + 1] = e0.y;
        r[16 + 2 * i] = e1.x; r[16 + 2 * i + 1] = e1.y;
      }
      TR(22);
      tc::tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + s * 32, r);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(a_full + s);
      TR(23);
      ++st_it;
    };
    if (DEEPBITS) {
      const int GPI = NK / 4;  // 4-stage groups per item (NK = Nc_pad / 64 is a multiple of 4)
      const int nitems = it1 > it0 ? (it1 - it0 + itstep - 1) / itstep : 0;
      const int ngroups = nitems * GPI;
      auto load_group = [&](int gi, uint32_t (&dst)[8]) {
        if (gi < ngroups) {
          const int itn = it0 + (gi / GPI) * itstep, kc0 = (gi % GPI) * 4;
          const int key = itn >> 1;
          const int b = key / MP, mp = key - b * MP;
          const int row = (2 * mp + (int)cta_rank) * BM + q * 32 + lane;
          const uint32_t* brow = bits + (size_t)b * NW * Nc_pad + row;
#pragma unroll
          for (int u = 0; u < 8; ++u) dst[u] = __ldg(brow + (size_t)(kc0 * 2 + u) * Nc_pad);
        }
      };
      uint32_t c0[8], c1[8], c2[8];
      load_group(0, c0);
      load_group(1, c1);
      for (int gi = 0; gi < ngroups; ++gi) {
        load_group(gi + 2, c2);
#pragma unroll
        for (int u = 0; u < 4; ++u) expand_stage(c0[2 * u], c0[2 * u + 1]);
#pragma unroll
        for (int u = 0; u < 8; ++u) { c0[u] = c1[u]; c1[u] = c2[u]; }
      }
    } else {
      for (int it = it0; it < it1; it += itstep) {
        const int key = it >> 1;
        const int b = key / MP, mp = key - b * MP;
        const int row = (2 * mp + (int)cta_rank) * BM + q * 32 + lane;
        const uint32_t* brow = bits + (size_t)b * NW * Nc_pad + row;
        // indicator words are fetched four stages (8 words) ahead: the loads of group g+1 are in flight while group g
        // is expanded, so the HBM/L2 latency of the bit matrix never sits on the per-stage critical path
        uint32_t cur[8], nxt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) cur[u] = __ldg(brow + (size_t)u * Nc_pad);
        for (int kc0 = 0; kc0 < NK; kc0 += 4) {  // NK = Nc_pad / 64 is a multiple of 4
          if (kc0 + 4 < NK) {
#pragma unroll
            for (int u = 0; u < 8; ++u) nxt[u] = __ldg(brow + (size_t)((kc0 + 4) * 2 + u) * Nc_pad);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) expand_stage(cur[2 * u], cur[2 * u + 1]);
#pragma unroll
          for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
        }
      }
    }
  } else {
    // ------------------------------ epilogue warps 4..11 ------------------------------
    // two warps per TMEM lane quadrant, each draining two of the four 32-column chunks of the accumulator
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    int tcount = 0;
    TR_DECL(3);
#ifdef SSP_TRACE
    if (!(warp == 4 && lane == 0)) tr_ = nullptr;
#endif
    for (int it = it0; it < it1; it += itstep, ++tcount) {
      const int key = it >> 1, dh = it & 1;
      const int b = key / MP, mp = key - b * MP;
      const int row = (2 * mp + (int)cta_rank) * BM + q * 32 + lane;
      const bool row_ok = row < Nc;
      float rs = 1.f;
      if (rowscale && row_ok) rs = rowscale[(size_t)b * Nc_pad + row];
      // sparse positive pairs of this row (and removal of their negative term, see desc_pos_coef_kernel): their
      // gathers hide behind the next item's main loop because this epilogue runs on its own warps
      const int* pl = plist ? plist + ((size_t)b * Nc_pad + row) * DESC_MAXP : nullptr;
      const float* pcf = plist ? pcoef + ((size_t)b * Nc_pad + row) * DESC_MAXP : nullptr;
      int npos = 0;
      if (pl && row_ok) {
#pragma unroll
        for (int n = 0; n < DESC_MAXP; ++n)
          if (pl[n] >= 0) npos = n + 1;
      }
      const int nmax = __reduce_max_sync(0xffffffffu, npos);
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      TR(30);
      tc::mbar_wait(d_full + as, aph);
      TR(31);
      tc::fence_after_sync();
      // adds the positive-pair terms of partner list entry n to the 32 channels of chunk ch held in val[]
      auto add_partner = [&](int n, int ch, float (&val)[32]) {
        int pc = n < npos ? pl[n] : -1;
        if (pc < 0) return;
        float pf = pcf[n];
        if (pos_hi) {
          // partner descriptor from the packed planes: the 32 channels of this chunk are 64 contiguous bytes per
          // plane (4 x 16 B per lane, every fetched sector fully used) instead of 32 words 4*Nc bytes apart
          const size_t o = (((size_t)b * Nc_pad + pc) * KD + dh * BG_N + ch * 32) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 h = __ldg(pos_hi + o + q);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
            uint32_t lw[4] = {0u, 0u, 0u, 0u};
            if (pos_lo) {
              const uint4 l = __ldg(pos_lo + o + q);
              lw[0] = l.x; lw[1] = l.y; lw[2] = l.z; lw[3] = l.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float e0 = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
              float e1 = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
              val[q * 8 + 2 * i] = fmaf(pf, e0, val[q * 8 + 2 * i]);
              val[q * 8 + 2 * i + 1] = fmaf(pf, e1, val[q * 8 + 2 * i + 1]);
            }
          }
        } else {
          const float* ps = possrc + ((size_t)b * KD + dh * BG_N + ch * 32) * Nc + pc;
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = fmaf(pf, __ldg(ps + (size_t)j * Nc), val[j]);
        }
      };
#pragma unroll 1
      for (int ch = chalf * 2; ch < chalf * 2 + 2; ++ch) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BG_N + ch * 32, v);
        tc::tmem_ld_wait();
        if (row_ok) {
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(v[j]) * rs;
          if (!LATEPOS) {
#pragma unroll 1
            for (int n = 0; n < nmax; ++n) add_partner(n, ch, val);
          }
          float* o = out + ((size_t)b * KD + dh * BG_N + ch * 32) * Nc + row;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)j * Nc] = val[j];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(d_empty + as);
      TR(32);
      if (LATEPOS && nmax > 0 && row_ok) {
        // the accumulator stage is free again; this thread re-reads its own stores and adds the partner terms
#pragma unroll 1
        for (int ch = chalf * 2; ch < chalf * 2 + 2; ++ch) {
          float* o = out + ((size_t)b * KD + dh * BG_N + ch * 32) * Nc + row;
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = o[(size_t)j * Nc];
#pragma unroll 1
          for (int n = 0; n < nmax; ++n) add_partner(n, ch, val);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)j * Nc] = val[j];
        }
      }
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync_all();
  if (warp == 13) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// packed plane [rows, 256] bf16 -> 2-D map with a [box_rows x 64] SWIZZLE_128B box
int make_plane_map(CUtensorMap* m, const void* base, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { ssp_set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return SSP_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)KD, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)KD * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ssp_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SSP_EARG; }
  return SSP_OK;
}

// launch with thread-block clusters of 2 along x
template <typename... KArgs, typename... Args>
int launch_cluster2(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) { ssp_set_error("cluster launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return SSP_OK;
}

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { ssp_set_error("cudaFuncSetAttribute(%d B smem) failed: %s", bytes, cudaGetErrorString(e)); return (int)e; }
  return SSP_OK;
}

}  // namespace

// Profiling aid: registers a device buffer of n_cta * 4 * TRACE_CAP int64 records for the timeline trace of the two tensor-core
// kernels (library built with -DSSP_TRACE only; otherwise SSP_EUNSUPPORTED).  NULL switches tracing off.
extern "C" int ssp_debug_trace(void* buf) {
#ifdef SSP_TRACE
  long long* p = reinterpret_cast<long long*>(buf);
  SSP_CUDA_CALL(cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)));
  return SSP_OK;
#else
  (void)buf;
  ssp_set_error("ssp_debug_trace: library built without -DSSP_TRACE");
  return SSP_EUNSUPPORTED;
#endif
}
extern "C" int ssp_debug_trace_cap(void) { return TRACE_CAP; }

// partial-sum slots of the forward kernel: one per (item, cluster rank, epilogue warp)
extern "C" int ssp_desc_dense_tc_nblocks(int B, int Nc) {
  int ncp = desc_nc_pad(Nc);
  return B * (ncp / (2 * BM)) * (ncp / BN) * 2 * 8;
}

// Ahi/Alo: packed planes of `descriptors`, Bhi/Blo: packed planes of `descriptors_warped`
// ([B, Nc_pad, 256] bf16).  Alo == Blo == NULL selects single-pass bf16; otherwise bf16x3.
// flags bit 0: bitsR drops the columns whose mask_valid is 0 (see epi_chunk_pre); only the EPI2 epilogue implements it
extern "C" int ssp_desc_dense_fwd_tc_ex(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo,
                                        const float* mv_pad, int B, int Hc, int Wc, float mneg, double* partials,
                                        uint32_t* bitsR, uint32_t* bitsC, float* dbgS, int flags, void* stream);

extern "C" int ssp_desc_dense_fwd_tc(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo,
                                     const float* mv_pad, int B, int Hc, int Wc, float mneg, double* partials,
                                     uint32_t* bitsR, uint32_t* bitsC, float* dbgS, void* stream) {
  return ssp_desc_dense_fwd_tc_ex(Ahi, Alo, Bhi, Blo, mv_pad, B, Hc, Wc, mneg, partials, bitsR, bitsC, dbgS, 0, stream);
}

extern "C" int ssp_desc_dense_fwd_tc_ex(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo,
                                        const float* mv_pad, int B, int Hc, int Wc, float mneg, double* partials,
                                        uint32_t* bitsR, uint32_t* bitsC, float* dbgS, int flags, void* stream) {
  SSP_REQUIRE(Ahi && Bhi && mv_pad && partials, "ssp_desc_dense_fwd_tc: null pointer");
  SSP_REQUIRE((Alo == nullptr) == (Blo == nullptr), "ssp_desc_dense_fwd_tc: lo planes must both be given or both null");
  SSP_REQUIRE((bitsR == nullptr) == (bitsC == nullptr), "ssp_desc_dense_fwd_tc: bitsR/bitsC must both be given or both null");
  SSP_REQUIRE(B > 0 && Hc > 0 && Wc > 0, "ssp_desc_dense_fwd_tc: bad sizes");
  SSP_REQUIRE(mneg > 0.f, "ssp_desc_dense_fwd_tc: margin_neg must be > 0 (zero padding relies on it)");
  SSP_REQUIRE((((uintptr_t)Ahi | (uintptr_t)Bhi | (uintptr_t)Alo | (uintptr_t)Blo | (uintptr_t)mv_pad) & 15) == 0,
              "ssp_desc_dense_fwd_tc: operands must be 16-byte aligned");
  DescGeom g;
  static const bool epi2 = [] { const char* e = getenv("SSP_FWD_EPI"); return e && e[0] == '2'; }();
  SSP_REQUIRE(flags == 0 || (flags == 1 && epi2), "ssp_desc_dense_fwd_tc_ex: flags=%d needs the EPI2 epilogue (SSP_FWD_EPI=2)", flags);
  g.B = B; g.Hc = Hc; g.Wc = Wc; g.Nc = Hc * Wc; g.Nc_pad = desc_nc_pad(g.Nc); g.Dch = KD;
  g.cell = flags;  // this kernel has no use for the cell size: the field carries the flags
  g.dist = 0.f; g.lamda = 0.f; g.mpos = 0.f; g.mneg = mneg;
  uint64_t rows = (uint64_t)B * g.Nc_pad;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_plane_map(&mAh, Ahi, rows, BM))) return rc;
  if ((rc = make_plane_map(&mBh, Bhi, rows, BN / 2))) return rc;  // half boxes: each CTA of a cluster fetches one half
  if ((rc = make_plane_map(&mAl, Alo ? Alo : Ahi, rows, BM))) return rc;
  if ((rc = make_plane_map(&mBl, Blo ? Blo : Bhi, rows, BN / 2))) return rc;
  // persistent: one 2-CTA cluster per SM pair (or fewer when there is less work)
  long long items = (long long)B * (g.Nc_pad / (2 * BM)) * (g.Nc_pad / BN);
  int nclusters = (int)std::min<long long>(items, std::max(1, ssp_num_sms() / 2));
  int grid = 2 * nclusters;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FWD(PP, BB, EE)                                                                                   \
  do {                                                                                                           \
    if ((rc = set_smem(desc_dense_fwd_tc_kernel<PP, BB, EE>, FwdCfg<PP>::SMEM))) return rc;                      \
    if ((rc = launch_cluster2(desc_dense_fwd_tc_kernel<PP, BB, EE>, grid, FWD_THREADS, FwdCfg<PP>::SMEM, st, mAh, mAl, \
                              mBh, mBl, mv_pad, g, partials, bitsR, bitsC, dbgS))) return rc;                      \
  } while (0)
#define LAUNCH_FWD2(PP, BB) do { if (epi2) LAUNCH_FWD(PP, BB, true); else LAUNCH_FWD(PP, BB, false); } while (0)
  if (Alo) { if (bitsR) LAUNCH_FWD2(2, true); else LAUNCH_FWD2(2, false); }
  else     { if (bitsR) LAUNCH_FWD2(1, true); else LAUNCH_FWD2(1, false); }
#undef LAUNCH_FWD2
#undef LAUNCH_FWD
  SSP_CUDA_CHECK_LAUNCH("desc_dense_fwd_tc_kernel");
  return SSP_OK;
}

// out[b, d, r] = rowscale[b, r] * sum_k bit(r, k) * (Bhi + Blo)[b, k, d]      (out is [B, 256, Nc] fp32)
//                + sum_n pcoef[b, r, n] * possrc[b, d, plist[b, r, n]]   (sparse positive pairs; plist may be NULL)
// The positive-pair source is either fp32 NCHW (possrc) or the packed planes pos_hi (+ pos_lo) [B, Nc_pad, 256] bf16.
static int bits_gemm_tc_launch(const uint32_t* bits, const void* Bhi, const void* Blo, const float* rowscale,
                               const int* plist, const float* pcoef, const float* possrc, const void* pos_hi,
                               const void* pos_lo, int B, int Nc, float* out, void* stream) {
  SSP_REQUIRE(bits && Bhi && out, "ssp_desc_bits_gemm_tc: null pointer");
  SSP_REQUIRE(!plist || (pcoef && (possrc || pos_hi)), "ssp_desc_bits_gemm_tc: plist needs pcoef and a positive-pair source");
  SSP_REQUIRE(B > 0 && Nc > 0, "ssp_desc_bits_gemm_tc: bad sizes");
  SSP_REQUIRE((((uintptr_t)Bhi | (uintptr_t)Blo | (uintptr_t)pos_hi | (uintptr_t)pos_lo) & 15) == 0,
              "ssp_desc_bits_gemm_tc: operands must be 16-byte aligned");
  int Nc_pad = desc_nc_pad(Nc);
  uint64_t rows = (uint64_t)B * Nc_pad;
  CUtensorMap mh, ml;
  int rc;
  if ((rc = make_plane_map(&mh, Bhi, rows, KT / 2))) return rc;  // half boxes (cluster multicast)
  if ((rc = make_plane_map(&ml, Blo ? Blo : Bhi, rows, KT / 2))) return rc;
  long long items = (long long)B * (Nc_pad / (2 * BM)) * 2;
  int nclusters = (int)std::min<long long>(items, std::max(1, ssp_num_sms() / 2));
  int grid = 2 * nclusters;
  cudaStream_t st = (cudaStream_t)stream;
  // ring depth: SSP_BG_NS=deep selects the deeper stage ring (6 x 32 KB for the split engine, 8 x 16 KB single pass)
  static const bool deep = [] { const char* e = getenv("SSP_BG_NS"); return e && e[0] == 'd'; }();
  // round-robin items by default (measured: 87 -> 77 us per launch at B=32, DRAM re-reads of the B planes gone);
  // SSP_BG_SCHED=contiguous restores the contiguous ranges.  The deeper ring measured slower (96 us) and stays opt-in.
  static const int sched = [] { const char* e = getenv("SSP_BG_SCHED"); return (e && e[0] == 'c') ? 0 : 1; }();
  static const bool deepbits = [] { const char* e = getenv("SSP_BG_BITS"); return e && e[0] == 'd'; }();
  static const bool latepos = [] { const char* e = getenv("SSP_BG_POS"); return e && e[0] == 'l'; }();
#define LAUNCH_BG(PP, NN)                                                                    \
  do {                                                                                       \
    if (deepbits) { if (latepos) LAUNCH_BG3(PP, NN, true, true); else LAUNCH_BG3(PP, NN, true, false); }   \
    else          { if (latepos) LAUNCH_BG3(PP, NN, false, true); else LAUNCH_BG3(PP, NN, false, false); } \
  } while (0)
#define LAUNCH_BG3(PP, NN, DD, LL)                                                                                    \
  do {                                                                                                                \
    if ((rc = set_smem(desc_bits_gemm_tc_kernel<PP, NN, DD, LL>, BgCfg<PP, NN>::SMEM))) return rc;                    \
    if ((rc = launch_cluster2(desc_bits_gemm_tc_kernel<PP, NN, DD, LL>, grid, BG_THREADS, BgCfg<PP, NN>::SMEM, st, mh, ml, bits, \
                              rowscale, plist, pcoef, possrc, (const uint4*)pos_hi, (const uint4*)pos_lo, B, Nc,      \
                              Nc_pad, sched, out))) return rc;                                                        \
  } while (0)
  if (Blo) { if (deep) LAUNCH_BG(2, 6); else LAUNCH_BG(2, 4); }
  else     { if (deep) LAUNCH_BG(1, 8); else LAUNCH_BG(1, 6); }
#undef LAUNCH_BG
#undef LAUNCH_BG3
  SSP_CUDA_CHECK_LAUNCH("desc_bits_gemm_tc_kernel");
  return SSP_OK;
}

extern "C" int ssp_desc_bits_gemm_tc(const uint32_t* bits, const void* Bhi, const void* Blo, const float* rowscale,
                                     const int* plist, const float* pcoef, const float* possrc, int B, int Nc,
                                     float* out, void* stream) {
  return bits_gemm_tc_launch(bits, Bhi, Blo, rowscale, plist, pcoef, possrc, nullptr, nullptr, B, Nc, out, stream);
}

// Same GEMM with the positive-pair partners read from packed planes (pos_lo may be NULL: single-pass bf16 engine).
extern "C" int ssp_desc_bits_gemm_tc_planes(const uint32_t* bits, const void* Bhi, const void* Blo, const float* rowscale,
                                            const int* plist, const float* pcoef, const void* pos_hi, const void* pos_lo,
                                            int B, int Nc, float* out, void* stream) {
  SSP_REQUIRE(pos_hi, "ssp_desc_bits_gemm_tc_planes: null pointer");
  return bits_gemm_tc_launch(bits, Bhi, Blo, rowscale, plist, pcoef, nullptr, pos_hi, pos_lo, B, Nc, out, stream);
}
