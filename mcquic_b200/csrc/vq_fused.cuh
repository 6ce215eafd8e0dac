// Fused multi-codebook VQ on tcgen05: distance + first-index argmin + soft logits + code histogram in ONE launch.
//
// Reference: mcquic/modules/quantizer.py:153-179 (_distance), :144-150 (encode), :181-183,204 (logits).
// The reference materialises the [n, m, h, w, k] distance tensor three times; here every (128 points x 128 codewords)
// block of x.c_k lives in TMEM only, and the single HBM-sized stream is the logits output (when asked for).
//
//   tile    = 128 consecutive latent points of one codebook m  (persistent CTAs, static round-robin over tiles)
//   A       = [x_hi | x_lo]  fp16, 2d elements per row: the fp32 latents are read once with coalesced 32 B loads by
//             4 producer warps, split into hi + lo/2048, and written straight into the SWIZZLE_128B K-major layout
//             tcgen05.mma consumes (no workspace, no prep launch); |x|^2 falls out of the same pass
//   B       = [c_lo | c_hi]  fp16 (codebook packed once on the host, c * 2^e), streamed by TMA in 128-codeword chunks
//   MMA     = D_hh = x_hi.c_hi  (K-steps [0, d/16) of A against K-steps [d/16, 2d/16) of B)
//             D_lo = x_hi.c_lo + x_lo.c_hi  (all 2d/16 K-steps, same offsets in A and B)   -> 3 fp16 passes, fp32-grade
//   TMEM    = 2 x (128 + 128) columns: chunk c+1 is multiplied while chunk c is drained
//   drain   = 16 warps (4 lane quarters x 4 column quarters): tcgen05.ld -> dist = (|x|^2 + |c|^2) - 2 x.c in the
//             reference's operation order -> running (distance, index) minimum (strict <: first index wins ties) ->
//             logits = dist * (-temperature / sqrt(k)) -> swizzled smem staging -> TMA store of a 32 x 32 fp32 box
//   |c|^2   = 512 B per chunk, bulk-copied (cp.async.bulk) into an 8-slot shared-memory ring with the codebook stage
//   finish  = the four column quarters are merged through shared memory; int64 codes and the histogram are written.
//
// d = 128 (qp = 1: one codebook over all 128 channels): an operand row is 256 fp16 = four 64-element chunks, 64 KB per
// 128-row buffer -- two A buffers and two such codebook stages would not fit.  There the latents are single-buffered
// (a tile lasts k / 128 = 4..64 chunks, the refill bubble is small) and the codebook is streamed in HALF rows: stage
// 2c holds the c_lo halves of chunk c (K-steps [0, d/16): the x_hi.c_lo part of D_lo), stage 2c+1 the c_hi halves
// (D_hh = x_hi.c_hi and the x_lo.c_hi part of D_lo) -- 32 KB stages like d = 64's.
//
// Supported: d in {32, 64, 128}, k % 128 == 0, (h*w) % 32 == 0 or 32 % (h*w) == 0; everything else -> vq_assign_kernel.
#pragma once
#include "conv_tc.cuh"

namespace mcq {

constexpr int VQF_BM = 128;
constexpr int VQF_BN = 128;
constexpr int VQF_PROD_WARPS = 4;
constexpr int VQF_EPI_WARPS = 16;                                                // 4 TMEM lane quarters x 4 column quarters
constexpr int VQF_FIRST_PROD = 2;
constexpr int VQF_FIRST_EPI = VQF_FIRST_PROD + VQF_PROD_WARPS;                 // 6 (6 % 4 == 2: quarters 2,3,0,1)
constexpr int VQF_THREADS = 32 * (VQF_FIRST_EPI + VQF_EPI_WARPS);               // 704
constexpr int VQF_CHUNK_BYTES = VQF_BM * 128;                                    // one 64-element K chunk of 128 rows
constexpr int VQF_STAGE_BYTES = 4096;                                            // 32 rows x 32 fp32
constexpr int VQF_C2_SLOTS = 8;   // ring of |c_k|^2 chunks (128 floats); > codebook stages + TMEM buffers + 1, see below

struct VqFusedArgs {
  const float* x;            // [P, m*d] NHWC latents
  const float* c2;           // [m, k]
  long long* codes;          // [n, m, hw]
  int* hist;                 // [m, k] or null
  const float* logit_scale;  // [m] or null
  int P, hw, m, k, d;
  int tiles_p;               // ceil(P / 128)
  int nb;                    // codebook stages
  int has_logits;
  float cb_scale;            // 2^-e of the packed codebook
  float inv_sqrt_k;
  int hist_on;
  int nst;                   // logits staging buffers per drain warp (2, or 1 when shared memory is short)
  int na;                    // latent (A) buffers: 2, or 1 for d = 128
  int halves;                // codebook stages per 128-codeword chunk: 1, or 2 (half rows) for d = 128
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// (fence_async_smem, bulk_commit, bulk_wait_read, bulk_wait_all: conv_tc.cuh)
// Wait for warps that are off the critical path (codebook stream, MMA issue, latent producer): between polls the warp
// sleeps, so its spin loop does not take issue slots from the drain warps sharing its scheduler (ncu: 45 % of all
// executed instructions were wait loops before this).  Same 4 s watchdog as mbar_wait.
__device__ __noinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, int code, unsigned ns) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  for (unsigned i = 1;; ++i) {
    __nanosleep(ns);
    if (mbar_try_wait(bar, parity)) return;
    if ((i & 1023u) == 0 && global_ns() - t0 > TC_WATCHDOG_NS) {
      atomicExch(&g_watchdog_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool LOGITS>
__global__ void __launch_bounds__(VQF_THREADS, 1)
vq_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmL, const VqFusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = a.d, nb = a.nb;
  const int kch = (2 * d) / TC_BK;                       // 64-element K chunks per operand row (1 or 2)
  const uint32_t op_bytes = (uint32_t)kch * VQF_CHUNK_BYTES;   // one A buffer (= one B stage when halves == 1)
  const int na = a.na, halves = a.halves;
  const int kch_st = kch / halves;                             // 64-element chunks per codebook stage
  const uint32_t bst_bytes = op_bytes / (uint32_t)halves;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)na * op_bytes;
  const uint32_t st_base = b_base + (uint32_t)nb * bst_bytes;                      // logits staging (1024-aligned)
  const uint32_t st_bytes = LOGITS ? (uint32_t)(VQF_EPI_WARPS * a.nst * VQF_STAGE_BYTES) : 0u;
  const uint32_t x2_off = (st_base - smem_base) + st_bytes;                        // float [2][128]
  const uint32_t red_off = x2_off + 2u * VQF_BM * 4u;                              // u64 [2][3][128]
  const uint32_t c2_off = red_off + 6u * VQF_BM * 8u;                              // float [VQF_C2_SLOTS][128]
  const uint32_t bar_base = smem_base + c2_off + (uint32_t)VQF_C2_SLOTS * VQF_BN * 4u;
  const uint32_t c2_base = smem_base + c2_off;
  float* x2buf = reinterpret_cast<float*>(smem_gen + x2_off);
  unsigned long long* red = reinterpret_cast<unsigned long long*>(smem_gen + red_off);
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto tfull = [&](int i) { return bar_base + 8u * (4 + i); };
  auto tempty = [&](int i) { return bar_base + 8u * (6 + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (8 + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (8 + nb + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8u * (8 + 2 * nb));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (LOGITS) tma_prefetch_desc(&tmL);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full(i), VQF_PROD_WARPS * 32);
      mbar_init(a_empty(i), 1 + VQF_EPI_WARPS);
      mbar_init(tfull(i), 1);
      mbar_init(tempty(i), VQF_EPI_WARPS);
    }
    for (int i = 0; i < nb; ++i) {
      mbar_init(b_full(i), 1);
      mbar_init(b_empty(i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  const int total_tiles = a.tiles_p * a.m;
  const int chunks = a.k / VQF_BN;

  if (warp == 0) {
    // ===================== codebook stream (TMA) =====================
    int gc = 0, gs = 0;                                     // chunk / stage counters
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mi = t / a.tiles_p;
      for (int c = 0; c < chunks; ++c, ++gc) {
        // |c_k|^2 of the chunk rides on the barrier of the chunk's first codebook stage but lives in its own ring: the
        // drain warps read it after the MMA has released the stage.  Slot gc % 8 was last read by the drain of chunk
        // gc - 8; this point is only reached after the MMAs of stage gs - nb completed (chunk gc - nb / halves at the
        // latest), hence after drain(gc - nb - 2) handed its accumulator back (its |c|^2 reads precede that hand-over),
        // and nb + 2 <= 6 < 8.
        const int cs = gc % VQF_C2_SLOTS;
        for (int hf = 0; hf < halves; ++hf, ++gs) {
          const int s = gs % nb;
          mbar_wait_sleep(b_empty(s), (((uint32_t)(gs / nb)) & 1u) ^ 1u, 31, 200);
          if (elect_one()) {
            mbar_expect_tx(b_full(s), bst_bytes + (hf == 0 ? VQF_BN * 4u : 0u));
            for (int kc = 0; kc < kch_st; ++kc)
              tma_load_2d(&tmB, b_base + (uint32_t)s * bst_bytes + (uint32_t)kc * VQF_CHUNK_BYTES, b_full(s),
                          (hf * kch_st + kc) * TC_BK, mi * a.k + c * VQF_BN);
            if (hf == 0)
              bulk_load_1d(c2_base + (uint32_t)cs * VQF_BN * 4u, a.c2 + (size_t)mi * a.k + c * VQF_BN, VQF_BN * 4u,
                           b_full(s));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(VQF_BN >> 3) << 17) | ((uint32_t)(VQF_BM >> 4) << 24);
    const int ks = d / 16;                                 // K-steps per half row
    int gc = 0, gs = 0, i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i % na;
      mbar_wait_sleep(a_full(ab), ((uint32_t)(i / na)) & 1u, 32, 100);
      const uint64_t a0 = make_sdesc(a_base + (uint32_t)ab * op_bytes);
      for (int c = 0; c < chunks; ++c, ++gc) {
        const int buf = gc & 1;
        mbar_wait_sleep(tempty(buf), (((uint32_t)(gc >> 1)) & 1u) ^ 1u, 33, 60);
        const uint32_t d_hh = tmem_base + (uint32_t)(buf * 2 * VQF_BN);
        const uint32_t d_lo = d_hh + (uint32_t)VQF_BN;
        // K-step j of a row lives in chunk j/4 at +32 B * (j%4); descriptor addresses are in 16 B units
        auto off = [](int j) { return (uint64_t)((j >> 2) * (VQF_CHUNK_BYTES >> 4) + (j & 3) * 2); };
        for (int hf = 0; hf < halves; ++hf, ++gs) {
          const int s = gs % nb;
          mbar_wait_sleep(b_full(s), ((uint32_t)(gs / nb)) & 1u, 34, 60);
          tc_fence_after();
          const uint64_t b0 = make_sdesc(b_base + (uint32_t)s * bst_bytes);
          if (elect_one()) {
            if (halves == 1) {
              for (int j = 0; j < ks; ++j) umma_f16(d_hh, a0 + off(j), b0 + off(ks + j), idesc, j > 0 ? 1u : 0u);
              for (int j = 0; j < 2 * ks; ++j) umma_f16(d_lo, a0 + off(j), b0 + off(j), idesc, j > 0 ? 1u : 0u);
            } else if (hf == 0) {       // stage = [c_lo] of the chunk: D_lo = x_hi . c_lo
              for (int j = 0; j < ks; ++j) umma_f16(d_lo, a0 + off(j), b0 + off(j), idesc, j > 0 ? 1u : 0u);
            } else {                    // stage = [c_hi]: D_hh = x_hi . c_hi, D_lo += x_lo . c_hi
              for (int j = 0; j < ks; ++j) umma_f16(d_hh, a0 + off(j), b0 + off(j), idesc, j > 0 ? 1u : 0u);
              for (int j = 0; j < ks; ++j) umma_f16(d_lo, a0 + off(ks + j), b0 + off(j), idesc, 1u);
            }
            umma_commit(b_empty(s));
            if (hf == halves - 1) {
              umma_commit(tfull(buf));
              if (c == chunks - 1) umma_commit(a_empty(ab));
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < VQF_FIRST_EPI) {
    // ===================== latent producer: fp32 -> split fp16, swizzled K-major smem; |x|^2 =====================
    const int ptid = threadIdx.x - VQF_FIRST_PROD * 32;
    const int g = d / 8;                                   // threads per row (8 floats each): 4 or 8
    const int rows_per_pass = (VQF_PROD_WARPS * 32) / g;
    const int C = a.m * d;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i % na;
      const int mi = t / a.tiles_p, p0 = (t - mi * a.tiles_p) * VQF_BM;
      mbar_wait_sleep(a_empty(ab), (((uint32_t)(i / na)) & 1u) ^ 1u, 35, 1000);
      const uint32_t abuf = a_base + (uint32_t)ab * op_bytes;
      for (int r0 = 0; r0 < VQF_BM; r0 += rows_per_pass) {
        const int row = r0 + ptid / g, j = ptid % g;
        const int pnt = p0 + row;
        float v[8];
        if (pnt < a.P) {
          load_f32v<8>(a.x, (size_t)pnt * C + (size_t)mi * d + j * 8, v);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(v[e], v[e], s);
        for (int o = 1; o < g; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f32x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        const uint32_t rbase = abuf + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        const int e_hi = j * 8, e_lo = d + j * 8;          // element index of this thread's 8 values in the 2d-wide row
        const uint32_t ad_hi = rbase + (uint32_t)(e_hi >> 6) * VQF_CHUNK_BYTES +
                               ((((uint32_t)(e_hi & 63) >> 3) ^ (uint32_t)(row & 7)) << 4);
        const uint32_t ad_lo = rbase + (uint32_t)(e_lo >> 6) * VQF_CHUNK_BYTES +
                               ((((uint32_t)(e_lo & 63) >> 3) ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad_hi), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                     "r"(hi[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad_lo), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]),
                     "r"(lo[3])
                     : "memory");
        if (j == 0) x2buf[ab * VQF_BM + row] = s;
      }
      fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(a_full(ab));
    }
  } else {
    // ===================== drain: distance, argmin, logits =====================
    const int q = warp & 3;
    const int ew = warp - VQF_FIRST_EPI;
    const int cg = ew >> 2;                                // column quarter of every chunk: columns [32 cg, 32 cg + 32)
    const int row = q * 32 + lane;
    const float neg2s = -2.0f * a.cb_scale;
    const int per = a.hw >= 32 ? a.hw / 32 : 1;            // 32-row blocks per image (hw >= 32)
    const int ib = a.hw >= 32 ? 1 : 32 / a.hw;             // images per 32-row block (hw < 32)
    const int col0 = cg * 32;
    // loop-invariant addresses.  Staging row `lane`, 16 B chunk j is stored at chunk j ^ (lane & 7) (the 128B swizzle
    // the store tensor map expects); with the base below that is simply base ^ (j << 4).
    const uint32_t stage0 = st_base + (uint32_t)ew * (uint32_t)(a.nst * VQF_STAGE_BYTES);
    const uint32_t stage_flip = a.nst == 2 ? (uint32_t)VQF_STAGE_BYTES : 0u;
    const uint32_t wlane = (uint32_t)lane * 128u + ((uint32_t)(lane & 7) << 4);
    const uint32_t t_acc0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
    const uint32_t c2_mine = c2_base + (uint32_t)col0 * 4u;
    const uint32_t tfull0 = tfull(0), tempty0 = tempty(0);
    uint32_t gc = 0;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i % na;
      const int mi = t / a.tiles_p, p0 = (t - mi * a.tiles_p) * VQF_BM;
      mbar_wait(a_full(ab), ((uint32_t)(i / na)) & 1u, 36);
      const float x2 = x2buf[ab * VQF_BM + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(a_empty(ab));
      const float lscale = -a.inv_sqrt_k * (a.logit_scale ? a.logit_scale[mi] : 1.0f);
      const int rb = (p0 + q * 32) >> 5;                   // global 32-row block of this warp
      const int nn0 = a.hw >= 32 ? rb / per : rb * ib;
      const int pix0 = a.hw >= 32 ? (rb - nn0 * per) * 32 : 0;
      const bool blk_live = (p0 + q * 32) < a.P;
      float best = INFINITY;
      int best_k = 0x7fffffff;
      int kcol = col0;                                     // codeword index of this warp's first column in the chunk
      for (int c = 0; c < chunks; ++c, ++gc, kcol += VQF_BN) {
        const uint32_t buf = gc & 1u;
        const uint32_t sbuf = stage0 + buf * stage_flip;
        const uint32_t c2a = c2_mine + (gc & (uint32_t)(VQF_C2_SLOTS - 1)) * (uint32_t)(VQF_BN * 4);
        mbar_wait(tfull0 + 8u * buf, (gc >> 1) & 1u, 37);
        tc_fence_after();
        const uint32_t t_acc = t_acc0 + buf * (uint32_t)(2 * VQF_BN);
        if (LOGITS) {
          // the TMA store that last read this staging buffer must have drained it
          if (lane == 0) { if (a.nst == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
          __syncwarp();
        }
        const uint32_t wb = sbuf + wlane;
        int best_e = -1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hh[16], ll[16];
          tmem_ld16(t_acc + (uint32_t)(half * 16), hh);
          tmem_ld16(t_acc + (uint32_t)(VQF_BN + half * 16), ll);
          float cc[16];                                     // |c_k|^2: broadcast reads from the shared-memory ring
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(cc[e]), "=f"(cc[e + 1]), "=f"(cc[e + 2]), "=f"(cc[e + 3])
                         : "r"(c2a + (uint32_t)((half * 16 + e) * 4)));
          tmem_ld_wait();
          if (half == 1) {                                  // accumulator fully read: hand the buffer back to the MMA
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * buf);
          }
#pragma unroll
          for (int e4 = 0; e4 < 16; e4 += 4) {
            float lg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e4 + u;
              const float tt = fmaf(__uint_as_float(ll[e]), kLoInv, __uint_as_float(hh[e]));
              const float dist = fmaf(tt, neg2s, x2 + cc[e]);    // (|x|^2 + |c|^2) - 2 x.c  (quantizer.py:176)
              if (dist < best) { best = dist; best_e = half * 16 + e; }   // strict <: the first index wins ties
              if (LOGITS) lg[u] = dist * lscale;
            }
            if (LOGITS)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                           ::"r"(wb ^ (uint32_t)((half * 4 + (e4 >> 2)) << 4)),
                           "f"(lg[0]), "f"(lg[1]), "f"(lg[2]), "f"(lg[3])
                           : "memory");
          }
        }
        if (best_e >= 0) best_k = kcol + best_e;
        if (LOGITS) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (blk_live) tma_store_4d(&tmL, sbuf, kcol, pix0, mi, nn0);
            bulk_commit();
          }
        }
      }
      // merge the four column quarters: (distance, index) lexicographic minimum = first index on ties
      unsigned long long key = ((unsigned long long)ordered_f32(best) << 32) | (unsigned long long)(uint32_t)best_k;
      unsigned long long* rbuf = red + (size_t)(i & 1) * 3 * VQF_BM;
      if (cg > 0) rbuf[(cg - 1) * VQF_BM + row] = key;
      named_bar(1 + q, 128);
      if (cg == 0) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const unsigned long long other = rbuf[o * VQF_BM + row];
          if (other < key) key = other;
        }
        const int pnt = p0 + row;
        if (pnt < a.P) {
          int code = (int)(key & 0xFFFFFFFFull);
          if (code < 0 || code >= a.k) code = 0;           // all-NaN row
          const int nn = pnt / a.hw, pix = pnt - nn * a.hw;
          a.codes[((size_t)nn * a.m + mi) * a.hw + pix] = code;
          if (a.hist_on) atomicAdd(a.hist + (size_t)mi * a.k + code, 1);
        }
      }
    }
    if (LOGITS && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

}  // namespace mcq
