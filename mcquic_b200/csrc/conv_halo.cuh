// tcgen05 3x3 stride-1 convolution with on-chip tap reuse ("halo" kernel).
//
// The per-tap kernel (conv_tc.cuh) re-reads the activation tile from L2 once per filter tap and streams the whole
// weight matrix per 128-pixel tile; at N=64 both the 1-pass and the 3-pass variant sit on the L2->SM bandwidth
// ceiling (~10 TB/s measured).  This kernel removes most of that traffic:
//
// * A: per 64-channel chunk ONE TMA box brings the (8+2) x (16+2) pixel halo of an 8x16 output tile into shared
//   memory ([halo row][halo col][64 ch], 128 B per pixel, SWIZZLE_128B).  All nine taps are then served from that
//   copy: the UMMA shared-memory descriptor of tap (r, s) simply starts (r*pitch + s) pixels further in, with the
//   8-row core-matrix groups (= 8 consecutive pixels of one tile row) one halo row (pitch*128 B) apart.
//   The 128B swizzle is a function of the absolute shared-memory address, so a start address that is 128 B- but
//   not 1024 B-aligned addresses the data exactly where TMA put it.     A traffic: 9x -> 1.4x.
// * B: the CL CTAs of a thread-block cluster work on CL neighbouring pixel tiles with the same weights; each
//   loads 1/CL of every weight stage and multicasts it to all of them.  B traffic: 1/CL.
//
// Roles and TMEM double buffering as in conv_tc.cuh.
#pragma once
#include <cooperative_groups.h>

#include "conv_tc.cuh"

namespace mcq {

constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_ROWS = HALO_TH + 2;

struct HaloArgs {
  int pitch;        // halo pixels per halo row in smem (10 = dense)
  int box_w;        // TMA box width (= pitch)
  int a_bytes;      // bytes of one halo plane (1024-aligned)
  int na, nbs;      // A buffers, B stages
  int tps;          // filter taps per weight stage (1 or 3): amortises the per-stage barrier round trip of the
                    // single MMA-issuing thread when a tap is only 4 MMAs (1-pass)
  int base_mode;    // 0: descriptor base_offset = 0; 1: base_offset = (start >> 7) & 7
  int tiles_m;      // pixel tiles (n * tiles_y * tiles_x)
  int groups_m;     // ceil(tiles_m / CL)
  int resident;     // the whole weight matrix of the N tile stays in shared memory (nbs = all stages): pair kernel 1-pass,
                    // halo kernel whenever a single N tile's stages all fit
  int per_ct;       // resident mode with several N tiles: clusters per N tile (cluster c serves N tile c % tiles_c only)
};

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t make_sdesc_halo(uint32_t saddr, uint32_t sbo_bytes, int base_mode) {
  const uint32_t lo = (saddr >> 4) & 0x3FFF;
  uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
  if (base_mode == 1) hi |= ((saddr >> 7) & 7u) << 17;   // matrix base offset, bits [49,52)
  return ((uint64_t)hi << 32) | lo;
}

template <int PASSES, int CL, int DRAIN = DRAIN_ROWS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const ConvArgs p, const HaloArgs hp, const __grid_constant__ OutMaps om) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = p.bn;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  constexpr int NP = (PASSES == 3) ? 2 : 1;              // planes per operand
  const uint32_t a_buf_bytes = (uint32_t)hp.a_bytes * NP;
  const uint32_t b_plane = (uint32_t)bn * TC_BK * 2;
  const uint32_t b_tap_bytes = b_plane * NP;             // one tap: [B_hi | B_lo]
  const uint32_t b_stage_bytes = b_tap_bytes * hp.tps;
  const uint32_t b_slice = b_plane / CL;                 // bytes of one CTA's share of a weight plane
  const int na = hp.na, nbs = hp.nbs;

  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + a_buf_bytes * na;
  const uint32_t bar_base = b_base + b_stage_bytes * nbs;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (na + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * na + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * na + nbs + i); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * na + 2 * nbs + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * na + 2 * nbs + 2 + b); };
  uint32_t* tmem_slot =
      reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8u * (2 * na + 2 * nbs + 4));
  const uint32_t epi_base = (bar_base + 8u * (2 * na + 2 * nbs + 4) + 16u + 511u) & ~511u;   // 512 B: swizzle period of the staging
  float* bias_smem = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base) + TC_EPI_WARPS * TC_EPI_STAGE_BYTES);
  const bool bias_staged = p.cout <= TC_BIAS_SMEM_FLOATS;
  if (bias_staged)
    for (int i = threadIdx.x; i < p.cout; i += TC_THREADS) bias_smem[i] = p.bias[i];

  const int acc_cols = NP * bn;
  const int nbuf = (2 * acc_cols <= (int)TC_TMEM_COLS) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (PASSES == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    for (int i = 0; i < na; ++i) {
      mbar_init(a_full(i), 1);
      mbar_init(a_empty(i), 1);
    }
    for (int i = 0; i < nbs; ++i) {
      mbar_init(b_full(i), 1);
      mbar_init(b_empty(i), CL);   // every CTA of the cluster must have consumed the stage before it is refilled
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), TC_EPI_WARPS);
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
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias staging
  // of constant weights) overlapped the tail of the previous kernel in the stream; from here on we read its outputs.
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  // cin need not fill its last 64-channel K chunk: TMA zero-fills the A box beyond the tensor's channel extent, and
  // whatever the weight box holds there (the next tap's columns, or zeros past the last one) is multiplied by those zeros
  const int kchunks = (p.cin + TC_BK - 1) / TC_BK;
  const int total_work = hp.groups_m * p.tiles_c;        // one work item = CL neighbouring pixel tiles x one N tile
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;

  auto decode_tile = [&](int w, int& ct, int& x0, int& y0, int& n) {
    ct = w / hp.groups_m;
    int mt = (w - ct * hp.groups_m) * CL + (int)crank;   // may be >= tiles_m (phantom tile: all stores masked)
    const int bx = mt % p.tiles_x;
    mt /= p.tiles_x;
    const int by = mt % p.tiles_y;
    n = mt / p.tiles_y;
    x0 = bx * HALO_TW;
    y0 = by * HALO_TH;
  };

  if (warp == 0) {
    {
      // (warp-uniform loop; one elected lane issues the TMA, see elect_one)
      // The halo of chunk g+na-1 is requested while the weights of chunk g stream: late enough that its buffer is
      // certainly free (the MMA has passed tap t* - nbs >= 0 of chunk g, so chunk g-1 is done -> the wait below
      // never blocks), early enough (>= one whole chunk of MMA time) to hide the TMA latency of the 23 KB box.
      const int my_work = (total_work - cluster_id + num_clusters - 1) / num_clusters;   // work items of this cluster
      const int total_chunks = my_work * kchunks;
      const int spc = 9 / hp.tps;                          // weight stages per chunk
      const int t_star = nbs < spc - 1 ? nbs : spc - 1;
      int a_issue = 0, ab = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto issue_a = [&]() {
        if (a_issue >= total_chunks) return;
        const int wi = cluster_id + (a_issue / kchunks) * num_clusters, kc = a_issue % kchunks;
        int ct, x0, y0, n;
        decode_tile(wi, ct, x0, y0, n);
        mbar_wait(a_empty(ab), aph ^ 1u, 11, p.wait_sleep_ns);
        const uint32_t sa = a_base + a_buf_bytes * ab;
        if (elect_one()) {
          // the whole box is always transferred (out-of-image pixels are zero-filled): tx = box bytes
          mbar_expect_tx(a_full(ab), (uint32_t)(hp.box_w * HALO_ROWS * 128) * NP);
          tma_load_5d(&tmA_hi, sa, a_full(ab), kc * TC_BK, x0 - 1, 0, y0 - 1, n);
          if (PASSES == 3) tma_load_5d(&tmA_lo, sa + hp.a_bytes, a_full(ab), kc * TC_BK, x0 - 1, 0, y0 - 1, n);
        }
        __syncwarp();
        if (++ab == na) { ab = 0; aph ^= 1u; }
        ++a_issue;
      };
      for (int i = 0; i < na - 1; ++i) issue_a();
      // hp.resident (single N tile whose weight stages all fit: narrow layers such as the 128 -> 12 output convolution or
      // Neon's 32-channel nets): the ring holds every stage of the N tile, filled during the first work item and never
      // recycled -- no weight traffic and no empty-barrier round trip for any later tile
      bool load_b = true;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        int ct, x0, y0, n;
        decode_tile(w, ct, x0, y0, n);
        const int row0 = ct * bn + (int)crank * (bn / CL);
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int sg = 0; sg < spc; ++sg) {
            if (!hp.resident) mbar_wait(b_empty(bs), bph ^ 1u, 12, p.wait_sleep_ns);
            if (load_b && elect_one()) {
              mbar_expect_tx(b_full(bs), b_stage_bytes);
              for (int tt = 0; tt < hp.tps; ++tt) {
                const uint32_t sb = b_base + b_stage_bytes * bs + b_tap_bytes * tt;
                const int kb = (sg * hp.tps + tt) * p.cin + kc * TC_BK;
                if (CL > 1) {
                  tma_load_2d_mc(&tmB_hi, sb + crank * b_slice, b_full(bs), kb, row0, kMask);
                  if (PASSES == 3) tma_load_2d_mc(&tmB_lo, sb + b_plane + crank * b_slice, b_full(bs), kb, row0, kMask);
                } else {
                  tma_load_2d(&tmB_hi, sb, b_full(bs), kb, row0);
                  if (PASSES == 3) tma_load_2d(&tmB_lo, sb + b_plane, b_full(bs), kb, row0);
                }
              }
            }
            __syncwarp();
            if (++bs == nbs) { bs = 0; bph ^= 1u; }
            if (sg == t_star) issue_a();
          }
        }
        if (hp.resident) load_b = false;
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t sbo = (uint32_t)hp.pitch * 128u;
      const uint32_t idesc_wide = (1u << 4) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const bool wide = (PASSES == 3) && (2 * bn <= 256);
      const int spc = 9 / hp.tps;
      // halo offset of tap (r, s) in 16 B units: (r * pitch + s) pixels x 128 B (computed, not tabulated: a table indexed
      // by the runtime stage index would live in local memory, see conv_pair.cuh)
      auto tap_off16 = [&](int tap) {
        const int r = tap / 3;
        return (uint32_t)(r * hp.pitch + (tap - 3 * r)) * 8u;
      };
      int ab = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters, ++it) {
        const int buf = (nbuf == 2) ? (it & 1) : 0;
        const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u, 13);
        tc_fence_after();
        const uint32_t d_hh = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t d_lo = d_hh + (uint32_t)bn;
        uint32_t acc = 0;   // the first MMA of a tile overwrites the accumulator, all later ones add
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(a_full(ab), aph, 14);
          // descriptor arithmetic is hoisted out of the issue loop: a single thread feeds the tensor core, and every
          // integer instruction between two tcgen05.mma costs issue latency (a 1-pass MMA lasts only 64 cycles)
          const uint64_t a_hi0 = make_sdesc_halo(a_base + a_buf_bytes * ab, sbo, hp.base_mode);
          const uint64_t a_lo0 = a_hi0 + (uint64_t)(hp.a_bytes >> 4);
          for (int sg = 0; sg < spc; ++sg) {
            if (!hp.resident || it == 0) mbar_wait(b_full(bs), bph, 15);
            tc_fence_after();
            const uint64_t b0 = make_sdesc(b_base + b_stage_bytes * bs);
            if (elect_one()) {
            uint32_t acc_i = acc;
#pragma unroll 3
            for (int tt = 0; tt < hp.tps; ++tt) {
              const int tap = sg * hp.tps + tt;
              const uint64_t at = (uint64_t)tap_off16(tap);
              const uint64_t bt = b0 + (uint64_t)(tt * (int)(b_tap_bytes >> 4));
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);   // +32 B per K = 16, in 16 B units
                if (PASSES == 3) {
                  if (wide) {   // one N = 2*bn MMA over the adjacent [B_hi; B_lo] stage: A_hi is fetched once
                    umma_f16(d_hh, a_hi0 + at + ko, bt + ko, idesc_wide, acc_i);
                  } else {
                    umma_f16(d_hh, a_hi0 + at + ko, bt + ko, idesc, acc_i);
                    umma_f16(d_lo, a_hi0 + at + ko, bt + (uint64_t)(b_plane >> 4) + ko, idesc, acc_i);
                  }
                  umma_f16(d_lo, a_lo0 + at + ko, bt + ko, idesc, 1u);
                } else {
                  umma_f16(d_hh, a_hi0 + at + ko, bt + ko, idesc, acc_i);
                }
                acc_i = 1u;
              }
            }
            if (!hp.resident) {
              if (CL > 1) umma_commit_mc(b_empty(bs), kMask); else umma_commit(b_empty(bs));
            }
            if (sg == spc - 1) umma_commit(a_empty(ab));   // halo buffer free once the chunk's last tap has been read
            }
            __syncwarp();
            acc = 1u;
            if (++bs == nbs) { bs = 0; bph ^= 1u; }
          }
          if (++ab == na) { ab = 0; aph ^= 1u; }
        }
        if (elect_one()) umma_commit(tfull_bar(buf));
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const uint32_t stage = epi_base + (uint32_t)(warp - 2) * TC_EPI_STAGE_BYTES;
    auto tile_pix = [&](int w, int& ct) {
      int x0, y0, n0;
      decode_tile(w, ct, x0, y0, n0);
      return [=, &p](int row, int& n, int& oy, int& ox) {
        ox = x0 + (row & (HALO_TW - 1));
        oy = y0 + (row >> 3);
        n = n0;
        return (ox < p.wout) && (oy < p.hout) && (n < p.n);
      };
    };
    int it = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++it) {
      int ct;
      auto pix = tile_pix(w, ct);
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
      if (DRAIN != DRAIN_TMA) prefetch_epilogue_operands(p, bn, ct, cg, q, lane, pix);
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols);
      drain_tile<PASSES, false, DRAIN>(p, &om, t_acc, bn, ct, cg, q, lane, stage, pix, bias_staged ? bias_smem : p.bias,
                                       effective_w_scale(p), [&]() {
                                         mbar_wait(tfull_bar(buf), use & 1u, 16, p.wait_sleep_ns);
                                         tc_fence_after();
                                       });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    if constexpr (DRAIN == DRAIN_TMA) {
      if (lane == 0) bulk_wait_all();          // this lane's bulk stores have left shared memory and are complete
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA may leave while peers can still multicast into it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

}  // namespace mcq
