// tcgen05 3x3 stride-1 convolution, CTA-pair variant (tcgen05.mma.cta_group::2) of conv_halo.cuh.
//
// The halo kernel's limiter is the shared-memory operand fetch of the tensor core: a single-CTA M=128 x N=128 x K=16
// MMA reads 8 KB per 64 cycles (the whole 128 B/clk port), leaving nothing for TMA fills and the epilogue.  Here the
// two CTAs of a cluster (two SMs of one TPC) execute ONE MMA over M = 256 pixels (each CTA's own 8x16 pixel tile from
// its own halo copy) while the weight operand is split between them, so each SM fetches half of B:
//
//   1-pass:  D[256 x N] += A_hi * B_hi          B_hi rows [64r, 64r+64) live in CTA r            6 KB / 64 clk / SM
//   3-pass:  D[:, 0:2N]  += A_hi * [B_hi | B_lo]   CTA 0 holds all of B_hi, CTA 1 all of B_lo     (one N = 2*bn MMA)
//            D[:, N:2N]  += A_lo * B_hi            rows [64r, 64r+64) of B_hi again in CTA r     14 KB / 192 clk / SM
//            (accumulator columns [0,N) = hi*hi, [N,2N) = hi*lo + lo*hi, same layout as the single-CTA kernels)
//
// Weight-stationary mode (hp.resident, 1-pass with a single N tile): half of B_hi for ALL taps and chunks is 144 KB per
// CTA and fits next to two halo buffers, so it is loaded once per launch instead of once per pixel tile.  That removes
// 295 KB of shared-memory writes per tile -- with the operand fetch of the MMA already at 96 B/clk of the 128 B/clk port,
// the weight refills were what kept the 1-pass layers at ~55 % of the tensor peak.
//
// Synchronisation (per stage, barriers at identical offsets in both CTAs):
//   * every CTA's TMA producer loads its own operands but signals the LEADER's (rank 0) full barrier
//     (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped into CTA 0); only the leader arms
//     expect_tx, with the byte count of both CTAs;
//   * the leader's MMA thread issues for the pair and multicasts tcgen05.commit to both CTAs' empty / tmem-full barriers;
//   * the peer's epilogue warps arrive remotely on the leader's tmem-empty barrier (count 2 x 16 warps).
#pragma once
#include "conv_halo.cuh"

namespace mcq {

// shared::cluster address of the same shared-memory object in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  return ra;
}

__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* map, uint32_t dst, uint32_t leader_bar, int c0, int c1,
                                             int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
// Arrive on the same barrier in CTA `cta` of the cluster.  Default semantics (.release at CTA scope), NOT .release.cluster:
// the latter compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive -- every drain warp of the peer CTA then waits, once
// per tile, until all its outstanding global stores are visible device-wide (ncu source view: the hottest non-wait
// instructions of the peer's drain).  What the arrive has to order is tensor-memory traffic -- tcgen05.wait::ld before it in
// program order, tcgen05.fence::before_thread_sync / ::after_thread_sync around the barrier -- not global memory.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar), "r"(cta)
      : "memory");
}

template <int PASSES, bool GN = false, int DRAIN = DRAIN_ROWS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi_full, const __grid_constant__ CUtensorMap tmB_lo_full,
                 const __grid_constant__ CUtensorMap tmB_hi_half, const ConvArgs p, const HaloArgs hp,
                 const __grid_constant__ OutMaps om) {
  // tmB_{hi,lo}_full: weight planes with box {64, bn} (3-pass only); tmB_hi_half: hi plane with box {64, bn / 2}
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = p.bn;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  constexpr int NP = (PASSES == 3) ? 2 : 1;
  const uint32_t a_buf_bytes = (uint32_t)hp.a_bytes * NP;
  const uint32_t b_plane = (uint32_t)bn * TC_BK * 2;                 // bn rows x 128 B
  const uint32_t b_half = b_plane / 2;
  // per tap and CTA: 3-pass [Bw: bn rows (hi plane in CTA 0, lo plane in CTA 1) | B2: bn/2 rows of hi]; 1-pass [B2]
  const uint32_t b_tap_bytes = (PASSES == 3) ? (b_plane + b_half) : b_half;
  const uint32_t b_stage_bytes = b_tap_bytes * hp.tps;
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
    tma_prefetch_desc(&tmB_hi_half);
    if (PASSES == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(leader ? &tmB_hi_full : &tmB_lo_full);
    }
    for (int i = 0; i < na; ++i) {
      mbar_init(a_full(i), 1);
      mbar_init(a_empty(i), 1);
    }
    for (int i = 0; i < nbs; ++i) {
      mbar_init(b_full(i), 1);
      mbar_init(b_empty(i), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 2 * TC_EPI_WARPS);   // both CTAs' epilogue warps (only the leader's copy is waited on)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();   // both CTAs' barriers exist before any remote signal; also required before a 2-SM TMEM alloc
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  // cin need not fill its last 64-channel K chunk: TMA zero-fills the A box beyond the tensor's channel extent, and
  // whatever the weight box holds there (the next tap's columns, or zeros past the last one) is multiplied by those zeros
  const int kchunks = (p.cin + TC_BK - 1) / TC_BK;
  const int total_work = hp.groups_m * p.tiles_c;        // one work item = 2 neighbouring pixel tiles x one N tile
  const int cluster_id = blockIdx.x / 2, num_clusters = gridDim.x / 2;
  const int spc = 9 / hp.tps;                            // weight stages per 64-channel chunk
  // work items of this cluster.  Default: round-robin over all (N tile, pixel-tile pair) items.  Weight-stationary mode
  // with several N tiles: cluster c keeps N tile c % tiles_c resident and walks that tile's pixel groups only.
  const int per_ct = hp.per_ct;
  const int my_ct = per_ct ? cluster_id % p.tiles_c : 0, my_lane = per_ct ? cluster_id / p.tiles_c : 0;
  const int my_work = per_ct ? (hp.groups_m - my_lane + per_ct - 1) / per_ct
                             : (total_work - cluster_id + num_clusters - 1) / num_clusters;
  auto work_item = [&](int i) {
    return per_ct ? my_ct * hp.groups_m + my_lane + i * per_ct : cluster_id + i * num_clusters;
  };

  auto decode_tile = [&](int w, int& ct, int& x0, int& y0, int& n) {
    ct = w / hp.groups_m;
    int mt = (w - ct * hp.groups_m) * 2 + (int)crank;    // may be >= tiles_m (phantom tile: all stores masked)
    const int bx = mt % p.tiles_x;
    mt /= p.tiles_x;
    const int by = mt % p.tiles_y;
    n = mt / p.tiles_y;
    x0 = bx * HALO_TW;
    y0 = by * HALO_TH;
  };

  const uint32_t a_tx = (uint32_t)(hp.box_w * HALO_ROWS * 128) * NP;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane per CTA; completion lands on the leader's barriers) =====
    {
      const int total_chunks = my_work * kchunks;
      const int t_star = nbs < spc - 1 ? nbs : spc - 1;
      int a_issue = 0, ab = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto issue_a = [&]() {
        if (a_issue >= total_chunks) return;
        const int wi = work_item(a_issue / kchunks), kc = a_issue % kchunks;
        int ct, x0, y0, n;
        decode_tile(wi, ct, x0, y0, n);
        mbar_wait(a_empty(ab), aph ^ 1u, 21, p.wait_sleep_ns);
        const uint32_t sa = a_base + a_buf_bytes * ab;
        const uint32_t lbar = map_to_cta(a_full(ab), 0);
        if (elect_one()) {
          if (leader) mbar_expect_tx(a_full(ab), 2u * a_tx);
          tma2_load_5d(&tmA_hi, sa, lbar, kc * TC_BK, x0 - 1, 0, y0 - 1, n);
          if (PASSES == 3) tma2_load_5d(&tmA_lo, sa + hp.a_bytes, lbar, kc * TC_BK, x0 - 1, 0, y0 - 1, n);
        }
        __syncwarp();
        if (++ab == na) { ab = 0; aph ^= 1u; }
        ++a_issue;
      };
      for (int i = 0; i < na - 1; ++i) issue_a();
      bool load_b = true;
      for (int wi_ = 0; wi_ < my_work; ++wi_) {
        const int w = work_item(wi_);
        int ct, x0, y0, n;
        decode_tile(w, ct, x0, y0, n);
        const int row_half = ct * bn + (int)crank * (bn / 2);
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int sg = 0; sg < spc; ++sg) {
            if (!hp.resident) mbar_wait(b_empty(bs), bph ^ 1u, 22, p.wait_sleep_ns);
            const uint32_t lbar = map_to_cta(b_full(bs), 0);
            if (load_b && elect_one()) {
              if (leader) mbar_expect_tx(b_full(bs), 2u * b_stage_bytes);
              for (int tt = 0; tt < hp.tps; ++tt) {
                const uint32_t sb = b_base + b_stage_bytes * bs + b_tap_bytes * tt;
                const int kb = (sg * hp.tps + tt) * p.cin + kc * TC_BK;
                if (PASSES == 3) {
                  tma2_load_2d(leader ? &tmB_hi_full : &tmB_lo_full, sb, lbar, kb, ct * bn);
                  tma2_load_2d(&tmB_hi_half, sb + b_plane, lbar, kb, row_half);
                } else {
                  tma2_load_2d(&tmB_hi_half, sb, lbar, kb, row_half);
                }
              }
            }
            __syncwarp();
            if (++bs == nbs) { bs = 0; bph ^= 1u; }
            if (sg == t_star) issue_a();
          }
        }
        if (hp.resident) load_b = false;   // every later work item reuses the weights already in shared memory
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected lane of the LEADER CTA, for the pair =====================
    if (leader) {
      // M = 256 (128 rows per CTA); N counts the rows both CTAs contribute together
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t sbo = (uint32_t)hp.pitch * 128u;
      // halo offset of tap (r, s) in 16 B units: (r * pitch + s) pixels x 128 B -- computed where it is used: a 9-entry
      // table indexed by the (runtime) stage index lives in LOCAL memory, and a local load of the one thread that feeds the
      // tensor core queues behind the drain warps' global accesses in the SM's L1TEX FIFO
      auto tap_off16 = [&](int tap) {
        const int r = tap / 3;
        return (uint32_t)(r * hp.pitch + (tap - 3 * r)) * 8u;
      };
      int ab = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (; it < my_work; ++it) {
        const int buf = (nbuf == 2) ? (it & 1) : 0;
        const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u, 23);
        tc_fence_after();
        const uint32_t d_hh = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t d_lo = d_hh + (uint32_t)bn;
        uint32_t acc = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(a_full(ab), aph, 24);
          const uint64_t a_hi0 = make_sdesc_halo(a_base + a_buf_bytes * ab, sbo, 0);
          const uint64_t a_lo0 = a_hi0 + (uint64_t)(hp.a_bytes >> 4);
          for (int sg = 0; sg < spc; ++sg) {
            if (!hp.resident || it == 0) mbar_wait(b_full(bs), bph, 25);
            tc_fence_after();
            const uint64_t b0 = make_sdesc(b_base + b_stage_bytes * bs);
            if (elect_one()) {
            uint32_t acc_i = acc;
#pragma unroll 3
            for (int tt = 0; tt < hp.tps; ++tt) {
              const uint64_t at = (uint64_t)tap_off16(sg * hp.tps + tt);
              const uint64_t bt = b0 + (uint64_t)(tt * (int)(b_tap_bytes >> 4));
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                if (PASSES == 3) {
                  umma2_f16(d_hh, a_hi0 + at + ko, bt + ko, idesc_2n, acc_i);                            // [hi*hi | hi*lo]
                  umma2_f16(d_lo, a_lo0 + at + ko, bt + (uint64_t)(b_plane >> 4) + ko, idesc_n, 1u);     // += lo*hi
                } else {
                  umma2_f16(d_hh, a_hi0 + at + ko, bt + ko, idesc_n, acc_i);
                }
                acc_i = 1u;
              }
            }
            if (!hp.resident) umma2_commit_mc(b_empty(bs), 3);
            if (sg == spc - 1) umma2_commit_mc(a_empty(ab), 3);
            }
            __syncwarp();
            acc = 1u;
            if (++bs == nbs) { bs = 0; bph ^= 1u; }
          }
          if (++ab == na) { ab = 0; aph ^= 1u; }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar(buf), 3);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps: every CTA drains its own 128 accumulator rows =====================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const uint32_t stage = epi_base + (uint32_t)(warp - 2) * TC_EPI_STAGE_BYTES;
    // the loop carries (it, w) and two bounds, not the cluster's place in the schedule (my_ct, my_lane, cluster_id, ...):
    // every value that stays live across the drain competes with it for the 96 registers
    const int w_step = per_ct ? per_ct : num_clusters;
    const int w_end = per_ct ? my_ct * hp.groups_m + hp.groups_m : total_work;
    int w = work_item(0);
    for (int it = 0; w < w_end; ++it, w += w_step) {
      int ct, x0, y0, n0;
      decode_tile(w, ct, x0, y0, n0);
      auto pix = [&](int row, int& n, int& oy, int& ox) {
        ox = x0 + (row & (HALO_TW - 1));
        oy = y0 + (row >> 3);
        n = n0;
        return (ox < p.wout) && (oy < p.hout) && (n < p.n);
      };
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
      // L2 prefetch of the fp32 operands: this tile's on the first trip, from then on the NEXT tile's (a whole tile of lead)
      // (bulk-store launches prefetch nothing: measured no gain, profiles/r2_drain_ab.txt)
      if (DRAIN != DRAIN_TMA && it == 0) prefetch_epilogue_operands(p, bn, ct, cg, q, lane, pix);
      if (DRAIN != DRAIN_TMA && w + w_step < w_end) {
        int ct2, x2, y2, n2;
        decode_tile(w + w_step, ct2, x2, y2, n2);
        prefetch_epilogue_operands(p, bn, ct2, cg, q, lane, [&](int row, int& n, int& oy, int& ox) {
          ox = x2 + (row & (HALO_TW - 1));
          oy = y2 + (row >> 3);
          n = n2;
          return (ox < p.wout) && (oy < p.hout) && (n < p.n);
        });
      }
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols);
      drain_tile<PASSES, GN, DRAIN>(p, &om, t_acc, bn, ct, cg, q, lane, stage, pix, bias_staged ? bias_smem : p.bias,
                                    effective_w_scale(p), [&]() {
                                      mbar_wait(tfull_bar(buf), use & 1u, 26, p.wait_sleep_ns);
                                      tc_fence_after();
                                    });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(buf));
        else mbar_arrive_remote(tempty_bar(buf), 0);
      }
    }
    if constexpr (DRAIN == DRAIN_TMA) {
      if (lane == 0) bulk_wait_all();          // this lane's bulk stores have left shared memory and are complete
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer must not exit (or free TMEM) while the leader's MMAs still read its smem / write its TMEM
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

}  // namespace mcq
