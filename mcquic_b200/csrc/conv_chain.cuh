// Layer-chain kernel: a whole run of dependent convolutions on small feature maps in ONE persistent launch.
//
// 137 of the 165 convolutions per direction work on <= 16x16 maps.  One such layer is
// ~5 GFLOP at batch 64 -- a few microseconds of tensor-core time -- but as its own launch it costs 16-26 us
// (launch + prologue + pipeline fill + drain), so the low-resolution tail took a third of the step.  Here the
// tail's layers are executed back to back by the same resident CTAs:
//
// * Images are independent through the whole path, so a thread-block CLUSTER of 8 CTAs owns a fixed group of
//   `ipc` images for every layer of the chain; a layer's tiles of that group (pixel tiles x N slices) are dealt
//   round-robin to the 8 CTAs.  The only synchronisation a layer boundary needs is therefore cluster-wide:
//   barrier.cluster (release/acquire) + fence.proxy.async, because the next layer's TMA loads (async proxy) read
//   what this layer's epilogue wrote with ordinary stores.  Activations travel through L2 (the tail's working
//   set is a few MB), weights stream from L2 as in conv_tc.cuh.
// * Layers that do not read anything written since the previous barrier (the two branches of an AttentionBlock,
//   dequantizationHead || sideHead) carry sync_before = 0 and simply extend the tile list: both branches fill
//   the pipeline together.
// * Per-layer constants (ConvArgs + 4 tensor maps) arrive as one __grid_constant__ kernel parameter; the biases of
//   all layers are staged in shared memory once.
//
// Tile body (TMA producer warp / MMA issuer warp / 16 epilogue warps, TMEM double buffering, fused epilogue) is
// the one of conv_tc.cuh; results are bit-identical to launching the layers one by one.
#pragma once
#include "conv_halo.cuh"
#include "conv_tc.cuh"

namespace mcq {

constexpr int CHAIN_MAX_LAYERS = 28;
constexpr int CHAIN_CL = 8;                  // CTAs per cluster (portable maximum)
constexpr int CHAIN_BIAS_FLOATS = 4096;      // shared-memory bias pool of one chain (16 KB)
constexpr int CHAIN_PROD1_WARP = 2 + TC_EPI_WARPS;               // warp 18: second TMA producer
constexpr int CHAIN_THREADS = TC_THREADS + 32;

struct alignas(64) ChainLayer {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  ConvArgs p;
  int sync_before;   // 1: the layer reads something an earlier layer of this chain wrote after the last barrier
  int bias_off;      // offset of its bias vector in the shared-memory pool (floats)
};

struct alignas(64) ChainParams {
  ChainLayer layers[CHAIN_MAX_LAYERS];
  int count;
  int ipc;           // images per cluster
  int stages;        // TMA pipeline depth
  int bias_total;
  long long* dbg;    // MCQ debug timeline (clock64 samples of CTA 0), or nullptr
  int debug;         // bit 0 (MCQ_CHAIN_NOSYNC=1): skip the layer barriers -- WRONG results, timing experiments only
};

__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// layer-boundary synchronisation, writer side (epilogue warps) / reader side (TMA producer) / bystander (MMA warp)
//   mode 0: fence.proxy.async + barrier.cluster release/acquire     mode 1: fence.proxy.async + relaxed barrier
//   mode 2: __threadfence + relaxed barrier (no proxy fence: timing experiment only)
__device__ __forceinline__ void chain_sync(int mode, int role) {
  if (mode == 0) {
    if (role == 2) asm volatile("fence.proxy.async;" ::: "memory");
    cluster_sync_all();
  } else if (mode == 1) {
    if (role == 2) asm volatile("fence.proxy.async;" ::: "memory");
    cluster_sync_relaxed();
  } else if (mode == 2) {
    if (role == 2) __threadfence();
    cluster_sync_relaxed();
  } else {
    cluster_sync_relaxed();   // mode 3: barrier only (timing experiment)
  }
}

template <int PASSES>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NP = (PASSES == 3) ? 2 : 1;
  constexpr int BN_MAX = (PASSES == 3) ? 64 : 128;
  constexpr uint32_t stage_bytes = (uint32_t)(TC_A_BYTES + BN_MAX * TC_BK * 2) * NP;   // fixed for every layer
  constexpr uint32_t ACC_STRIDE = 256;   // TMEM columns between the two accumulator buffers (independent of bn)
  const int stages = cp.stages;
  const int count = cp.count;
  const bool nosync = (cp.debug & 1) != 0;
  const int sync_mode = (cp.debug >> 1) & 3;

  const uint32_t bar_base = smem_base + stage_bytes * stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * stages + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + stage_bytes * stages + 8u * (2 * stages + 4));
  const uint32_t epi_base = (bar_base + 8u * (2 * stages + 4) + 16u + 127u) & ~127u;
  float* bias_pool = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base) + TC_EPI_WARPS * TC_EPI_STAGE_BYTES);

  // biases are weights, not outputs of the previous kernel: stage them before the grid dependency resolves
  for (int l = 0; l < count; ++l) {
    const float* src = cp.layers[l].p.bias;
    float* dst = bias_pool + cp.layers[l].bias_off;
    for (int i = threadIdx.x; i < cp.layers[l].p.cout; i += CHAIN_THREADS) dst[i] = src[i];
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  const int crank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x / CHAIN_CL;
  const int img0 = cluster_id * cp.ipc;

  // geometry of layer l restricted to this cluster's images
  struct LayerGeo {
    int nimg, tiles_m, total;
  };
  auto geo = [&](const ConvArgs& p) {
    LayerGeo g;
    g.nimg = min(cp.ipc, p.n - img0);
    if (g.nimg < 0) g.nimg = 0;
    const int tiles_nl = (g.nimg + p.tn - 1) / p.tn;
    g.tiles_m = tiles_nl * p.tiles_y * p.tiles_x;
    g.total = g.tiles_m * p.tiles_c;
    return g;
  };

  if (warp == 0 || warp == CHAIN_PROD1_WARP) {
    // ===================== TMA producers =====================
    // Two warps share the k-iterations (even / odd in the CTA-wide sequence): issuing one stage costs a serial chain
    // of ~400 cycles (barrier check, coordinates, uniform-register moves, expect_tx, 2-4 UTMALDG), more than the MMAs
    // of a narrow tile take, so a single producer warp was the bottleneck of the whole k-loop.
    const int pj = (warp == 0) ? 0 : 1;
    int dbg_k = 0;
    int kg = 0;          // k-iterations this CTA has gone through before the current tile (same count in both warps)
    int my_k = pj;       // index of this warp's next k-iteration in that sequence
    int s = pj % stages;
    uint32_t ph = (uint32_t)(pj / stages) & 1u;
    for (int l = 0; l < count; ++l) {
      const ChainLayer& L = cp.layers[l];
      const ConvArgs& p = L.p;
      // constants of the layer (immutable kernel parameters): fetched before the barrier, off the critical path
      if (pj == 0 && lane == 0) {
        tma_prefetch_desc(&L.tmA_hi);
        tma_prefetch_desc(&L.tmB_hi);
        if (PASSES == 3) {
          tma_prefetch_desc(&L.tmA_lo);
          tma_prefetch_desc(&L.tmB_lo);
        }
      }
      const LayerGeo g = geo(p);
      const int bn = p.bn;
      const uint32_t b_bytes = (uint32_t)bn * TC_BK * 2;
      const uint32_t tx_bytes = (TC_A_BYTES + b_bytes) * NP;
      const int ksize = p.ksize, stride = p.stride, cin = p.cin, ch_off = p.ch_off;
      const int kchunks = cin / TC_BK;
      const int kiters = ksize * ksize * kchunks;
      const int pad = ksize >> 1;
      const int tiles_x = p.tiles_x, tiles_y = p.tiles_y, tw = p.tw, th = p.th, tn = p.tn;
      if (l > 0 && L.sync_before && !nosync) {
        if (cp.dbg && blockIdx.x == 0 && lane == 0 && pj == 0 && l < 32) cp.dbg[2560 + 8 * l + 4] = clock64();
        chain_sync(sync_mode, 0);
        if (cp.dbg && blockIdx.x == 0 && lane == 0 && pj == 0 && l < 32) cp.dbg[2560 + 8 * l + 5] = clock64();
      }
      for (int t = crank; t < g.total; t += CHAIN_CL) {
        const int ct = t / g.tiles_m;
        int mt = t - ct * g.tiles_m;
        const int bx = mt % tiles_x;
        mt /= tiles_x;
        const int by = mt % tiles_y;
        const int bz = mt / tiles_y;
        const int x0 = bx * tw, y0 = by * th, n0 = img0 + bz * tn, c0 = ct * bn;
        for (int kk = my_k - kg; kk < kiters; kk += 2) {
          int tap, kc;
          if (kchunks == 2) { tap = kk >> 1; kc = kk & 1; } else { tap = kk / kchunks; kc = kk - tap * kchunks; }
          // filter tap (r, s) -> TMA coordinates in the 5-D activation view (see mcq_api.cu fill_taps)
          const int tr = (tap * 11) >> 5;                    // tap / 3 for tap < 9
          const int r = (ksize == 3 ? tr : 0) - pad, sx = (ksize == 3 ? tap - 3 * tr : 0) - pad;
          int t_c, t_dx, t_py, t_dy;
          if (stride == 1) { t_c = 0; t_dx = sx; t_py = 0; t_dy = r; }
          else { t_py = r & 1; t_dy = r < 0 ? -1 : 0; t_c = (sx & 1) * cin; t_dx = sx < 0 ? -1 : 0; }
          mbar_wait(empty_bar(s), ph ^ 1u, 31);
          if (cp.dbg && blockIdx.x == 0 && lane == 0 && dbg_k < 512) cp.dbg[512 * pj + dbg_k++] = clock64();
          const uint32_t sa = smem_base + stage_bytes * s;
          const int ca = ch_off + t_c + kc * TC_BK;
          const int kb = tap * cin + kc * TC_BK;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), tx_bytes);
            tma_load_5d(&L.tmA_hi, sa, full_bar(s), ca, x0 + t_dx, t_py, y0 + t_dy, n0);
            if (PASSES == 3) {
              tma_load_5d(&L.tmA_lo, sa + TC_A_BYTES, full_bar(s), ca, x0 + t_dx, t_py, y0 + t_dy, n0);
              tma_load_2d(&L.tmB_hi, sa + 2 * TC_A_BYTES, full_bar(s), kb, c0);
              tma_load_2d(&L.tmB_lo, sa + 2 * TC_A_BYTES + b_bytes, full_bar(s), kb, c0);
            } else {
              tma_load_2d(&L.tmB_hi, sa + TC_A_BYTES, full_bar(s), kb, c0);
            }
          }
          __syncwarp();
          my_k += 2;
          s += 2;
          while (s >= stages) { s -= stages; ph ^= 1u; }
        }
        kg += kiters;
        if (my_k < kg) {          // (only if a tile had fewer than two k-iterations) catch up, keeping this warp's parity
          const int d = ((kg - my_k + 1) >> 1) << 1;
          my_k += d;
          s += d;
          while (s >= stages) { s -= stages; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int dbg_k = 0;
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int l = 0; l < count; ++l) {
      const ChainLayer& L = cp.layers[l];
      const ConvArgs& p = L.p;
      const LayerGeo g = geo(p);
      const int bn = p.bn;
      const uint32_t b_bytes = (uint32_t)bn * TC_BK * 2;
      const int kiters = p.ksize * p.ksize * (p.cin / TC_BK);
      if (l > 0 && L.sync_before && !nosync) chain_sync(sync_mode, 1);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc_wide = (1u << 4) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int t = crank; t < g.total; t += CHAIN_CL, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u, 32);
        tc_fence_after();
        const uint32_t d_hh = tmem_base + (uint32_t)buf * ACC_STRIDE;
        const uint32_t d_lo = d_hh + (uint32_t)bn;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full_bar(s), ph, 33);
          tc_fence_after();
          if (cp.dbg && blockIdx.x == 0 && lane == 0 && dbg_k < 1024) cp.dbg[1024 + dbg_k++] = clock64();
          const uint32_t sa = smem_base + stage_bytes * s;
          const uint64_t a_hi = make_sdesc(sa);
          const uint64_t a_lo = make_sdesc(sa + TC_A_BYTES);
          const uint64_t b_hi = make_sdesc(sa + NP * TC_A_BYTES);
          const uint32_t acc0 = ki > 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint32_t acc = k > 0 ? 1u : acc0;
              const uint64_t ko = (uint64_t)(k * 2);
              if (PASSES == 3) {
                // [B_hi ; B_lo] adjacent in smem, acc_hh / acc_lo adjacent in TMEM: one N = 2*bn MMA (bn <= 64)
                umma_f16(d_hh, a_hi + ko, b_hi + ko, idesc_wide, acc);
                umma_f16(d_lo, a_lo + ko, b_hi + ko, idesc, 1u);
              } else {
                umma_f16(d_hh, a_hi + ko, b_hi + ko, idesc, acc);
              }
            }
            umma_commit(empty_bar(s));
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(tfull_bar(buf));
        __syncwarp();
      }
      (void)b_bytes;
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const uint32_t stage = epi_base + (uint32_t)(warp - 2) * TC_EPI_STAGE_BYTES;
    int it = 0;
    for (int l = 0; l < count; ++l) {
      const ChainLayer& L = cp.layers[l];
      const ConvArgs& p = L.p;
      if (l > 0 && L.sync_before && !nosync) {
        // our stores of the layers before must be visible to the TMA engines of the whole cluster
        if (cp.dbg && blockIdx.x == 0 && threadIdx.x == 64 && l < 32) cp.dbg[2560 + 8 * l + 0] = clock64();
        if (cp.dbg && blockIdx.x < 8 && threadIdx.x == 64 && l < 16) cp.dbg[3072 + 64 * blockIdx.x + 4 * l] = global_ns();
        if (cp.dbg && blockIdx.x < 8 && threadIdx.x == 64 && l < 16) cp.dbg[3072 + 64 * blockIdx.x + 4 * l + 1] = clock64();
        chain_sync(sync_mode, 2);
        if (cp.dbg && blockIdx.x == 0 && threadIdx.x == 64 && l < 32) cp.dbg[2560 + 8 * l + 2] = clock64();
        if (cp.dbg && blockIdx.x < 8 && threadIdx.x == 64 && l < 16) cp.dbg[3072 + 64 * blockIdx.x + 4 * l + 2] = global_ns();
        if (cp.dbg && blockIdx.x < 8 && threadIdx.x == 64 && l < 16) cp.dbg[3072 + 64 * blockIdx.x + 4 * l + 3] = clock64();
      }
      const LayerGeo g = geo(p);
      const int bn = p.bn;
      const int n_end = img0 + g.nimg;
      const float* bias_src = bias_pool + L.bias_off;
      for (int t = crank; t < g.total; t += CHAIN_CL, ++it) {
        const int ct = t / g.tiles_m;
        int mt = t - ct * g.tiles_m;
        const int bx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int by = mt % p.tiles_y;
        const int bz = mt / p.tiles_y;
        auto pix = [=, &p](int row, int& n, int& oy, int& ox) {
          ox = bx * p.tw + row % p.tw;
          oy = by * p.th + (row / p.tw) % p.th;
          n = img0 + bz * p.tn + row / (p.tw * p.th);
          return (ox < p.wout) && (oy < p.hout) && (n < n_end);
        };
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tfull_bar(buf), use & 1u, 34);
        tc_fence_after();
        if (cp.dbg && blockIdx.x == 0 && threadIdx.x == 64 && it < 256) cp.dbg[2048 + 2 * it] = clock64();
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * ACC_STRIDE;
        drain_tile<PASSES>(p, nullptr, t_acc, bn, ct, cg, q, lane, stage, pix, bias_src, effective_w_scale(p), []() {});
        if (cp.dbg && blockIdx.x == 0 && threadIdx.x == 64 && it < 256) cp.dbg[2048 + 2 * it + 1] = clock64();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));
      }
    }
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
