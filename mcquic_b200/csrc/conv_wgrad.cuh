// Weight gradient of a convolution on tcgen05 (training step, SURVEY.md section 8f NEXT-3):
//
//   dW[co, tap, ci] = sum over output pixels p of  dY[p, co] * X[p @ tap, ci]
//
// i.e. per filter tap a GEMM with M = cout, N = cin and the REDUCTION over pixels.  Both operands are the NHWC split-fp16
// planes the forward / dgrad convolutions use anyway (X = the forward A operand, dY = the dgrad A operand): a
// (128 pixels x 64 channels) TMA box of such a plane, 128B-swizzled, is exactly tcgen05's canonical *MN-major* operand
// layout (rows = K = pixels, 128 contiguous bytes = 64 MN elements; 8-row groups 1024 B apart = SBO, 64-channel chunks
// one box apart = LBO), so no transposed copy of either tensor is ever made -- the instruction descriptor just marks A
// and B as MN-major.  The tap shift of X (and its zero padding) is the TMA coordinate offset, as in conv_tc.cuh.
//
//   work item  = (128 output channels) x (128 input channels) x (group of up to 3 taps): 3 x 128 TMEM columns
//   split-K    = the pixel tiles of a work item are dealt round-robin to `splits` CTAs; every CTA writes its fp32
//                partial block [tap][128][128] to a workspace, wgrad_reduce_kernel sums the splits in a fixed order
//                (deterministic), applies the scale and scatters into nn.Conv2d's [cout, cin, kh, kw] layout
//   precision  = one fp16 pass (hi planes), fp32 accumulation: TF32-grade, what the reference trains with
//                (mcquic/train/utils.py: allow_tf32)
//   warps      = 0: TMA producer, 1: MMA issuer, 2..5: drain (one per TMEM lane quarter)
#pragma once
#include "conv_tc.cuh"

namespace mcq {

constexpr int WG_THREADS = 32 * 6;
constexpr int WG_MAX_TAPS = 3;                       // accumulators per CTA (3 x 128 of the 512 TMEM columns)
constexpr int WG_A_BYTES = 2 * TC_A_BYTES;           // dY tile: 128 pixels x 128 channels (two 64-channel boxes)
constexpr int WG_B_BYTES = 2 * TC_A_BYTES;           // X tile of one tap: 128 pixels x 128 channels
constexpr int WG_NA = 2, WG_NB = 4;                  // ring depths: 64 KB + 128 KB

struct WgradArgs {
  float* partial;              // [items][splits][taps_per_item][128][128] fp32
  int n, hout, wout;           // output (dY) grid
  int cin, cout;
  int ksize, ntaps;            // ntaps = ksize^2
  int tw, th, tn, tiles_x, tiles_y, tiles_n;   // pixel tile = (tw x th x tn) box of OUTPUT pixels, 128 of them
  int tiles_ci, tiles_co;      // ceil(c / 128)
  int tap_groups, taps_per_group;
  int splits;
  // per-tap TMA coordinate offsets in the 5-D view of X (stride 1: plain; stride 2: parity view, see conv_tc.cuh)
  int tap_c[9], tap_dx[9], tap_py[9], tap_dy[9];
};

// MN-major SWIZZLE_128B operand: 64-element (128 B) rows, 8-row groups `sbo` bytes apart, 64-element column blocks
// `lbo` bytes apart
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + WG_NA * WG_A_BYTES;
  const uint32_t bar_base = b_base + WG_NB * WG_B_BYTES;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (WG_NA + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * WG_NA + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (2 * WG_NA + WG_NB + i); };
  const uint32_t done_bar = bar_base + 8u * (2 * WG_NA + 2 * WG_NB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8u * (2 * WG_NA + 2 * WG_NB + 1));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < WG_NA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < WG_NB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    mbar_init(done_bar, 1);
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

  // work item of this CTA
  const int split = blockIdx.x % a.splits;
  int item = blockIdx.x / a.splits;
  const int tg = item % a.tap_groups;
  item /= a.tap_groups;
  const int ci_t = item % a.tiles_ci, co_t = item / a.tiles_ci;
  const int tap0 = tg * a.taps_per_group;
  const int ntap = min(a.taps_per_group, a.ntaps - tap0);
  const int tiles_pix = a.tiles_n * a.tiles_y * a.tiles_x;
  const int my_tiles = (tiles_pix - split + a.splits - 1) / a.splits;     // tiles split, split + splits, ...

  if (warp == 0) {
    // ===================== TMA producer =====================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int mt = split + i * a.splits;
      const int bx = mt % a.tiles_x;
      mt /= a.tiles_x;
      const int by = mt % a.tiles_y;
      const int bz = mt / a.tiles_y;
      const int x0 = bx * a.tw, y0 = by * a.th, n0 = bz * a.tn;
      mbar_wait(a_empty(as), aph ^ 1u, 51);
      if (elect_one()) {
        mbar_expect_tx(a_full(as), WG_A_BYTES);
        const uint32_t sa = a_base + (uint32_t)as * WG_A_BYTES;
        tma_load_5d(&tmDY, sa, a_full(as), co_t * 128, x0, 0, y0, n0);
        tma_load_5d(&tmDY, sa + TC_A_BYTES, a_full(as), co_t * 128 + 64, x0, 0, y0, n0);
      }
      __syncwarp();
      if (++as == WG_NA) { as = 0; aph ^= 1u; }
      for (int t = 0; t < ntap; ++t) {
        const int tap = tap0 + t;
        mbar_wait(b_empty(bs), bph ^ 1u, 52);
        if (elect_one()) {
          mbar_expect_tx(b_full(bs), WG_B_BYTES);
          const uint32_t sb = b_base + (uint32_t)bs * WG_B_BYTES;
          const int c = a.tap_c[tap] + ci_t * 128;
          tma_load_5d(&tmX, sb, b_full(bs), c, x0 + a.tap_dx[tap], a.tap_py[tap], y0 + a.tap_dy[tap], n0);
          tma_load_5d(&tmX, sb + TC_A_BYTES, b_full(bs), c + 64, x0 + a.tap_dx[tap], a.tap_py[tap], y0 + a.tap_dy[tap], n0);
        }
        __syncwarp();
        if (++bs == WG_NB) { bs = 0; bph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D[128 co x 128 ci] (fp32) += A^T B: A = dY tile, B = X tile, both MN-major (bits 15 / 16), K = 16 pixels per MMA
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(a_full(as), aph, 53);
      const uint64_t adesc = make_sdesc_mn(a_base + (uint32_t)as * WG_A_BYTES, TC_A_BYTES, 1024u);
      for (int t = 0; t < ntap; ++t) {
        mbar_wait(b_full(bs), bph, 54);
        tc_fence_after();
        const uint64_t bdesc = make_sdesc_mn(b_base + (uint32_t)bs * WG_B_BYTES, TC_A_BYTES, 1024u);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BM / 16; ++k) {      // 16 pixel rows = 2048 B per K step (descriptor units of 16 B)
            const uint64_t ko = (uint64_t)(k * 128);
            umma_f16(tmem_base + (uint32_t)(t * 128), adesc + ko, bdesc + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(b_empty(bs));
          if (t == ntap - 1) umma_commit(a_empty(as));
        }
        __syncwarp();
        if (++bs == WG_NB) { bs = 0; bph ^= 1u; }
      }
      if (++as == WG_NA) { as = 0; aph ^= 1u; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ===================== drain: TMEM lane = output channel row, columns = input channels =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* out = a.partial + ((size_t)blockIdx.x * WG_MAX_TAPS) * 128 * 128;
    if (my_tiles > 0) {
      mbar_wait(done_bar, 0, 55);
      tc_fence_after();
    }
    for (int t = 0; t < ntap; ++t) {
      float* dst = out + ((size_t)t * 128 + row) * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        if (my_tiles > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 128 + c0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
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

// dw[co, ci, kh, kw] = scale * (*dev_scale) * sum_split partial[item(co, ci, tap)][split][tap_local][co % 128][ci % 128]
__global__ void wgrad_reduce_kernel(const float* partial, float* dw, const WgradArgs a, float scale, const float* dev_scale,
                                    int accumulate) {
  const long long total = (long long)a.cout * a.cin * a.ntaps;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // thread order: ci fastest (coalesced partial reads), then tap, then co
  const int ci = (int)(idx % a.cin);
  const int tap = (int)((idx / a.cin) % a.ntaps);
  const int co = (int)(idx / ((long long)a.cin * a.ntaps));
  const int tg = tap / a.taps_per_group, tl = tap - tg * a.taps_per_group;
  const int item = ((co >> 7) * a.tiles_ci + (ci >> 7)) * a.tap_groups + tg;
  const float* src = partial + ((size_t)item * a.splits * WG_MAX_TAPS + tl) * 128 * 128 + (size_t)(co & 127) * 128 + (ci & 127);
  float s = 0.f;
  for (int sp = 0; sp < a.splits; ++sp) s += src[(size_t)sp * WG_MAX_TAPS * 128 * 128];
  s *= scale * (dev_scale ? __ldg(dev_scale) : 1.0f);
  float* d = dw + ((size_t)co * a.cin + ci) * a.ntaps + tap;       // [cout][cin][kh*kw]
  *d = accumulate ? *d + s : s;
}

}  // namespace mcq
