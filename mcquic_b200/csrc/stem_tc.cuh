// Tensor-core stem: the analysis transform's first layer, conv3x3 stride 2 (RGB -> C) on the NCHW image
// (mcquic/modules/compressor.py:124), with AlignedPadding's reflect pad (mcquic/data/transforms.py:86-99) folded into
// the operand loads and -- for uint8 images -- the reference's input transform `convert_image_dtype` + `(x - 0.5) * 2`
// (mcquic/demo.py:110-118) folded in as well, so a host batch crosses PCIe as bytes.
//
// The FFMA version (stem_conv_kernel) is issue-bound at ~18 % of the FP32 peak (one shared-memory broadcast per four
// FFMAs, 156 registers) and took 0.6 ms of the 14 ms step although it only has to write 1.07 GB.  Here the layer is a
// GEMM with K = 27 (padded to 32) on tcgen05, fp32-grade through the same 3-pass split as every encode-side layer:
//
//   tile   = 128 consecutive output pixels (flattened n, y, x), all C <= 128 output channels
//   A      = [x_hi | x_lo] fp16, 32 + 32 elements per row = one 128 B swizzle row: two producer warps gather the 27
//            taps of a pixel from the image (reflect / zero padding resolved per tap), split them into hi + lo / 2048
//            and write the row straight into the SWIZZLE_128B K-major layout (the fused VQ kernel's A-operand path)
//   B      = [w_lo | w_hi] fp16 rows of w * 2^e (packed once on the host), C x 128 B, resident in shared memory
//   MMA    = D_hh = x_hi.w_hi (K-steps 0,1 of A against 2,3 of B), D_lo = x_hi.w_lo + x_lo.w_hi (K-steps 0..3 of both)
//            -> 6 MMAs per tile; the accumulator pair has the layout of the 3-pass convolution kernels
//   drain  = the convolution kernels' drain_tile<3> (bias, fp32 output, SiLU / raw planes): the layer is bound by the
//            1.07 GB it writes
#pragma once
#include "vq_fused.cuh"

namespace mcq {

constexpr int STC_PROD_WARPS = 2;                                   // 64 producer threads, two tile rows each (19 warps
                                                                    // -> 96 registers per thread; 21 warps would get 80)
constexpr int STC_FIRST_EPI = 1 + STC_PROD_WARPS;                  // warp 0 = TMEM owner + MMA issuer
constexpr int STC_THREADS = 32 * (STC_FIRST_EPI + TC_EPI_WARPS);   // 608
constexpr int STC_BM = 128;
constexpr int STC_A_BYTES = STC_BM * 128;                           // 16 KB: 128 rows x [32 hi | 32 lo] fp16

struct StemTcArgs {
  const float* x_f32;          // [n, 3, h, w] fp32 in [-1, 1]   (exactly one of x_f32 / x_u8)
  const unsigned char* x_u8;   // [n, 3, h, w] uint8: value = (u / 255 - 0.5) * 2
  const __half* w_lohi;        // [cout_pad, 64]: [w_lo(27 + 5 zeros) | w_hi(27 + 5 zeros)], K order ci*9 + r*3 + s
  int h, w, pad_top, pad_left, hp, wp;
  int cout_pad;
};

// the reference's uint8 -> [-1, 1] transform, same fp32 operations in the same order (demo.py:110,117)
__device__ __forceinline__ float u8_to_unit(unsigned char u) { return (__fdiv_rn((float)u, 255.0f) - 0.5f) * 2.0f; }

__device__ __forceinline__ int stem_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <int DRAIN>
__global__ void __launch_bounds__(STC_THREADS, 1) stem_tc_kernel(const ConvArgs p, const StemTcArgs s,
                                                                 const __grid_constant__ OutMaps om) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = p.cout_pad;                                 // <= 128, multiple of 16
  const uint32_t a_base = smem_base;                          // 2 x 16 KB
  const uint32_t b_base = a_base + 2u * STC_A_BYTES;          // cout_pad x 128 B (<= 16 KB, 1024-aligned)
  const uint32_t bar_base = b_base + 16384u;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto tfull = [&](int i) { return bar_base + 8u * (4 + i); };
  auto tempty = [&](int i) { return bar_base + 8u * (6 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 64u);
  const uint32_t epi_base = bar_base + 512u;                                          // 16 x 2 KB drain staging (512 B aligned)
  float* bias_smem = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base) + TC_EPI_WARPS * TC_EPI_STAGE_BYTES);

  float* u8_lut = bias_smem + 128;                                                    // 256 floats
  for (int i = threadIdx.x; i < p.cout; i += STC_THREADS) bias_smem[i] = p.bias[i];
  if (s.x_u8)
    for (int i = threadIdx.x; i < 256; i += STC_THREADS) u8_lut[i] = u8_to_unit((unsigned char)i);
  // weights -> shared memory in the swizzled K-major layout (16 B chunk j of row r at chunk j ^ (r & 7))
  for (int i = threadIdx.x; i < bn * 8; i += STC_THREADS) {
    const int r = i >> 3, j = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(s.w_lohi + (size_t)r * 64 + j * 8);
    const uint32_t dst = b_base + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + ((uint32_t)(j ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  fence_async_smem();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full(i), STC_PROD_WARPS * 32);
      mbar_init(a_empty(i), 1);
      mbar_init(tfull(i), 1);
      mbar_init(tempty(i), TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
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

  const long long total_pix = (long long)p.n * p.hout * p.wout;
  const int total_tiles = (int)((total_pix + STC_BM - 1) / STC_BM);
  const int hw_out = p.hout * p.wout;
  const int acc_cols = 2 * bn;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(STC_BM >> 4) << 24);
    const uint64_t b0 = make_sdesc(b_base);
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i & 1;
      const uint32_t use = (uint32_t)(i >> 1);
      mbar_wait_sleep(tempty(ab), (use & 1u) ^ 1u, 41, 60);
      mbar_wait_sleep(a_full(ab), use & 1u, 42, 60);
      tc_fence_after();
      const uint64_t a0 = make_sdesc(a_base + (uint32_t)ab * STC_A_BYTES);
      const uint32_t d_hh = tmem_base + (uint32_t)(ab * acc_cols);
      const uint32_t d_lo = d_hh + (uint32_t)bn;
      if (elect_one()) {
        // K-step j (16 elements) of a row sits at +32 B * j; descriptor addresses are in 16 B units
        for (int j = 0; j < 2; ++j) umma_f16(d_hh, a0 + (uint64_t)(2 * j), b0 + (uint64_t)(2 * (2 + j)), idesc, j > 0 ? 1u : 0u);
        for (int j = 0; j < 4; ++j) umma_f16(d_lo, a0 + (uint64_t)(2 * j), b0 + (uint64_t)(2 * j), idesc, j > 0 ? 1u : 0u);
        umma_commit(a_empty(ab));
        umma_commit(tfull(ab));
      }
      __syncwarp();
    }
  } else if (warp < STC_FIRST_EPI) {
    // ===================== operand producer: im2col rows, split fp16, swizzled K-major =====================
    const int ptid = threadIdx.x - 32;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i & 1;
      mbar_wait_sleep(a_empty(ab), (((uint32_t)(i >> 1)) & 1u) ^ 1u, 43, 200);
      const uint32_t abuf = a_base + (uint32_t)ab * STC_A_BYTES;
      // both rows of this thread: all 54 tap loads are issued before any is consumed (one memory round trip per tile
      // instead of two -- the producers' latency, not the tensor pipe, paces this kernel's mainloop)
      float v[2][27];
      bool ok[2][9];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = rr * (STC_PROD_WARPS * 32) + ptid;
        const long long f = (long long)t * STC_BM + row;
        const bool live = f < total_pix;
        const long long fc = live ? f : 0;
        const int n = (int)(fc / hw_out);
        const int rem = (int)(fc - (long long)n * hw_out);
        const int oy = rem / p.wout, ox = rem - oy * p.wout;
        // the 3 source rows / columns of this pixel's taps: always in-range indices (the loads are unconditional --
        // predicated loads were serialised by the compiler: 27 dependent L2 round trips per row); zero padding is
        // applied to the loaded values
        int sy[3], sx[3];
        bool vy[3], vx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int Y = 2 * oy + k - 1, X = 2 * ox + k - 1;       // coordinates in the padded image (zeros outside it)
          vy[k] = live && Y >= 0 && Y < s.hp;
          vx[k] = X >= 0 && X < s.wp;
          sy[k] = min(max(stem_reflect(Y - s.pad_top, s.h), 0), s.h - 1);
          sx[k] = min(max(stem_reflect(X - s.pad_left, s.w), 0), s.w - 1);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) ok[rr][k] = vy[k / 3] && vx[k % 3];
        const size_t img = (size_t)n * 3 * s.h * s.w;
        if (s.x_u8) {
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                v[rr][ci * 9 + r * 3 + c] =
                    __uint_as_float((uint32_t)__ldg(s.x_u8 + img + ((size_t)ci * s.h + sy[r]) * s.w + sx[c]));
        } else {
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                v[rr][ci * 9 + r * 3 + c] = __ldg(s.x_f32 + img + ((size_t)ci * s.h + sy[r]) * s.w + sx[c]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = rr * (STC_PROD_WARPS * 32) + ptid;
        if (s.x_u8) {
          // uint8 -> [-1, 1] through a 256-entry table of the reference's transform (filled with the same fp32 operations)
#pragma unroll
          for (int k = 0; k < 27; ++k) v[rr][k] = u8_lut[__float_as_uint(v[rr][k])];
        }
#pragma unroll
        for (int k = 0; k < 27; ++k)
          if (!ok[rr][k % 9]) v[rr][k] = 0.f;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 13; ++e) split_f32x2(v[rr][2 * e], v[rr][2 * e + 1], hi[e], lo[e]);
        split_f32x2(v[rr][26], 0.f, hi[13], lo[13]);
        hi[14] = hi[15] = lo[14] = lo[15] = 0u;
        const uint32_t rbase = abuf + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((((uint32_t)j) ^ sw) << 4)),
                       "r"(hi[4 * j]), "r"(hi[4 * j + 1]), "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((((uint32_t)(4 + j)) ^ sw) << 4)),
                       "r"(lo[4 * j]), "r"(lo[4 * j + 1]), "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3])
                       : "memory");
        }
      }
      fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(a_full(ab));
    }
  } else {
    // ===================== drain: the convolution kernels' fused epilogue =====================
    const int q = warp & 3;
    const int cg = (warp - STC_FIRST_EPI) >> 2;
    const uint32_t stage = epi_base + (uint32_t)(warp - STC_FIRST_EPI) * TC_EPI_STAGE_BYTES;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int ab = i & 1;
      const long long f0 = (long long)t * STC_BM;
      auto pix = [&](int row, int& n, int& oy, int& ox) {
        const long long f = f0 + row;
        if constexpr (DRAIN == DRAIN_TMA) {
          // bulk-store drain: the output is viewed as [pixels, channels] (host: total_pix < 2^31), "x" = flat pixel index
          n = 0; oy = 0; ox = (int)f;
          return f < total_pix;
        }
        const long long fc = f < total_pix ? f : 0;
        n = (int)(fc / hw_out);
        const int rem = (int)(fc - (long long)n * hw_out);
        oy = rem / p.wout;
        ox = rem - oy * p.wout;
        return f < total_pix;
      };
      mbar_wait(tfull(ab), ((uint32_t)(i >> 1)) & 1u, 44);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * acc_cols);
      drain_tile<3, false, DRAIN>(p, &om, t_acc, bn, 0, cg, q, lane, stage, pix, bias_smem, effective_w_scale(p), []() {});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(ab));
    }
    if constexpr (DRAIN == DRAIN_TMA) {
      if (lane == 0) bulk_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

}  // namespace mcq
