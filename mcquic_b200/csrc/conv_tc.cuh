// tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel @ tap, cin] * W[cout, tap, cin]
//
// * A is never materialised: for every filter tap the 128-pixel x 64-channel operand tile is one TMA box of
//   the NHWC activation (5-D view, zero fill outside the image = the conv's zero padding), landing in shared
//   memory already in the K-major SWIZZLE_128B layout tcgen05.mma consumes.  Stride-2 convolutions use the
//   view [n, h/2, 2, w/2, 2*c] so that every tap is again a unit-stride box.
// * B (packed weights [cout_pad, taps*cin]) is a 2-D TMA box per (tap, 64-channel chunk).
// * One elected thread issues tcgen05.mma (kind::f16, M=128, N=bn, K=16) into TMEM accumulators.
//   PASSES==3: acc_hh += Ahi*Bhi, acc_lo += Ahi*Blo + Alo*Bhi (split-fp16, fp32-grade); PASSES==1: Ahi*Bhi.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..17 = epilogue
//   (tcgen05.ld -> registers -> fused epilogue -> global; warp w owns TMEM lanes 32*(w%4).. and every 4th
//   32-column chunk, so 16 warps share one 128x128 tile).  The accumulator is double-buffered in TMEM
//   when it fits, so the epilogue of tile i overlaps the MMAs of tile i+1.  Persistent CTAs, static schedule.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mcq {

constexpr int TC_EPI_WARPS = 16;                 // 4 TMEM lane quarters x 4 column groups
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                        // fp16 elements per stage along K (= 128 B swizzle span)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;    // 16 KB
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr unsigned long long TC_WATCHDOG_NS = 4ull * 1000 * 1000 * 1000;   // 4 s: far beyond any legitimate wait
constexpr int TC_EPI_STAGE_BYTES = 2048;          // per epilogue warp: 32 rows x 16 fp32 columns
constexpr int TC_BIAS_SMEM_FLOATS = 384;            // bias vectors up to this length are staged in shared memory

__device__ int g_watchdog_flag = 0;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait: the thread is suspended in hardware until the phase completes or a time limit elapses.
// MCQ_WAIT_HINT_NS > 0 passes an explicit suspend-time hint (ptxas turns it into NANOSLEEP.SYNCS <ns>);
// 0 = the plain form with the system-dependent limit.
#ifndef MCQ_WAIT_HINT_NS
#define MCQ_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if defined(MCQ_WAIT_SPIN)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#elif MCQ_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)MCQ_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug (wrong tx byte count, bad tensor map) traps instead of hanging the GPU.
// sleep_ns > 0: the warp sleeps between polls.  For the warps that are not on the critical path of the tensor pipe (TMA
// producer waiting for a free stage, drain warps waiting for an accumulator) this takes their poll loops -- a fifth of all
// executed instructions in the ncu captures -- out of the issue slots and the power budget of a power-capped GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int code, unsigned sleep_ns = 0) {
  const unsigned long long t0 = global_ns();
  for (unsigned i = 1;; ++i) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if (mbar_try_wait(bar, parity)) return;
    if ((i & 63u) == 0 && global_ns() - t0 > TC_WATCHDOG_NS) {
      atomicExch(&g_watchdog_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code, unsigned sleep_ns = 0) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, code, sleep_ns);
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// one lane of a converged warp; values computed warp-uniformly outside the guarded region stay in uniform registers,
// which is what UTMALDG / UTCHMMA take (inside an `if (lane == 0)` region the compiler has to broadcast every operand
// through R2UR in a loop: ~20 instructions per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 operands, fp32 accumulate), one CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 256-bit global load / store (sm_100: LDG.E.256 / STG.E.256): one full 32 B sector per lane
__device__ __forceinline__ void ld_global_256(const float* ptr, float* dst) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]),
                 "=f"(dst[7])
               : "l"(ptr));
}
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  const uint32_t lo = (saddr >> 4) & 0x3FFF;                       // start address, LBO = 0 (unused, swizzled K-major)
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 B, version = 1 (sm_100), SWIZZLE_128B
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }


// ---- bulk-tensor stores (shared -> global) of the drain; bulk groups belong to the issuing thread
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// TMA-store views of a launch's outputs (DRAIN_TMA instantiations only): 5-D {channels, W, 1 | sub-pixel row, H, N} with a
// box of 16 channels x the 32 pixels one drain warp owns; fp32: SWIZZLE_64B, fp16 planes: SWIZZLE_32B
struct OutMaps {
  CUtensorMap f32, o0_hi, o0_lo, o1_hi, o1_lo;
};

constexpr int DRAIN_ROWS = 0;   // row per lane: direct / smem-transposed global stores (any shape)
constexpr int DRAIN_TMA = 2;    // row per lane, outputs staged in shared memory and written by bulk-tensor stores
// (1 was a quad-layout drain with full-line global accesses: measured, slower than DRAIN_TMA everywhere, removed --
//  profiles/r2_ncu_drain_variants.txt keeps its numbers)

// ---------------------------------------------------------------- epilogue
// TMA-store drain.
//
// ncu (profiles/r2_ncu_drain_variants.txt) shows what bounds the layers with fp32 outputs: the L1TEX data pipe, which also
// feeds the tensor core its shared-memory operands, is ~75 % busy, and a global STORE costs it one wavefront per 32 B
// sector (two in the row-per-lane layout) however well it coalesces: 8-11 k wavefronts per 128 x 128 tile beside the MMA's
// 8 k (3-pass).  A shared-memory store moves 128 B per wavefront.  So the drain keeps its row-per-lane layout, writes every
// 16-column batch of every output into the warp's 2 KB staging buffer (the same XOR pattern as the transposed drain =
// SWIZZLE_64B for fp32 rows of 64 B, SWIZZLE_32B for fp16 rows of 32 B) and one lane issues a bulk-tensor store of the
// 16 channels x 32 pixels box.  The buffer is single: a store must have been read out (wait_group.read) before the next
// batch is staged -- by then the warp has spent a batch's worth of math.  Image borders and phantom tiles need no
// masking: the TMA clips the box.
template <int PASSES, class PixFn, class WaitFn>
__device__ __forceinline__ void drain_tile_tma(const ConvArgs& p, const OutMaps* om, uint32_t t_acc, int bn, int ct, int cg,
                                               int q, int lane, uint32_t stage, PixFn pix, const float* bias_src,
                                               const float wscale, WaitFn wait_accumulator) {
  constexpr bool FAST = PASSES == 1;
  const bool shuffled = p.store == MCQ_STORE_SHUFFLE_NHWC;
  int n, oy, ox;
  const bool ok = pix(q * 32 + lane, n, oy, ox);
  const size_t pixoff = (((size_t)n * p.hout + oy) * p.wout + ox) * (size_t)p.cout;
  // box origin = the warp's first row (lane 0); the box covers the 32 rows in tile order (x, then y, then n)
  const int bx0 = __shfl_sync(0xffffffffu, ox, 0), by0 = __shfl_sync(0xffffffffu, oy, 0);
  const int bn0 = __shfl_sync(0xffffffffu, n, 0);
  const uint32_t row64 = stage + (uint32_t)lane * 64u;               // fp32: 64 B per row, chunk c at c ^ ((row >> 1) & 3)
  const uint32_t sw64 = (uint32_t)((lane >> 1) & 3);
  const uint32_t row32 = stage + (uint32_t)lane * 32u;               // fp16: 32 B per row, chunk c at c ^ ((row >> 2) & 1)
  const uint32_t sw32 = (uint32_t)((lane >> 2) & 1);
  const bool skip = p.debug_skip_store != 0;
  const float* src1 = (p.mode == MCQ_EPI_LINEAR || p.mode == MCQ_EPI_GATE) ? p.res1 : nullptr;
  const float* src2 = (p.mode == MCQ_EPI_LINEAR) ? p.res2 : p.aux;
  // this warp's 16-column batches: b -> first column (batches 2i, 2i + 1 are the halves of its i-th 32-column chunk)
  auto col_of = [&](int b) { return cg * 32 + (b >> 1) * 128 + (b & 1) * 16; };
  auto live = [&](int b) { const int c = col_of(b); return c < bn && ct * bn + c < p.cout; };   // warp-uniform, monotone
  auto off_of = [&](int b) {
    const int c0 = ct * bn + col_of(b);
    return shuffled ? epilogue_offset(p, n, oy, ox, c0) : pixoff + c0;
  };
  // The fp32 operands of a batch are requested one batch ahead, into the registers the previous batch has just consumed:
  // batch 0's before the accumulator wait, batch b + 1's as soon as batch b has added its operands -- its activation math
  // and its staging then run while the loads are in flight.
  // (one operand is requested ahead: the residual, or the GDN / IGDN operand.  Launches with two -- the attention gate,
  // `z - dequant + side` -- load the second one where it is used: 16 more live registers would spill the tile loop's state,
  // and a spill reload costs an L2 round trip here, the whole L1 being carved out as shared memory)
  const float* ahead = src1 ? src1 : src2;
  const float* late = src1 ? src2 : nullptr;
  float o1[16];
  auto request = [&](int b) {
    if (!ok || !ahead) return;
    const size_t off = off_of(b);
    ld_global_256(ahead + off, o1);
    ld_global_256(ahead + off + 8, o1 + 8);
  };
  if (live(0)) request(0);
  wait_accumulator();
#pragma unroll 1
  for (int b = 0; live(b); ++b) {
    const int col0 = col_of(b);
    const int c0 = ct * bn + col0;
    int cx = c0, cz = 0;                                             // channel / sub-pixel-row coordinates of the box
    if (shuffled) {
      const int cq = p.cout >> 2, sub = c0 / cq;
      cx = (sub & 1) * cq + (c0 - sub * cq);
      cz = sub >> 1;
    }
    uint32_t r[16];
    float y[16];
    tmem_ld16(t_acc + (uint32_t)col0, r);
    if (PASSES == 3) {
      uint32_t l[16];
      tmem_ld16(t_acc + (uint32_t)(bn + col0), l);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = fmaf(__uint_as_float(l[j]), kLoInv, __uint_as_float(r[j]));
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(r[j]);
    }
    {
      float bb[16];
      load_f32v<16>(bias_src, c0, bb);
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = fmaf(y[j], wscale, bb[j]);
    }
    if (ok) {                                                        // rows outside the image: values unused (clipped)
      if (p.mode == MCQ_EPI_LINEAR) {
        if (src1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = y[j] + p.res1_scale * o1[j];
          if (late) {
            float o2[16];
            const size_t off = off_of(b);
            ld_global_256(late + off, o2);
            ld_global_256(late + off + 8, o2 + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = y[j] + o2[j];
          }
        } else if (ahead) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = y[j] + o1[j];
        }
      } else if (p.mode == MCQ_EPI_GATE) {
        float o2[16];
        const size_t off = off_of(b);
        ld_global_256(late + off, o2);
        ld_global_256(late + off + 8, o2 + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = o2[j] * sigmoid_f<FAST>(y[j]) + o1[j];
      } else if (p.mode == MCQ_EPI_GDN) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (FAST) y[j] = o1[j] * rsqrtf(y[j]);
          else y[j] = o1[j] * (1.0f / sqrtf(y[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (FAST) y[j] = o1[j] * (y[j] * rsqrtf(y[j]));
          else y[j] = o1[j] * sqrtf(y[j]);
        }
      }
    }
    if (live(b + 1)) request(b + 1);                                 // o1 is dead from here on
    if (p.out_f32) {
      if (lane == 0) bulk_wait_read<0>();                            // the store that last read the staging buffer
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row64 + ((((uint32_t)c) ^ sw64) << 4)),
                     "f"(y[4 * c]), "f"(y[4 * c + 1]), "f"(y[4 * c + 2]), "f"(y[4 * c + 3])
                     : "memory");
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (!skip) tma_store_5d(&om->f32, stage, cx, bx0, cz, by0, bn0);
        bulk_commit();
      }
    }
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
      const bool on = slot == 0 ? p.o0_hi != nullptr : p.o1_hi != nullptr;
      const bool with_lo = slot == 0 ? p.o0_lo != nullptr : p.o1_lo != nullptr;
      if (!on) continue;
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      // (eight channels at a time: the activation / split temporaries of all sixteen at once are what pushed the tile
      //  loop's state out of the 96 registers, and a spill reload queues behind this warp's global loads)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float yy[8], t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] = y[8 * c + j];
        act_group<8, FAST>(yy, t, slot == 0 ? p.o0_act : p.o1_act);
        uint32_t h[4], lw[4];
        if (with_lo) {
#pragma unroll
          for (int j = 0; j < 4; ++j) split_f32x2(t[2 * j], t[2 * j + 1], h[j], lw[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = f2h2_sat(t[2 * j], t[2 * j + 1]);
        }
        const uint32_t a = row32 + ((((uint32_t)c) ^ sw32) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                     : "memory");
        if (with_lo)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 1024u), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]),
                       "r"(lw[3])
                       : "memory");
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (!skip) {
          tma_store_5d(slot == 0 ? &om->o0_hi : &om->o1_hi, stage, cx, bx0, cz, by0, bn0);
          if (with_lo) tma_store_5d(slot == 0 ? &om->o0_lo : &om->o1_lo, stage + 1024u, cx, bx0, cz, by0, bn0);
        }
        bulk_commit();
      }
    }
  }
}

// Row-per-lane drain with global stores (DRAIN_ROWS; any shape, argmin epilogue, GroupNorm statistics).
template <int PASSES, bool GN, class PixFn>
__device__ __forceinline__ void drain_tile_rows(const ConvArgs& p, uint32_t t_acc, int bn, int ct, int cg, int q, int lane,
                                                uint32_t stage, PixFn pix, const float* bias_src, const float wscale) {
  int n_[2], oy_[2], ox_[2];
  bool ok_[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) ok_[it] = pix(q * 32 + it * 16 + (lane >> 1), n_[it], oy_[it], ox_[it]);
  const int col8 = (lane & 1) * 8;
  if (p.mode == EPI_ARGMIN) {
    // tensor-core VQ: this thread owns GEMM row q*32+lane (one latent point) and scans its columns (codewords):
    // distance = (|x|^2 + |c_k|^2) - 2 x.c_k in the reference's order (quantizer.py:176); the running minimum is
    // merged across column groups / N tiles with one 64-bit atomicMin per thread: (distance, index) lexicographic,
    // i.e. the first index wins ties like torch.argmin.
    int n, oy, ox;
    const bool ok = pix(q * 32 + lane, n, oy, ox);
    const size_t point = ((size_t)n * p.hout + oy) * p.wout + ox;
    const float x2 = ok ? p.aux[point * p.argmin_stride] : 0.f;
    float best = INFINITY;
    int best_k = 0x7fffffff;
    for (int cc = cg * 32; cc < bn; cc += 128) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col0 = cc + half * 16;
        if (col0 >= bn) break;
        uint32_t r[16], l[16];
        tmem_ld16(t_acc + (uint32_t)col0, r);
        if (PASSES == 3) tmem_ld16(t_acc + (uint32_t)(bn + col0), l);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = ct * bn + col0 + j;
          float v = __uint_as_float(r[j]);
          if (PASSES == 3) v = fmaf(__uint_as_float(l[j]), kLoInv, v);
          v *= wscale;
          if (col < p.cout) {
            const float dist = (x2 + p.bias[col]) - 2.0f * v;
            if (dist < best) { best = dist; best_k = col; }
          }
        }
      }
    }
    if (ok && best_k != 0x7fffffff)
      atomicMin(p.argmin_keys + point * p.argmin_stride,
                ((unsigned long long)ordered_f32(best) << 32) | (unsigned long long)(uint32_t)best_k);
    return;
  }
  // ---- direct drain: no shared-memory transpose.  The lane keeps its GEMM row (= pixel) and moves 16 consecutive
  // channels per step with 256-bit accesses: one full 32 B sector per lane and plane, two per fp32 tensor.  A quarter of
  // the instructions of the transposed drain below (~800 per tile and warp), at the price of more L1TEX wavefronts per
  // byte (32 lines per instruction).  Used where instructions / energy are the scarce resource: plane-only outputs (first
  // conv of every ResidualBlock), 3-pass layers, 1x1 layers (64 B per lane in flight instead of 32 B for their HBM-bound
  // operand reads) -- and, measured, everywhere else too.  At 64x64: 1-pass plane->plane 64 -> 60 us (mainloop alone
  // 45), 3-pass 164 -> 157 us; whole step -2 %.
  const bool plane_only = p.mode == MCQ_EPI_LINEAR && !p.res1 && !p.res2 && !p.out_f32 && p.o0_hi && !p.o1_hi;
  // direct_epilogue (A/B knob, mcq_set_option("direct_epi")): >= 3 = as 1 here (the host picks DRAIN_TMA instantiations);
  // 1 = direct drain for every NHWC / PixelShuffle-NHWC store, 2 = plane-only outputs only, 0 = never
  const bool direct_ok = p.direct_epilogue == 1 || p.direct_epilogue >= 3 || (p.direct_epilogue == 2 && plane_only);
  const bool shuffled = p.store == MCQ_STORE_SHUFFLE_NHWC;
  if (direct_ok && ((p.store == MCQ_STORE_NHWC && p.cout % 16 == 0) || (shuffled && (p.cout >> 2) % 16 == 0))) {
    int n, oy, ox;
    const bool ok = pix(q * 32 + lane, n, oy, ox) && !p.debug_skip_store;
    const size_t pixoff = (((size_t)n * p.hout + oy) * p.wout + ox) * (size_t)p.cout;
    for (int cc = cg * 32; cc < bn; cc += 128) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col0 = cc + half * 16;
        if (col0 >= bn) break;
        const int c0 = ct * bn + col0;
        const bool live = ok && c0 < p.cout;
        const size_t off = shuffled ? epilogue_offset(p, n, oy, ox, c0) : pixoff + c0;
        uint32_t r[16];
        float y[16];
        tmem_ld16(t_acc + (uint32_t)col0, r);
        // fp32 operand (residual, or the GDN / IGDN operand): requested before the accumulator wait.  A second operand
        // (attention gate, `z - dequant + side`: rare) is loaded where it is used -- see drain_tile_tma.
        float o1[16];
        const float* src1 = (p.mode == MCQ_EPI_LINEAR || p.mode == MCQ_EPI_GATE) ? p.res1 : nullptr;
        const float* src2 = (p.mode == MCQ_EPI_LINEAR) ? p.res2 : p.aux;
        const float* ahead = src1 ? src1 : src2;
        const float* late = src1 ? src2 : nullptr;
        if (live && ahead) { ld_global_256(ahead + off, o1); ld_global_256(ahead + off + 8, o1 + 8); }
        if (PASSES == 3) {
          uint32_t l[16];
          tmem_ld16(t_acc + (uint32_t)(bn + col0), l);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaf(__uint_as_float(l[j]), kLoInv, __uint_as_float(r[j]));
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(r[j]);
        }
        float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};   // GN: this lane's sums per channel quad
        if (live) {
        {
          float b[16];
          load_f32v<16>(bias_src, c0, b);
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaf(y[j], wscale, b[j]);
        }
        if (p.mode == MCQ_EPI_LINEAR) {
          if (src1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = y[j] + p.res1_scale * o1[j];
            if (late) {
              float o2[16];
              ld_global_256(late + off, o2);
              ld_global_256(late + off + 8, o2 + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) y[j] = y[j] + o2[j];
            }
          } else if (ahead) {
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = y[j] + o1[j];
          }
        } else if (p.mode == MCQ_EPI_GATE) {
          float o2[16];
          ld_global_256(late + off, o2);
          ld_global_256(late + off + 8, o2 + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = o2[j] * sigmoid_f<PASSES == 1>(y[j]) + o1[j];
        } else if (p.mode == MCQ_EPI_GDN) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if constexpr (PASSES == 1) y[j] = o1[j] * rsqrtf(y[j]);
            else y[j] = o1[j] * (1.0f / sqrtf(y[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if constexpr (PASSES == 1) y[j] = o1[j] * (y[j] * rsqrtf(y[j]));
            else y[j] = o1[j] * sqrtf(y[j]);
          }
        }
        if (p.out_f32) {
          st_global_256(p.out_f32 + off, reinterpret_cast<const uint32_t(&)[8]>(y[0]));
          st_global_256(p.out_f32 + off + 8, reinterpret_cast<const uint32_t(&)[8]>(y[8]));
        }
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          __half* hi_p = slot == 0 ? p.o0_hi : p.o1_hi;
          __half* lo_p = slot == 0 ? p.o0_lo : p.o1_lo;
          if (!hi_p) continue;
          float t[16];
          act_group<16, PASSES == 1>(y, t, slot == 0 ? p.o0_act : p.o1_act);
          uint32_t h[8], lw[8];
          if (lo_p) {
#pragma unroll
            for (int j = 0; j < 8; ++j) split_f32x2(t[2 * j], t[2 * j + 1], h[j], lw[j]);
            st_global_256(hi_p + off, h);
            st_global_256(lo_p + off, lw);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = f2h2_sat(t[2 * j], t[2 * j + 1]);
            st_global_256(hi_p + off, h);
          }
        }
        if constexpr (GN) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            gs[u] = (y[4 * u] + y[4 * u + 1]) + (y[4 * u + 2] + y[4 * u + 3]);
            gq[u] = fmaf(y[4 * u + 3], y[4 * u + 3], fmaf(y[4 * u + 2], y[4 * u + 2],
                         fmaf(y[4 * u + 1], y[4 * u + 1], y[4 * u] * y[4 * u])));
          }
        }
        }   // live
        if constexpr (GN) {
          // GroupNorm partials of the fp32 output: (sum, sum of squares) per gn_unit channels over this warp's 32 pixels
          // (dead lanes contribute zeros), fixed reduction order -> bit-reproducible; one float2 store per unit.
          if (p.gn_ws != nullptr && c0 < p.cout) {          // warp-uniform
            const int unit = p.gn_unit;                    // 4, 8 or 16 channels
            if (unit >= 8) {
              gs[0] += gs[1]; gq[0] += gq[1];
              gs[1] = gs[2] + gs[3]; gq[1] = gq[2] + gq[3];
            }
            if (unit == 16) { gs[0] += gs[1]; gq[0] += gq[1]; }
            const int nu = 16 / unit;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (u < nu) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  gs[u] += __shfl_xor_sync(0xffffffffu, gs[u], o);
                  gq[u] += __shfl_xor_sync(0xffffffffu, gq[u], o);
                }
              }
            }
            const int n0 = __shfl_sync(0xffffffffu, n, 0), oy0 = __shfl_sync(0xffffffffu, oy, 0);
            const int ox0 = __shfl_sync(0xffffffffu, ox, 0);
            const bool ok0 = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;     // lane 0 = first pixel of the row block
            if (ok0 && lane < nu) {
              const int rb = (oy0 / (32 / p.tw)) * p.tiles_x + ox0 / p.tw;
              const float sv = lane == 0 ? gs[0] : lane == 1 ? gs[1] : lane == 2 ? gs[2] : gs[3];
              const float qv = lane == 0 ? gq[0] : lane == 1 ? gq[1] : lane == 2 ? gq[2] : gq[3];
              p.gn_ws[((size_t)n0 * p.gn_rb + rb) * p.gn_units + c0 / unit + lane] = make_float2(sv, qv);
            }
          }
        }
      }
    }
    return;
  }
  for (int cc = cg * 32; cc < bn; cc += 128) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col0 = cc + half * 16;
      if (col0 >= bn) break;
      uint32_t r[16];
      float v[16];
      tmem_ld16(t_acc + (uint32_t)col0, r);
      if (PASSES == 3) {
        uint32_t l[16];
        tmem_ld16(t_acc + (uint32_t)(bn + col0), l);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(l[j]), kLoInv, __uint_as_float(r[j]));
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);   // w_scale is applied with the bias (one FFMA)
      }
      // row-per-lane -> smem (64 B per row; 16 B chunk c stored at c ^ ((row >> 1) & 3))
      const uint32_t wbase = stage + (uint32_t)lane * 64u;
      const uint32_t wsw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wbase + ((((uint32_t)c) ^ wsw) << 4)),
                     "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                     : "memory");
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int rl = it * 16 + (lane >> 1);
        const uint32_t rbase = stage + (uint32_t)rl * 64u;
        const uint32_t rsw = (uint32_t)((rl >> 1) & 3);
        float vv[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t chunk = (uint32_t)((col8 >> 2) + c);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(vv[4 * c]), "=f"(vv[4 * c + 1]), "=f"(vv[4 * c + 2]), "=f"(vv[4 * c + 3])
                       : "r"(rbase + ((chunk ^ rsw) << 4))
                       : "memory");
        }
        if (ok_[it] && !p.debug_skip_store)
          epilogue_store<8, PASSES == 1>(p, n_[it], oy_[it], ox_[it], ct * bn + col0 + col8, vv, bias_src, wscale);
      }
      __syncwarp();
    }
  }
}

// DRAIN (chosen by the host, drain_kind() in mcq_api.cu): each variant is a separate kernel instantiation -- two drains in
// one kernel do not fit its 96 registers per thread.
// wait_accumulator(): blocks until the tile's accumulator is complete (mbarrier wait + tcgen05 fence); the bulk-store drain
// calls it after it has requested its first fp32 operands, the others on entry.
template <int PASSES, bool GN = false, int DRAIN = DRAIN_ROWS, class PixFn, class WaitFn>
__device__ __forceinline__ void drain_tile(const ConvArgs& p, const OutMaps* om, uint32_t t_acc, int bn, int ct, int cg,
                                           int q, int lane, uint32_t stage, PixFn pix, const float* bias_src,
                                           const float wscale, WaitFn wait_accumulator) {
  if constexpr (DRAIN == DRAIN_TMA) {
    static_assert(!GN, "the GroupNorm-statistics drain stores from registers");
    drain_tile_tma<PASSES>(p, om, t_acc, bn, ct, cg, q, lane, stage, pix, bias_src, wscale, wait_accumulator);
  } else {
    wait_accumulator();
    drain_tile_rows<PASSES, GN>(p, t_acc, bn, ct, cg, q, lane, stage, pix, bias_src, wscale);
  }
}

// The fp32 operands of the fused epilogue (residuals, GDN/gate operand) were written two or more launches ago and
// have left L2 at batch 64.  A drain warp reads them 32 B per lane with the loads right in front of their use, so one
// SM has ~16 KB in flight against ~1 us of DRAM latency: 16 GB/s per SM, a third of its HBM share.  Requesting the
// warp's 32 rows x 128 B into L2 before it blocks on the accumulator turns those loads into L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
template <class PixFn>
__device__ __forceinline__ void prefetch_epilogue_operands(const ConvArgs& p, int bn, int ct, int cg, int q, int lane,
                                                           PixFn pix) {
  if (p.store != MCQ_STORE_NHWC || p.mode == EPI_ARGMIN) return;
  if (!p.res1 && !p.res2 && !p.aux) return;
  int n, oy, ox;
  if (!pix(q * 32 + lane, n, oy, ox)) return;
  const size_t pixoff = (((size_t)n * p.hout + oy) * p.wout + ox) * p.cout;
  for (int cc = cg * 32; cc < bn; cc += 128) {
    const int c = ct * bn + cc;
    if (c >= p.cout) break;
    if (p.res1) prefetch_l2(p.res1 + pixoff + c);
    if (p.res2) prefetch_l2(p.res2 + pixoff + c);
    if (p.aux) prefetch_l2(p.aux + pixoff + c);
  }
}

// ---------------------------------------------------------------- kernel
template <int PASSES, int DRAIN = DRAIN_ROWS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const ConvArgs p, const __grid_constant__ OutMaps om) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16 B aligned: round up to the 1024 B the 128B swizzle needs
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = p.bn;
  const uint32_t b_bytes = (uint32_t)bn * TC_BK * 2;
  const uint32_t stage_bytes = (TC_A_BYTES + b_bytes) * (PASSES == 3 ? 2u : 1u);
  const int stages = p.stages;

  // smem carve-up: [stage][A_hi | A_lo | B_hi | B_lo] ... barriers
  const uint32_t bar_base = smem_base + stage_bytes * stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * stages + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + stage_bytes * stages + 8u * (2 * stages + 4));
  const uint32_t epi_base = (bar_base + 8u * (2 * stages + 4) + 16u + 511u) & ~511u;   // 16 x 2 KB staging buffers (512 B: swizzle period)
  float* bias_smem = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base) + TC_EPI_WARPS * TC_EPI_STAGE_BYTES);
  const bool bias_staged = (p.mode != EPI_ARGMIN) && (p.cout <= TC_BIAS_SMEM_FLOATS);
  if (bias_staged)
    for (int i = threadIdx.x; i < p.cout; i += TC_THREADS) bias_smem[i] = p.bias[i];

  const int acc_cols = (PASSES == 3 ? 2 : 1) * bn;   // TMEM columns per accumulator buffer
  const int nbuf = (2 * acc_cols <= (int)TC_TMEM_COLS) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (PASSES == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), TC_EPI_WARPS);  // one arrive per epilogue warp
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
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias staging
  // of constant weights) overlapped the tail of the previous kernel in the stream; from here on we read its outputs.
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  const int tiles_m = p.tiles_n * p.tiles_y * p.tiles_x;
  const int total_tiles = tiles_m * p.tiles_c;
  const int ntaps = p.ksize * p.ksize;
  // cin need not fill its last 64-channel K chunk: TMA zero-fills the A box beyond the tensor's channel extent, and
  // whatever the weight box holds there (the next tap's columns, or zeros past the last one) is multiplied by those zeros
  const int kchunks = (p.cin + TC_BK - 1) / TC_BK;
  const int kiters = ntaps * kchunks;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ct = t / tiles_m;
        int mt = t - ct * tiles_m;
        const int bx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int by = mt % p.tiles_y;
        const int bz = mt / p.tiles_y;
        const int x0 = bx * p.tw, y0 = by * p.th, n0 = bz * p.tn, c0 = ct * bn;
        for (int tap = 0; tap < ntaps; ++tap) {
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(empty_bar(s), ph ^ 1u, 1, p.wait_sleep_ns);
            const uint32_t sa = smem_base + stage_bytes * s;
            const int ca = p.ch_off + p.tap_c[tap] + kc * TC_BK;
            const int kb = tap * p.cin + kc * TC_BK;
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), stage_bytes);
              tma_load_5d(&tmA_hi, sa, full_bar(s), ca, x0 + p.tap_dx[tap], p.tap_py[tap], y0 + p.tap_dy[tap], n0);
              if (PASSES == 3) {
                tma_load_5d(&tmA_lo, sa + TC_A_BYTES, full_bar(s), ca, x0 + p.tap_dx[tap], p.tap_py[tap],
                            y0 + p.tap_dy[tap], n0);
                tma_load_2d(&tmB_hi, sa + 2 * TC_A_BYTES, full_bar(s), kb, c0);
                tma_load_2d(&tmB_lo, sa + 2 * TC_A_BYTES + b_bytes, full_bar(s), kb, c0);
              } else {
                tma_load_2d(&tmB_hi, sa + TC_A_BYTES, full_bar(s), kb, c0);
              }
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc_wide = (1u << 4) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const bool wide = (PASSES == 3) && (2 * bn <= 256) && (bn % 8 == 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int buf = (nbuf == 2) ? (it & 1) : 0;
        const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_hh = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t d_lo = d_hh + (uint32_t)bn;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full_bar(s), ph, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + stage_bytes * s;
          const uint64_t a_hi = make_sdesc(sa);
          const uint64_t a_lo = make_sdesc(sa + TC_A_BYTES);
          const uint64_t b_hi = make_sdesc(sa + (PASSES == 3 ? 2 : 1) * TC_A_BYTES);
          const uint64_t b_lo = make_sdesc(sa + 2 * TC_A_BYTES + b_bytes);
          const uint32_t acc0 = ki > 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint32_t acc = k > 0 ? 1u : acc0;
              const uint64_t ko = (uint64_t)(k * 2);  // +32 B per K=16 step, in 16 B units
              if (PASSES == 3) {
                if (wide) {
                  // B_hi and B_lo are adjacent in smem and acc_hh / acc_lo adjacent in TMEM: one N = 2*bn MMA computes
                  // [A_hi*B_hi | A_hi*B_lo] and reads A_hi once (operand fetch from smem is the scarce resource)
                  umma_f16(d_hh, a_hi + ko, b_hi + ko, idesc_wide, acc);
                } else {
                  umma_f16(d_hh, a_hi + ko, b_hi + ko, idesc, acc);
                  umma_f16(d_lo, a_hi + ko, b_lo + ko, idesc, acc);
                }
                umma_f16(d_lo, a_lo + ko, b_hi + ko, idesc, 1u);
              } else {
                umma_f16(d_hh, a_hi + ko, b_hi + ko, idesc, acc);
              }
            }
            umma_commit(empty_bar(s));  // frees this smem stage once the MMAs above have read it
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(tfull_bar(buf));  // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access (hardware rule: warp % 4)
    const int cg = (warp - 2) >> 2;     // column group: this warp handles 32-column chunks cg, cg+4, ...
    const uint32_t stage = epi_base + (uint32_t)(warp - 2) * TC_EPI_STAGE_BYTES;
    // GEMM row inside a tile = pixel (x fastest, then y, then n)
    auto tile_pix = [&](int t, int& ct) {
      ct = t / tiles_m;
      int mt = t - ct * tiles_m;
      const int bx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int by = mt % p.tiles_y;
      const int bz = mt / p.tiles_y;
      return [=, &p](int row, int& n, int& oy, int& ox) {
        ox = bx * p.tw + row % p.tw;
        oy = by * p.th + (row / p.tw) % p.th;
        n = bz * p.tn + row / (p.tw * p.th);
        return (ox < p.wout) && (oy < p.hout) && (n < p.n);
      };
    };
    int it = 0;
    const float wscale = effective_w_scale(p);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int ct;
      auto pix = tile_pix(t, ct);
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
      // L2 prefetch of the fp32 operands: this tile's on the first trip, from then on the NEXT tile's (a whole tile of lead)
      // (bulk-store launches prefetch nothing: measured, their operand read is bound by the L1TEX wavefronts of its loads, not
      //  by latency -- profiles/r2_drain_ab.txt)
      if constexpr (DRAIN != DRAIN_TMA) {
        if (it == 0) prefetch_epilogue_operands(p, bn, ct, cg, q, lane, pix);
        if (t + (int)gridDim.x < total_tiles) {
          int ct2;
          auto pix2 = tile_pix(t + (int)gridDim.x, ct2);
          prefetch_epilogue_operands(p, bn, ct2, cg, q, lane, pix2);
        }
      }
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols);
      drain_tile<PASSES, false, DRAIN>(p, &om, t_acc, bn, ct, cg, q, lane, stage, pix, bias_staged ? bias_smem : p.bias, wscale,
                                       [&]() {
                                         mbar_wait(tfull_bar(buf), use & 1u, 4, p.wait_sleep_ns);
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
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

}  // namespace mcq
