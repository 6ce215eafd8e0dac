// Shared device-side definitions: kernel parameter block, split-fp16 helpers, the fused epilogue.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcquic_b200.h"

namespace mcq {

constexpr float kLoScale = 2048.0f;          // lo plane = (a - hi) * 2^11
constexpr float kLoInv = 1.0f / 2048.0f;
constexpr int EPI_ARGMIN = 4;   // internal epilogue mode of the tensor-core VQ (vq_assign_tc): bias = |c|^2, aux = |x|^2

// monotone map float -> uint32 (total order incl. negatives), so that (key << 32 | index) orders by (distance, index)
__device__ __forceinline__ uint32_t ordered_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Device view of one convolution launch (mcq_conv_params + derived geometry).
struct ConvArgs {
  const __half* a_hi;
  const __half* a_lo;
  const __half* w_hi;
  const __half* w_lo;
  const float* bias;
  const float* res1;
  const float* res2;
  const float* aux;
  float* out_f32;
  unsigned char* out_u8;          // MCQ_STORE_SHUFFLE_NCHW only: uint8 pixels (DeTransform, utils/vision.py:135-146) instead of fp32
  __half* o0_hi;
  __half* o0_lo;
  __half* o1_hi;
  __half* o1_lo;
  float w_scale;
  float res1_scale;
  const float* dev_scale;        // optional device scalar multiplied into w_scale (loss-scale removal of the backward convs)
  int n, hin, win, cin;
  int hout, wout;  // conv output grid (before any pixel shuffle)
  int cout, cout_pad, ksize, stride, ktotal;
  int mode, store, o0_act, o1_act, passes;
  // tensor-core tiling: a tile is a (tw x th x tn) box of output pixels (x fastest), tw*th*tn == 128
  int tw, th, tn;
  int tiles_x, tiles_y, tiles_n;  // tiles along W, H, N
  int bn;                         // N tile (GEMM columns per CTA tile)
  int tiles_c;                    // cout_pad / bn
  int stages;
  int cin_total, ch_off;          // A is the channel slice [ch_off, ch_off + cin) of a [.., cin_total] tensor
  unsigned long long* argmin_keys;  // EPI_ARGMIN: per-point packed (ordered distance << 32 | index) minima
  int argmin_stride;              // points are argmin_stride keys / aux entries apart (= number of codebooks)
  int debug_skip_store;           // MCQ_EPI_SKIP=1: drain TMEM but store nothing (profiling aid)
  int wait_sleep_ns;              // nanosleep between mbarrier polls of the producer / drain warps (MCQ_WAIT_SLEEP_NS)
  int direct_epilogue;            // mcq_set_option("direct_epi"): which drain a launch takes, see drain_kind() (mcq_api.cu) and
                                  // drain_tile_rows (conv_tc.cuh); default 3
  // GroupNorm statistics fused into the drain (pair kernel, GN instantiation only): every drain warp writes, per
  // gn_unit consecutive channels, (sum y, sum y^2) over its 32 pixels to gn_ws[(n * gn_rb + row block) * gn_units + unit]
  float2* gn_ws;
  int gn_unit, gn_units, gn_rb;
  // per-tap TMA coordinate offsets in the 5-D view of A (see conv_tc.cuh)
  int tap_c[9], tap_dx[9], tap_py[9], tap_dy[9];
};

__device__ __forceinline__ float effective_w_scale(const ConvArgs& p) {
  return p.dev_scale ? p.w_scale * __ldg(p.dev_scale) : p.w_scale;
}

// Activation math.  Both variants use the SFU (ex2.approx, rcp.approx); the exact variant adds one Newton step to the
// reciprocal so that sigmoid/SiLU carry <= ~2 ulp error (the exponential's 2^-22), i.e. fp32-grade like the rest of
// the 3-pass path, at a quarter of the instructions of expf() + IEEE division.  FAST (1-pass, TF32-grade) skips it.
__device__ __forceinline__ float rcp_approx(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// FAST (1-pass path, whose operands are rounded to fp16 anyway): one SFU op, sigmoid(x) = 0.5 + 0.5 tanh(x / 2) with
// tanh.approx's 2^-11 relative error -- 3 instructions instead of 6 and half the SFU traffic of the ex2 + rcp form.
template <bool FAST>
__device__ __forceinline__ float sigmoid_f(float x) {
  if constexpr (FAST) {
    return fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f);
  } else {
    const float d = 1.0f + __expf(fminf(-x, 80.0f));   // clamp keeps d finite (inf * 0 in the Newton step would be NaN)
    float r = rcp_approx(d);
    r = r * fmaf(-d, r, 2.0f);
    return r;
  }
}
template <bool FAST>
__device__ __forceinline__ float silu_f(float x) {
  if constexpr (FAST) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);                   // x sigmoid(x) = h + h tanh(h)
  } else {
    return x * sigmoid_f<false>(x);
  }
}

template <bool FAST>
__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == MCQ_ACT_SILU) return silu_f<FAST>(y);
  if (act == MCQ_ACT_SQUARE) return y * y * MCQ_SQUARE_SCALE;
  return y;
}

__device__ __forceinline__ unsigned short f2h_sat(float a) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float h2f(unsigned short h) { return __half2float(__ushort_as_half(h)); }

// a ~= hi + lo / 2048, both fp16
__device__ __forceinline__ void split_f32(float a, unsigned short& hi, unsigned short& lo) {
  hi = f2h_sat(a);
  lo = f2h_sat((a - h2f(hi)) * kLoScale);
}

// two values at once: packed fp16x2 words (element 0 in the low half)
__device__ __forceinline__ uint32_t f2h2_sat(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
__device__ __forceinline__ void split_f32x2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = f2h2_sat(e0, e1);
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = f2h2_sat((e0 - back.x) * kLoScale, (e1 - back.y) * kLoScale);
}

// The activation selector is launch-uniform: branch on it ONCE per group of NV values, not per element (the per-element
// form cost two compare+branch pairs per value and plane -- a third of all instructions the drain warps executed).
template <int NV, bool FAST>
__device__ __forceinline__ void act_group(const float (&y)[NV], float (&t)[NV], int act) {
  if (act == MCQ_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] = silu_f<FAST>(y[j]);
  } else if (act == MCQ_ACT_SQUARE) {
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] = y[j] * y[j] * MCQ_SQUARE_SCALE;
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) t[j] = y[j];
  }
}

template <int NV, bool FAST = false>
__device__ __forceinline__ void store_planes(__half* hi_p, __half* lo_p, size_t off, const float (&y)[NV], int act) {
  static_assert(NV == 4 || NV == 8, "NV");
  uint32_t h[NV / 2], l[NV / 2];
  float t[NV];
  act_group<NV, FAST>(y, t, act);
  if (lo_p) {
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) split_f32x2(t[2 * j], t[2 * j + 1], h[j], l[j]);
    if constexpr (NV == 8) {
      *reinterpret_cast<uint4*>(hi_p + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo_p + off) = make_uint4(l[0], l[1], l[2], l[3]);
    } else {
      *reinterpret_cast<uint2*>(hi_p + off) = make_uint2(h[0], h[1]);
      *reinterpret_cast<uint2*>(lo_p + off) = make_uint2(l[0], l[1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) h[j] = f2h2_sat(t[2 * j], t[2 * j + 1]);
    if constexpr (NV == 8) {
      *reinterpret_cast<uint4*>(hi_p + off) = make_uint4(h[0], h[1], h[2], h[3]);
    } else {
      *reinterpret_cast<uint2*>(hi_p + off) = make_uint2(h[0], h[1]);
    }
  }
}

template <int NV>
__device__ __forceinline__ void load_f32v(const float* p, size_t off, float (&r)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; j += 4) {
    float4 t = *reinterpret_cast<const float4*>(p + off + j);
    r[j] = t.x; r[j + 1] = t.y; r[j + 2] = t.z; r[j + 3] = t.w;
  }
}

// DeTransform of the reference (mcquic/utils/vision.py:135-146): [-1, 1] -> uint8, the same fp32 operations in the same
// order: ((x + 1) / 2 * 255.999).clamp(0, 255).byte()  (.byte() truncates)
__device__ __forceinline__ unsigned char detransform_u8(float x) {
  float t = __fdiv_rn(x - (-1.0f), 2.0f) * 255.999f;
  t = fminf(fmaxf(t, 0.0f), 255.0f);
  return (unsigned char)(int)t;
}

// Output element offset of GEMM columns [c0, ..) of conv pixel (n, oy, ox) for the NHWC stores
// (PixelShuffle convs: column (2i+j)*C/4 + c of pixel (oy, ox) is channel c of pixel (2oy+i, 2ox+j)).
__device__ __forceinline__ size_t epilogue_offset(const ConvArgs& p, int n, int oy, int ox, int c0) {
  int C = p.cout, H = p.hout, W = p.wout, py = oy, px = ox, c = c0;
  if (p.store == MCQ_STORE_SHUFFLE_NHWC) {
    const int cq = p.cout >> 2;
    const int sub = c0 / cq;
    c = c0 - sub * cq;
    py = 2 * oy + (sub >> 1);
    px = 2 * ox + (sub & 1);
    H *= 2; W *= 2; C = cq;
  }
  return (((size_t)n * H + py) * W + px) * C + c;
}

// Fused epilogue for NV consecutive GEMM columns [c0, c0+NV) of output pixel (n, oy, ox).
// v[] * scale = accumulator * w_scale (bias NOT yet added; scale is a power of two, so folding it into the bias FFMA
// rounds exactly like multiplying first).  c0 % NV == 0.
// bias_src: where to read the bias from (the tensor-core kernels keep a copy in shared memory: with the whole
// carve-out given to smem there is no L1, and a global bias fetch per call is a ~300-cycle stall)
template <int NV, bool FAST = false>
__device__ __forceinline__ void epilogue_store(const ConvArgs& p, int n, int oy, int ox, int c0, float (&v)[NV],
                                               const float* bias_src = nullptr, float scale = 1.0f) {
  if (c0 >= p.cout) return;
  if (bias_src == nullptr) bias_src = p.bias;
  if (p.store == MCQ_STORE_SHUFFLE_NCHW) {
    // last layer (compressor.py:139): GEMM column 4c+2i+j -> out[n, c, 2oy+i, 2ox+j], fp32 NCHW
    const int cq = p.cout >> 2, H2 = p.hout * 2, W2 = p.wout * 2;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = c0 + j;
      if (col < p.cout) {
        const int c = col >> 2, i = (col >> 1) & 1, jj = col & 1;
        const size_t o = (((size_t)n * cq + c) * H2 + (2 * oy + i)) * W2 + (2 * ox + jj);
        const float y = fmaf(v[j], scale, bias_src[col]);
        if (p.out_u8) p.out_u8[o] = detransform_u8(y);
        else p.out_f32[o] = y;
      }
    }
    return;
  }
  int C = p.cout, H = p.hout, W = p.wout, py = oy, px = ox, c = c0;
  if (p.store == MCQ_STORE_SHUFFLE_NHWC) {
    const int cq = p.cout >> 2;
    const int sub = c0 / cq;
    c = c0 - sub * cq;
    py = 2 * oy + (sub >> 1);
    px = 2 * ox + (sub & 1);
    H *= 2; W *= 2; C = cq;
  }
  const size_t off = (((size_t)n * H + py) * W + px) * C + c;
  float b[NV], y[NV];
  load_f32v<NV>(bias_src, c0, b);
#pragma unroll
  for (int j = 0; j < NV; ++j) y[j] = fmaf(v[j], scale, b[j]);
  if (p.mode == MCQ_EPI_LINEAR) {
    if (p.res1) {
      float r[NV];
      load_f32v<NV>(p.res1, off, r);
#pragma unroll
      for (int j = 0; j < NV; ++j) y[j] = y[j] + p.res1_scale * r[j];
    }
    if (p.res2) {
      float r[NV];
      load_f32v<NV>(p.res2, off, r);
#pragma unroll
      for (int j = 0; j < NV; ++j) y[j] = y[j] + r[j];
    }
  } else if (p.mode == MCQ_EPI_GATE) {
    float r[NV], a[NV];
    load_f32v<NV>(p.res1, off, r);
    load_f32v<NV>(p.aux, off, a);
#pragma unroll
    for (int j = 0; j < NV; ++j) y[j] = a[j] * sigmoid_f<FAST>(y[j]) + r[j];
  } else {
    float a[NV];
    load_f32v<NV>(p.aux, off, a);
    if (p.mode == MCQ_EPI_GDN) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if constexpr (FAST) y[j] = a[j] * rsqrtf(y[j]);
        else y[j] = a[j] * (1.0f / sqrtf(y[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if constexpr (FAST) y[j] = a[j] * (y[j] * rsqrtf(y[j]));
        else y[j] = a[j] * sqrtf(y[j]);
      }
    }
  }
  if (p.out_f32) {
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      *reinterpret_cast<float4*>(p.out_f32 + off + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
  }
  if (p.o0_hi) store_planes<NV, FAST>(p.o0_hi, p.o0_lo, off, y, p.o0_act);
  if (p.o1_hi) store_planes<NV, FAST>(p.o1_hi, p.o1_lo, off, y, p.o1_act);
}

}  // namespace mcq
