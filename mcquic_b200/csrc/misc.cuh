// Boundary kernels: the RGB stem convolution (with AlignedPadding's reflect pad folded in), layout
// conversions between the reference's NCHW fp32 tensors and the internal NHWC / split-fp16 planes.
#pragma once
#include "common.cuh"

namespace mcq {

struct StemArgs {
  const float* x;   // [n, 3, h, w] fp32, or
  const unsigned char* x_u8;   // uint8: value = (u / 255 - 0.5) * 2 (mcquic/demo.py:110-118)
  const float* w;   // [cout, 27]
  const float* bias;
  float* out_f32;
  __half *o_hi, *o_lo;
  int o_act;
  int n, h, w_, pad_top, pad_left, hp, wp, cout, hout, wout;
};

// reflect index as F.pad(mode="reflect") does (mcquic/data/transforms.py:99)
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// One block = STEM_PIX consecutive output pixels of one output row x all cout.  The 3 x (2*STEM_PIX+1) x 3 input
// patch (reflect pad / zero pad resolved while loading) and the transposed weights sit in shared memory; a warp owns
// STEM_PIX/4 pixels, lane l the 4 output channels [4l, 4l+4) (+128 per further channel group): its 108 weights stay
// in registers, the 27 inputs of a pixel are warp-uniform broadcasts, and every store is 512 contiguous bytes.
constexpr int STEM_PIX = 64;
constexpr int STEM_THREADS = 128;   // 156 registers/thread (108 weights): 3 blocks per SM

__global__ void __launch_bounds__(STEM_THREADS) stem_conv_kernel(const StemArgs a) {
  extern __shared__ float stem_smem[];
  float* ws = stem_smem;                         // [27][cout]
  float* patch = stem_smem + 27 * a.cout;        // [3 ch][3 rows][2*STEM_PIX + 1]
  constexpr int PW = 2 * STEM_PIX + 1;
  for (int i = threadIdx.x; i < 27 * a.cout; i += STEM_THREADS) {
    const int co = i % a.cout, t = i / a.cout;
    ws[t * a.cout + co] = a.w[co * 27 + t];
  }
  const int strips = (a.wout + STEM_PIX - 1) / STEM_PIX;
  const int strip = blockIdx.x % strips;
  const int oy = (blockIdx.x / strips) % a.hout;
  const int n = blockIdx.x / (strips * a.hout);
  const int ox0 = strip * STEM_PIX;
  for (int i = threadIdx.x; i < 9 * PW; i += STEM_THREADS) {
    const int col = i % PW, rr = (i / PW) % 3, ci = i / (3 * PW);
    const int Y = 2 * oy + rr - 1, X = 2 * ox0 + col - 1;   // coordinates in the padded image (zeros outside it)
    float v = 0.f;
    if (Y >= 0 && Y < a.hp && X >= 0 && X < a.wp) {
      const int sy = reflect_idx(Y - a.pad_top, a.h), sx = reflect_idx(X - a.pad_left, a.w_);
      const size_t idx = (((size_t)n * 3 + ci) * a.h + sy) * a.w_ + sx;
      v = a.x_u8 ? (__fdiv_rn((float)a.x_u8[idx], 255.0f) - 0.5f) * 2.0f : a.x[idx];
    }
    patch[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = lane * 4; c0 < a.cout; c0 += 128) {
    float4 w[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) w[t] = *reinterpret_cast<const float4*>(ws + t * a.cout + c0);
    const float4 b = *reinterpret_cast<const float4*>(a.bias + c0);
#pragma unroll 2
    for (int pp = 0; pp < STEM_PIX / (STEM_THREADS / 32); ++pp) {
      const int px = warp * (STEM_PIX / (STEM_THREADS / 32)) + pp;
      const int ox = ox0 + px;
      if (ox >= a.wout) break;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 27; ++t) {      // t = ci*9 + r*3 + s  (nn.Conv2d weight order)
        const float v = patch[(t / 9) * 3 * PW + ((t / 3) % 3) * PW + 2 * px + (t % 3)];
        acc[0] = fmaf(v, w[t].x, acc[0]);
        acc[1] = fmaf(v, w[t].y, acc[1]);
        acc[2] = fmaf(v, w[t].z, acc[2]);
        acc[3] = fmaf(v, w[t].w, acc[3]);
      }
      float y[4] = {acc[0] + b.x, acc[1] + b.y, acc[2] + b.z, acc[3] + b.w};
      const size_t off = (((size_t)n * a.hout + oy) * a.wout + ox) * a.cout + c0;
      if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
      if (a.o_hi) store_planes<4>(a.o_hi, a.o_lo, off, y, a.o_act);
    }
  }
}

__global__ void split_planes_kernel(const float* x, long long count4, int act, __half* hi, __half* lo,
                                    const float* dev_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float y[4];
  load_f32v<4>(x, (size_t)i * 4, y);
  if (dev_scale) {
    const float s = __ldg(dev_scale);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] *= s;
  }
  store_planes<4>(hi, lo, (size_t)i * 4, y, act);
}

// out = x + alpha * y (one rounding: alpha = +-1 gives exactly the reference's `latent - currentLatent` /
// `quantized + formerLevel`, mcquic/modules/quantizer.py:686,701), as fp32 and/or split planes of act(out)
__global__ void add_scaled_kernel(const float* x, const float* y, float alpha, long long count4, float* out_f32,
                                  int act, __half* hi, __half* lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float a[4], b[4], v[4];
  load_f32v<4>(x, (size_t)i * 4, a);
  load_f32v<4>(y, (size_t)i * 4, b);
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = fmaf(alpha, b[j], a[j]);
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  if (hi) store_planes<4>(hi, lo, (size_t)i * 4, v, act);
}

struct LayoutArgs {
  const float* x;
  float* out_f32;
  __half *o0_hi, *o0_lo, *o1_hi, *o1_lo;
  int o0_act, o1_act;
  int n, c, hw;
};

// NCHW fp32 -> NHWC (fp32 and/or planes) through a 32x32 smem transpose; block (32, 8)
__global__ void nchw_to_nhwc_kernel(const LayoutArgs a) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < a.c && p < a.hw) ? a.x[((size_t)n * a.c + c) * a.hw + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < a.hw && c < a.c) {
      const float y = tile[threadIdx.x][j];
      const size_t off = ((size_t)n * a.hw + p) * a.c + c;
      if (a.out_f32) a.out_f32[off] = y;
      if (a.o0_hi) {
        unsigned short h, l;
        split_f32(apply_act<false>(y, a.o0_act), h, l);
        a.o0_hi[off] = __ushort_as_half(h);
        if (a.o0_lo) a.o0_lo[off] = __ushort_as_half(l);
      }
      if (a.o1_hi) {
        unsigned short h, l;
        split_f32(apply_act<false>(y, a.o1_act), h, l);
        a.o1_hi[off] = __ushort_as_half(h);
        if (a.o1_lo) a.o1_lo[off] = __ushort_as_half(l);
      }
    }
  }
}

__global__ void nhwc_to_nchw_kernel(const float* x, int c_, int hw, float* out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < c_ && p < hw) ? x[((size_t)n * hw + p) * c_ + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < c_ && p < hw) out[((size_t)n * c_ + c) * hw + p] = tile[threadIdx.x][j];
  }
}

}  // namespace mcq
