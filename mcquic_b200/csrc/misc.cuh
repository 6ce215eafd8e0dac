// Boundary kernels: the RGB stem convolution (with AlignedPadding's reflect pad folded in), layout
// conversions between the reference's NCHW fp32 tensors and the internal NHWC / split-fp16 planes.
#pragma once
#include "common.cuh"

namespace mcq {

struct StemArgs {
  const float* x;   // [n, 3, h, w]
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

// block = (cout/4 threads for channels) x (pixels): each thread 4 output channels of one pixel
__global__ void __launch_bounds__(256) stem_conv_kernel(const StemArgs a) {
  extern __shared__ float ws[];  // [27][cout] transposed weights
  for (int i = threadIdx.x; i < 27 * a.cout; i += blockDim.x) {
    const int co = i % a.cout, t = i / a.cout;
    ws[t * a.cout + co] = a.w[co * 27 + t];
  }
  __syncthreads();
  const int cg = a.cout / 4;
  const int pix_per_block = blockDim.x / cg;
  const int lp = threadIdx.x / cg, c0 = (threadIdx.x % cg) * 4;
  const long long pix = (long long)blockIdx.x * pix_per_block + lp;
  const long long total = (long long)a.n * a.hout * a.wout;
  if (lp >= pix_per_block || pix >= total) return;
  const int ox = (int)(pix % a.wout);
  const long long t = pix / a.wout;
  const int oy = (int)(t % a.hout);
  const int n = (int)(t / a.hout);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int Y = 2 * oy + r - 1;  // coordinate in the padded image (zero padding outside it)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int X = 2 * ox + s - 1;
        float v = 0.f;
        if (Y >= 0 && Y < a.hp && X >= 0 && X < a.wp) {
          const int sy = reflect_idx(Y - a.pad_top, a.h), sx = reflect_idx(X - a.pad_left, a.w_);
          v = a.x[(((size_t)n * 3 + ci) * a.h + sy) * a.w_ + sx];
        }
        const float4 wv = *reinterpret_cast<const float4*>(ws + (ci * 9 + r * 3 + s) * a.cout + c0);
        acc[0] = fmaf(v, wv.x, acc[0]);
        acc[1] = fmaf(v, wv.y, acc[1]);
        acc[2] = fmaf(v, wv.z, acc[2]);
        acc[3] = fmaf(v, wv.w, acc[3]);
      }
    }
  }
  float y[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) y[j] = acc[j] + a.bias[c0 + j];
  const size_t off = (size_t)pix * a.cout + c0;
  if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
  if (a.o_hi) store_planes<4>(a.o_hi, a.o_lo, off, y, a.o_act);
}

__global__ void split_planes_kernel(const float* x, long long count4, int act, __half* hi, __half* lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float y[4];
  load_f32v<4>(x, (size_t)i * 4, y);
  store_planes<4>(hi, lo, (size_t)i * 4, y, act);
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
