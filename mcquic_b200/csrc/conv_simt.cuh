// CUDA-core (fp32 FFMA) implicit-GEMM convolution over the same split-fp16 planes and the same fused
// epilogue as the tcgen05 kernel.  It is the exact-fp32 cross-check for the tensor-core path
// (MCQ_IMPL_SIMT) and serves channel counts the tensor-core tiling does not (cin % 64 != 0).
#pragma once
#include "common.cuh"

namespace mcq {

constexpr int SIMT_TM = 64;   // output pixels per block
constexpr int SIMT_TN = 64;   // GEMM columns per block
constexpr int SIMT_TK = 16;   // K chunk
constexpr int SIMT_THREADS = 256;

__global__ void __launch_bounds__(SIMT_THREADS) conv_simt_kernel(const ConvArgs p) {
  __shared__ float As[SIMT_TK][SIMT_TM + 4];
  __shared__ float Bs[SIMT_TK][SIMT_TN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // column group: columns tx*4 .. tx*4+3
  const int ty = tid >> 4;   // pixel group:  pixels  ty*4 .. ty*4+3
  const long long M = (long long)p.n * p.hout * p.wout;
  const long long m0 = (long long)blockIdx.x * SIMT_TM;
  const int n0 = blockIdx.y * SIMT_TN;

  // loader mapping: 64 rows x 16 k -> each thread 4 consecutive k of one row
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;

  // pixel handled by this thread as loader
  const long long lm = m0 + lrow;
  int ln = 0, loy = 0, lox = 0;
  const bool lvalid = lm < M;
  if (lvalid) {
    lox = (int)(lm % p.wout);
    long long t = lm / p.wout;
    loy = (int)(t % p.hout);
    ln = (int)(t / p.hout);
  }
  const int pad = p.ksize / 2;
  const bool use_lo = (p.passes == 3) && p.a_lo != nullptr;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = p.ksize * p.ksize;
  for (int tap = 0; tap < ntaps; ++tap) {
    const int r = tap / p.ksize, s = tap % p.ksize;
    const int iy = loy * p.stride + r - pad, ix = lox * p.stride + s - pad;
    const bool in_ok = lvalid && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win;
    const size_t a_off = in_ok ? (((size_t)ln * p.hin + iy) * p.win + ix) * p.cin : 0;
    for (int c0 = 0; c0 < p.cin; c0 += SIMT_TK) {
      // A tile
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (in_ok && c0 + lk < p.cin) {
        const uint2 h = *reinterpret_cast<const uint2*>(p.a_hi + a_off + c0 + lk);
        av[0] = h2f(h.x & 0xffff); av[1] = h2f(h.x >> 16); av[2] = h2f(h.y & 0xffff); av[3] = h2f(h.y >> 16);
        if (use_lo) {
          const uint2 l = *reinterpret_cast<const uint2*>(p.a_lo + a_off + c0 + lk);
          av[0] += h2f(l.x & 0xffff) * kLoInv; av[1] += h2f(l.x >> 16) * kLoInv;
          av[2] += h2f(l.y & 0xffff) * kLoInv; av[3] += h2f(l.y >> 16) * kLoInv;
        }
      }
      // B tile: row = GEMM column n0 + lrow
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (n0 + lrow < p.cout_pad && c0 + lk < p.cin) {
        const size_t b_off = (size_t)(n0 + lrow) * p.ktotal + (size_t)tap * p.cin + c0 + lk;
        const uint2 h = *reinterpret_cast<const uint2*>(p.w_hi + b_off);
        bv[0] = h2f(h.x & 0xffff); bv[1] = h2f(h.x >> 16); bv[2] = h2f(h.y & 0xffff); bv[3] = h2f(h.y >> 16);
        if (p.passes == 3 && p.w_lo) {
          const uint2 l = *reinterpret_cast<const uint2*>(p.w_lo + b_off);
          bv[0] += h2f(l.x & 0xffff) * kLoInv; bv[1] += h2f(l.x >> 16) * kLoInv;
          bv[2] += h2f(l.y & 0xffff) * kLoInv; bv[3] += h2f(l.y >> 16) * kLoInv;
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lk + j][lrow] = av[j];
        Bs[lk + j][lrow] = bv[j];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SIMT_TK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int ox = (int)(m % p.wout);
    const long long t = m / p.wout;
    const int oy = (int)(t % p.hout);
    const int n = (int)(t / p.hout);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] * effective_w_scale(p);
    epilogue_store<4>(p, n, oy, ox, n0 + tx * 4, v);
  }
}

}  // namespace mcq
