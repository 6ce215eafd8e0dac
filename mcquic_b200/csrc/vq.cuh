// Multi-codebook vector quantizer kernels.
//
// vq_assign_kernel: one pass over the latent grid that computes, per point and codebook, the squared-L2
// distance to all K codewords in the reference's operation order  (|x|^2 + |c_k|^2) - 2 x.c_k
// (mcquic/modules/quantizer.py:176), the first-index argmin (:148) and optionally the soft logits
// (:181-183,204) and the code histogram -- without materialising the [n,m,h,w,k] distance tensor.
// Codebook tiles are staged in shared memory; the latent tile is read once with coalesced float4 loads.
#pragma once
#include "common.cuh"

namespace mcq {

constexpr int VQ_TP = 64;        // points per block
constexpr int VQ_TK = 128;       // codewords per shared-memory tile
constexpr int VQ_THREADS = 256;  // 16 (k groups of 8) x 16 (point groups of 4)
constexpr int VQ_MAX_D = 128;

struct VqArgs {
  const float* x;         // [P, m*d] NHWC
  const float* codebook;  // [m, k, d]
  const float* c2;        // [m, k]
  long long* codes;       // [n, m, hw]
  float* logits;          // [n, m, hw, k] or null
  const float* logit_scale;  // [m] or null
  int* hist;              // [m, k] or null
  int P, hw, m, k, d;
  float inv_sqrt_k;
};

__global__ void __launch_bounds__(VQ_THREADS) vq_assign_kernel(const VqArgs a) {
  extern __shared__ __align__(16) float vq_smem[];
  const int d = a.d;
  float* Xs = vq_smem;                   // [d][VQ_TP]  (transposed: d-major)
  float* Cs = Xs + (size_t)d * VQ_TP;    // [d][VQ_TK]
  float* X2 = Cs + (size_t)d * VQ_TK;    // [VQ_TP]

  const int tid = threadIdx.x;
  const int mi = blockIdx.y;
  const int p0 = blockIdx.x * VQ_TP;
  const int C = a.m * d;

  // ---- stage the latent tile (coalesced along channels), transposed into Xs[d][p]
  for (int i = tid; i < VQ_TP * (d / 4); i += VQ_THREADS) {
    const int pp = i % VQ_TP, j4 = (i / VQ_TP) * 4;   // lanes -> consecutive points: conflict-free smem stores
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + pp < a.P) v = *reinterpret_cast<const float4*>(a.x + (size_t)(p0 + pp) * C + mi * d + j4);
    Xs[(j4 + 0) * VQ_TP + pp] = v.x;
    Xs[(j4 + 1) * VQ_TP + pp] = v.y;
    Xs[(j4 + 2) * VQ_TP + pp] = v.z;
    Xs[(j4 + 3) * VQ_TP + pp] = v.w;
  }
  __syncthreads();
  if (tid < VQ_TP) {
    float s = 0.f;
    for (int j = 0; j < d; ++j) {
      const float v = Xs[j * VQ_TP + tid];
      s = fmaf(v, v, s);
    }
    X2[tid] = s;
  }

  const int tk = tid & 15;   // codeword group: tile columns tk*8 .. tk*8+7
  const int tp = tid >> 4;   // point group:    points tp*4 .. tp*4+3
  float best[4];
  int best_k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; best_k[i] = 0x7fffffff; }
  const float lscale = a.logits ? -a.inv_sqrt_k * (a.logit_scale ? a.logit_scale[mi] : 1.0f) : 0.f;

  for (int k0 = 0; k0 < a.k; k0 += VQ_TK) {
    __syncthreads();  // previous tile fully consumed (also orders X2 writes before first use)
    // ---- stage codebook tile transposed: Cs[d][kk]
    for (int i = tid; i < VQ_TK * (d / 4); i += VQ_THREADS) {
      const int kk = i % VQ_TK, j4 = (i / VQ_TK) * 4;  // lanes -> consecutive codewords
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < a.k) v = *reinterpret_cast<const float4*>(a.codebook + ((size_t)mi * a.k + k0 + kk) * d + j4);
      Cs[(j4 + 0) * VQ_TK + kk] = v.x;
      Cs[(j4 + 1) * VQ_TK + kk] = v.y;
      Cs[(j4 + 2) * VQ_TK + kk] = v.z;
      Cs[(j4 + 3) * VQ_TK + kk] = v.w;
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int j = 0; j < d; ++j) {
      const float4 xv = *reinterpret_cast<const float4*>(Xs + j * VQ_TP + tp * 4);
      const float4 c0 = *reinterpret_cast<const float4*>(Cs + j * VQ_TK + tk * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(Cs + j * VQ_TK + tk * 8 + 4);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      const float ca[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(xa[i], ca[jj], acc[i][jj]);
    }
    // ---- distance in the reference's order, running first-index argmin, optional logits
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int kg = k0 + tk * 8 + jj;
      if (kg < a.k) {
        const float c2 = a.c2[(size_t)mi * a.k + kg];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dist = (X2[tp * 4 + i] + c2) - 2.0f * acc[i][jj];
          if (dist < best[i]) { best[i] = dist; best_k[i] = kg; }   // strict <: keeps the lowest index in-thread
          acc[i][jj] = dist;
        }
      }
    }
    if (a.logits) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pg = p0 + tp * 4 + i;
        if (pg < a.P) {
          const int nn = pg / a.hw, pix = pg - nn * a.hw;
          float* dst = a.logits + (((size_t)nn * a.m + mi) * a.hw + pix) * a.k + k0 + tk * 8;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (k0 + tk * 8 + jj < a.k) dst[jj] = acc[i][jj] * lscale;
        }
      }
    }
  }
  // ---- reduce (distance, index) lexicographically over the 16 codeword groups (lanes differing in low 4 bits)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best[i], off);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k[i], off);
      if (ob < best[i] || (ob == best[i] && ok < best_k[i])) { best[i] = ob; best_k[i] = ok; }
    }
    const int pg = p0 + tp * 4 + i;
    if (tk == 0 && pg < a.P) {
      const int nn = pg / a.hw, pix = pg - nn * a.hw;
      a.codes[((size_t)nn * a.m + mi) * a.hw + pix] = best_k[i];
      if (a.hist) atomicAdd(a.hist + (size_t)mi * a.k + best_k[i], 1);
    }
  }
}

// ---- tensor-core VQ helpers (vq_assign_tc): the x.c_k GEMM runs as a 1x1 tcgen05 'convolution' (3-pass split-fp16)
// whose epilogue does the distance + argmin (conv_tc.cuh, EPI_ARGMIN).
// prep: fp32 latents -> split-fp16 planes, |x|^2 per (point, codebook) summed in channel order, keys = +inf.
__global__ void vq_prep_kernel(const float* x, int P, int m, int d, __half* hi, __half* lo, float* x2,
                               unsigned long long* keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * m) return;
  const int pt = i / m, mi = i - pt * m;
  const size_t base = (size_t)pt * m * d + (size_t)mi * d;
  float s = 0.f;
  for (int j = 0; j < d; j += 4) {
    float y[4];
    load_f32v<4>(x, base + j, y);
#pragma unroll
    for (int q = 0; q < 4; ++q) s = fmaf(y[q], y[q], s);
    store_planes<4>(hi, lo, base + j, y, MCQ_ACT_NONE);
  }
  x2[i] = s;
  keys[i] = 0xFFFFFFFFFFFFFFFFull;
}

// finalize: packed minima -> int64 codes [n, m, hw] (+ histogram)
__global__ void vq_finalize_kernel(const unsigned long long* keys, int P, int hw, int m, int k, long long* codes,
                                   int* hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * m) return;
  const int pt = i / m, mi = i - pt * m;
  const int code = (int)(keys[i] & 0xFFFFFFFFull);
  const int nn = pt / hw, pix = pt - nn * hw;
  codes[((size_t)nn * m + mi) * hw + pix] = code;
  if (hist && code >= 0 && code < k) atomicAdd(hist + (size_t)mi * k + code, 1);
}

// codes [n, m, hw] -> codebook rows, NHWC [n, hw, m*d]; one thread per 4 channels
struct DequantArgs {
  const long long* codes;
  const float* codebook;
  float* out_f32;
  __half *o0_hi, *o0_lo, *o1_hi, *o1_lo;
  int o0_act, o1_act;
  int* status;
  int P, hw, m, k, d;
};

__global__ void vq_dequant_kernel(const DequantArgs a) {
  const int C = a.m * a.d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.P * (C / 4);
  if (i >= total) return;
  const int pg = (int)(i / (C / 4));
  const int c = (int)(i % (C / 4)) * 4;
  const int mi = c / a.d, j = c - mi * a.d;
  const int nn = pg / a.hw, pix = pg - nn * a.hw;
  long long code = a.codes[((size_t)nn * a.m + mi) * a.hw + pix];
  if (code < 0 || code >= a.k) {
    if (a.status) atomicExch(a.status, MCQ_ERR_CODE_RANGE);
    code = 0;
  }
  float y[4];
  load_f32v<4>(a.codebook, ((size_t)mi * a.k + code) * a.d + j, y);
  const size_t off = (size_t)pg * C + c;
  if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
  if (a.o0_hi) store_planes<4>(a.o0_hi, a.o0_lo, off, y, a.o0_act);
  if (a.o1_hi) store_planes<4>(a.o1_hi, a.o1_lo, off, y, a.o1_act);
}

__global__ void code_histogram_kernel(const long long* codes, int n, int m, int hw, int k, int* hist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * m * hw;
  if (i >= total) return;
  const int mi = (int)((i / hw) % m);
  const long long code = codes[i];
  if (code >= 0 && code < k) atomicAdd(hist + (size_t)mi * k + code, 1);
}

}  // namespace mcq
