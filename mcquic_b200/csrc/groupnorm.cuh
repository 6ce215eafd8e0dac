// nn.GroupNorm(groups, C) between the two convolutions of ResidualBlock(denseNorm=True) (mcquic/nn/blocks.py:198):
// statistics + affine normalisation + split into the next convolution's A-operand planes in ONE launch.
//
// One thread-block cluster works on one image at a time: CTA `rank` owns a slice of the image's pixels.  Phase 1 reduces
// the slice to per-group (sum, sum of squares) -- fp32 per thread over <= a few hundred values, every combination step
// after that in double and in a fixed order (bit-reproducible, no atomics).  The cluster's CTAs then read each other's
// partials through distributed shared memory, and phase 2 re-reads the slice and writes
// y = x * (rstd * gamma) + (beta - mean * rstd * gamma), the form PyTorch's CPU kernel evaluates.
// The grid is persistent (clusters loop over images) and sized by the host so that the images in flight fit in L2:
// the phase-2 read is then an L2 hit and HBM traffic = the fp32 activation once in, the planes (and/or fp32) once out
// (outputs use streaming stores so that they do not evict the slices still waiting for their second read).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace mcq {

struct GroupNormArgs {
  const float* x;       // [n, hw, c] fp32 NHWC
  const float* gamma;   // [c]
  const float* beta;    // [c]
  float* out_f32;       // optional
  __half* o_hi;         // optional split-fp16 planes of act(y)
  __half* o_lo;
  int o_act;
  int n, hw, c, groups;
  float eps;
};

constexpr int GN_MAX_C = 512;
constexpr int GN_MAX_CLUSTER = 8;
constexpr int gn_smem_bytes(int threads) { return GN_MAX_C * (8 + 8 + 16 + 8) + threads * 4 * 4 * 2; }

template <int GN_THREADS>
__global__ void __launch_bounds__(GN_THREADS) groupnorm_kernel(const GroupNormArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char gn_smem[];
  double2* part = reinterpret_cast<double2*>(gn_smem);            // per-group (sum, sumsq) of this CTA's slice: read by the cluster
  double* ch_s = reinterpret_cast<double*>(part + GN_MAX_C);      // per-channel sums of this CTA's slice
  double* ch_q = ch_s + GN_MAX_C;
  float2* stat = reinterpret_cast<float2*>(ch_q + GN_MAX_C);      // per-group (mean, rstd) of the whole image
  float* row_s = reinterpret_cast<float*>(stat + GN_MAX_C);       // [rows][c] per-thread partials, rows * c <= 4 * GN_THREADS
  float* row_q = row_s + 4 * GN_THREADS;

  const int slices = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int num_clusters = gridDim.x / slices;
  const int c4 = a.c >> 2;
  const int rows = GN_THREADS / c4;                   // pixels handled per sweep
  const int t = threadIdx.x;
  const bool active = t < rows * c4;
  const int q = t % c4, r = t / c4;
  const int per = (a.hw + slices - 1) / slices;
  const int p0 = rank * per, p1 = min(a.hw, p0 + per);
  const int cg_ = a.c / a.groups;                     // channels per group
  const int warp = t >> 5, lane = t & 31;
  const double cnt = (double)a.hw * (double)cg_;
  float gam[4] = {0.f, 0.f, 0.f, 0.f}, bet[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gam[j] = a.gamma[4 * q + j];
      bet[j] = a.beta[4 * q + j];
    }
  }

  for (int img = blockIdx.x / slices; img < a.n; img += num_clusters) {
    const float* xin = a.x + (size_t)img * a.hw * a.c + 4 * q;

    // ---- phase 1: slice -> per-group partial sums
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll 8
      for (int p = p0 + r; p < p1; p += rows) {
        const float4 v = *reinterpret_cast<const float4*>(xin + (size_t)p * a.c);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]);
        ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        row_s[r * a.c + 4 * q + j] = s[j];
        row_q[r * a.c + 4 * q + j] = ss[j];
      }
    }
    __syncthreads();
    for (int ch = t; ch < a.c; ch += GN_THREADS) {
      double S = 0.0, Q = 0.0;
      for (int rr = 0; rr < rows; ++rr) {
        S += (double)row_s[rr * a.c + ch];
        Q += (double)row_q[rr * a.c + ch];
      }
      ch_s[ch] = S;
      ch_q[ch] = Q;
    }
    __syncthreads();
    for (int g = warp; g < a.groups; g += GN_THREADS / 32) {
      double S = 0.0, Q = 0.0;
      for (int i = lane; i < cg_; i += 32) {
        S += ch_s[g * cg_ + i];
        Q += ch_q[g * cg_ + i];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        Q += __shfl_xor_sync(0xffffffffu, Q, o);
      }
      if (lane == 0) part[g] = make_double2(S, Q);
    }
    cluster.sync();   // every CTA's `part` is complete and visible cluster-wide

    // ---- image statistics: the slices' partials summed in rank order (the same order in every CTA)
    for (int g = t; g < a.groups; g += GN_THREADS) {
      double S = 0.0, Q = 0.0;
      for (int rk = 0; rk < slices; ++rk) {
        const double2* remote = cluster.map_shared_rank(part, rk);
        const double2 v = remote[g];
        S += v.x;
        Q += v.y;
      }
      const double mean = S / cnt;
      double var = Q / cnt - mean * mean;               // biased variance, as nn.GroupNorm
      if (var < 0.0) var = 0.0;
      stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)a.eps)));
    }
    __syncthreads();

    // ---- phase 2: normalise, affine, activation, split
    if (active) {
      float sc[4], sh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 st = stat[(4 * q + j) / cg_];
        sc[j] = st.y * gam[j];
        sh[j] = fmaf(-sc[j], st.x, bet[j]);
      }
      const size_t base = (size_t)img * a.hw * a.c + 4 * q;
#pragma unroll 4
      for (int p = p0 + r; p < p1; p += rows) {
        const float4 v = *reinterpret_cast<const float4*>(xin + (size_t)p * a.c);
        float y[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
        const size_t off = base + (size_t)p * a.c;
        if (a.out_f32) __stcs(reinterpret_cast<float4*>(a.out_f32 + off), make_float4(y[0], y[1], y[2], y[3]));
        if (a.o_hi) {
          float tv[4];
          act_group<4, false>(y, tv, a.o_act);
          uint32_t h[2], l[2];
          if (a.o_lo) {
            split_f32x2(tv[0], tv[1], h[0], l[0]);
            split_f32x2(tv[2], tv[3], h[1], l[1]);
            __stcs(reinterpret_cast<uint2*>(a.o_hi + off), make_uint2(h[0], h[1]));
            __stcs(reinterpret_cast<uint2*>(a.o_lo + off), make_uint2(l[0], l[1]));
          } else {
            h[0] = f2h2_sat(tv[0], tv[1]);
            h[1] = f2h2_sat(tv[2], tv[3]);
            __stcs(reinterpret_cast<uint2*>(a.o_hi + off), make_uint2(h[0], h[1]));
          }
        }
      }
    }
    cluster.sync();   // `part` / `stat` / the row buffers may be overwritten only after every CTA has read them
  }
}

// ---- GroupNorm from statistics the producing convolution's epilogue already accumulated (drain_tile<.., GN = true>):
// partials float2 [n][rb][units] = (sum y, sum y^2) per (32-pixel row block, `unit` consecutive channels).
// gn_finalize_kernel: one CTA per (group, image) sums its partials in a fixed order in double -> stats[n][groups] =
// (mean, rstd).  gn_apply_kernel: one streaming pass, y = x * (rstd * gamma) + (beta - mean * rstd * gamma).
struct GroupNormApplyArgs {
  const float* x;
  const float2* partials;
  float2* stats;
  const float* gamma;
  const float* beta;
  float* out_f32;
  __half* o_hi;
  __half* o_lo;
  int o_act;
  int n, hw, c, groups, rb, unit;
  float eps;
};

constexpr int GNF_THREADS = 256;

__global__ void __launch_bounds__(GNF_THREADS) gn_finalize_kernel(const GroupNormApplyArgs a) {
  __shared__ double red_s[GNF_THREADS], red_q[GNF_THREADS];
  const int g = blockIdx.x, img = blockIdx.y;
  const int cg_ = a.c / a.groups;
  const int upg = cg_ / a.unit;                      // units per group
  const int units = a.c / a.unit;
  const float2* src = a.partials + (size_t)img * a.rb * units + (size_t)g * upg;
  double S = 0.0, Q = 0.0;
  const int total = a.rb * upg;
  for (int i = threadIdx.x; i < total; i += GNF_THREADS) {
    const float2 v = src[(size_t)(i / upg) * units + (i % upg)];
    S += (double)v.x;
    Q += (double)v.y;
  }
  red_s[threadIdx.x] = S;
  red_q[threadIdx.x] = Q;
  __syncthreads();
  for (int o = GNF_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red_s[threadIdx.x] += red_s[threadIdx.x + o];
      red_q[threadIdx.x] += red_q[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cnt = (double)a.hw * (double)cg_;
    const double mean = red_s[0] / cnt;
    double var = red_q[0] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    a.stats[(size_t)img * a.groups + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)a.eps)));
  }
}

// One block = `ppb` consecutive pixels of one image (blockIdx.y); a thread keeps one channel quad, so its scale / shift
// are computed once and the loop body is a load, four FFMAs, the split and the stores (32-bit index arithmetic only:
// the first version spent 160 instructions per float4 on 64-bit div/mod and was issue-bound at 3.8 TB/s).
constexpr int GNA_THREADS = 256;

__global__ void __launch_bounds__(GNA_THREADS) gn_apply_kernel(const GroupNormApplyArgs a, int ppb) {
  const int img = blockIdx.y;
  const int c4 = a.c >> 2;
  const int rows = GNA_THREADS / c4;
  const int t = threadIdx.x;
  if (t >= rows * c4) return;
  const int q = t % c4, r = t / c4;
  const int ch = 4 * q;
  const float2 st = a.stats[(size_t)img * a.groups + ch / (a.c / a.groups)];   // channels per group % 4 == 0
  const float4 gm = *reinterpret_cast<const float4*>(a.gamma + ch);
  const float4 bt = *reinterpret_cast<const float4*>(a.beta + ch);
  const float sc[4] = {st.y * gm.x, st.y * gm.y, st.y * gm.z, st.y * gm.w};
  const float sh[4] = {fmaf(-sc[0], st.x, bt.x), fmaf(-sc[1], st.x, bt.y), fmaf(-sc[2], st.x, bt.z),
                       fmaf(-sc[3], st.x, bt.w)};
  const int p0 = blockIdx.x * ppb, p1 = min(a.hw, p0 + ppb);
  const size_t base = (size_t)img * a.hw * a.c + ch;
#pragma unroll 4
  for (int p = p0 + r; p < p1; p += rows) {
    const size_t off = base + (size_t)p * a.c;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(a.x + off));
    float y[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
    if (a.o_hi) store_planes<4>(a.o_hi, a.o_lo, off, y, a.o_act);
  }
}

}  // namespace mcq
