// C ABI of libmcquic_b200.so (see include/mcquic_b200.h).  Host-side launch logic only: argument
// validation, tile geometry, TMA tensor maps, kernel launches on the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "conv_simt.cuh"
#include "conv_halo.cuh"
#include "conv_pair.cuh"
#include "conv_tc.cuh"
#include "conv_chain.cuh"
#include "groupnorm.cuh"
#include "misc.cuh"
#include "vq.cuh"
#include "vq_fused.cuh"
#include "stem_tc.cuh"
#include "conv_wgrad.cuh"

using namespace mcq;

namespace {

std::atomic<int> g_launches{0};
long long* g_chain_dbg = nullptr;   // mcq_debug_timeline(): device buffer for the chain kernel's clock samples
// optional per-launch timing events of the current mcq_conv2d call (thread-local: the library is re-entrant per thread)
thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
struct EvScope {
  cudaStream_t st;
  explicit EvScope(cudaStream_t s) : st(s) { if (g_ev_start) cudaEventRecord(g_ev_start, st); }
  ~EvScope() { if (g_ev_stop) cudaEventRecord(g_ev_stop, st); }
};

#define MCQ_CHECK_ARG(cond)            \
  do {                                 \
    if (!(cond)) return MCQ_ERR_BAD_ARG; \
  } while (0)

inline int cuda_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// ---- driver entry point for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Tuning / debugging knobs.  They are plain process-wide values with fixed defaults, changed ONLY through the explicit
// mcq_set_option() call (tools/ and tests use it for A/B measurements); nothing on the launch path reads the environment.
struct Option { const char* name; int value; };
Option g_options[] = {
    {"direct_epi", 3},        // drain (drain_kind()): 3 = bulk-tensor stores for large launches, else direct; 5 = bulk-tensor
                              // stores wherever they apply; 1 = direct (row per lane) for every NHWC / PixelShuffle store,
                              // 2 = direct for plane-only outputs, 0 = smem-transposed
    {"wait_sleep_ns", 0},     // nanosleep between mbarrier polls of the producer / drain warps
    {"tc_spread", 33},        // conv_tc: narrow the N tile until this % of the SMs have a tile
    {"pdl", 1},               // programmatic dependent launch
    {"epi_skip", 0},          // profiling aid: drain TMEM but store nothing
    {"pair", -1},             // CTA-pair kernel: -1 = auto (3-pass and weight-resident 1-pass layers), 0 / 1 = force
    {"pair_resident", 1},     // weight-stationary 1-pass mode of the pair kernel
    {"pair_96", 1},           // pair kernel with 96-column N tiles for 3-pass layers of 96 k output channels (C = 192 models)
    {"pair_narrow", 1},       // pair kernel also for layers of 16 / 32 / 64 (padded) output channels
    {"pair_nbs", 0},          // A/B knob: cap on the weight stages of the streaming (3-pass) mode; 0 = all that fit
    {"halo", 1},              // single-CTA halo kernel for 3x3 stride-1 layers the pair kernel does not take
    {"halo_resident", 1},     // its weight-stationary mode (single N tile, all weight stages fit)
    {"halo_pitch", 10}, {"halo_base", 0}, {"halo_cl", 2}, {"halo_tps", 0}, {"halo_nbs", 0},
    {"small_grid_pct", 100},  // grid cap (% of the SMs) of launches with fewer work items than clusters
    {"chain", 1}, {"chain_ipc", 0}, {"chain_nosync", 0}, {"chain_sync_mode", 0},
    {"gn_threads", 256}, {"gn_ctas_per_sm", 4}, {"gn_apply_iters", 8}, {"gn_fused", 1},
    {"vq128_fused", 1},       // one-launch tcgen05 VQ for d = 128 (qp = 1); 0 = prep + GEMM-epilogue argmin + finalize
};
int* find_option(const char* name) {
  for (auto& o : g_options)
    if (std::strcmp(o.name, name) == 0) return &o.value;
  return nullptr;
}
int opt(const char* name) {
  const int* v = find_option(name);
  return v ? *v : 0;
}

int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// Which drain a tensor-core convolution launch runs (conv_tc.cuh; one kernel instantiation per kind).  `work_per_cta`:
// tiles each CTA walks -- a launch with less than two is latency-bound (small maps) and keeps the plain drain, whose
// stores need no staging round trip and no bulk-group wait before the kernel can end.
//   direct_epi = 3 (default): DRAIN_TMA for large launches that can take it;  5: DRAIN_TMA wherever it applies;
//   0 / 1 / 2: DRAIN_ROWS (see the option table)
int drain_kind(const ConvArgs& a, long long work_per_cta) {
  if (a.direct_epilogue < 3 || a.mode == EPI_ARGMIN || a.gn_ws) return DRAIN_ROWS;
  const int cper = a.store == MCQ_STORE_NHWC ? a.cout : (a.store == MCQ_STORE_SHUFFLE_NHWC ? a.cout >> 2 : 0);
  if (cper == 0) return DRAIN_ROWS;
  if (a.direct_epilogue == 4) return DRAIN_ROWS;   // (was the quad-layout drain: removed)
  if (a.bn % 16 != 0 || cper % 16 != 0) return DRAIN_ROWS;
  if (a.direct_epilogue == 3 && work_per_cta < 2) return DRAIN_ROWS;
  // N tiles below 64 columns leave 12 of the 16 drain warps without a column: the few that work then serialise on their
  // single staging buffer (measured: the 32-channel Neon training step 163 -> 167 ms with the bulk-store drain)
  if (a.direct_epilogue == 3 && a.bn < 64) return DRAIN_ROWS;
  // bulk-tensor stores need 16-byte aligned tensors (torch allocations are; a caller of the C ABI may pass anything)
  const uintptr_t ptrs = (uintptr_t)a.out_f32 | (uintptr_t)a.o0_hi | (uintptr_t)a.o0_lo | (uintptr_t)a.o1_hi | (uintptr_t)a.o1_lo;
  if (ptrs & 15) return DRAIN_ROWS;
  return DRAIN_TMA;
}

// TMA-store view of one output tensor for DRAIN_TMA: 5-D {channels, W, 1 | sub-pixel row, H, N} over the conv's output
// grid, box = 16 channels x the 32 consecutive tile rows of one drain warp (tile rows run x, then y, then n)
int encode_out_map(CUtensorMap* map, const void* ptr, bool f32, const ConvArgs& a) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  const cuuint64_t es = f32 ? 4 : 2;
  const int bw = a.tw < 32 ? a.tw : 32;
  const int bh = a.th < 32 / bw ? a.th : 32 / bw;
  const int bnn = 32 / (bw * bh);
  cuuint64_t dims[5], strides[4];
  if (a.store == MCQ_STORE_NHWC) {
    const cuuint64_t c = (cuuint64_t)a.cout, w = (cuuint64_t)a.wout, h = (cuuint64_t)a.hout;
    dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = (cuuint64_t)a.n;
    strides[0] = c * es; strides[1] = w * c * es; strides[2] = w * c * es; strides[3] = h * w * c * es;
  } else {
    // PixelShuffle: GEMM column (2i + j) * cq + c of conv pixel (oy, ox) is channel c of output pixel (2 oy + i, 2 ox + j)
    const cuuint64_t cq = (cuuint64_t)(a.cout >> 2), w = (cuuint64_t)a.wout, h = (cuuint64_t)a.hout, w2 = 2 * w;
    dims[0] = 2 * cq; dims[1] = w; dims[2] = 2; dims[3] = h; dims[4] = (cuuint64_t)a.n;
    strides[0] = 2 * cq * es; strides[1] = w2 * cq * es; strides[2] = 2 * w2 * cq * es; strides[3] = 2 * h * w2 * cq * es;
  }
  cuuint32_t box[5] = {16u, (cuuint32_t)bw, 1u, (cuuint32_t)bh, (cuuint32_t)bnn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MCQ_ERR_DRIVER;
}
int encode_out_maps(OutMaps& om, const ConvArgs& a) {
  int rc = 0;
  if (a.out_f32) rc = encode_out_map(&om.f32, a.out_f32, true, a);
  if (!rc && a.o0_hi) rc = encode_out_map(&om.o0_hi, a.o0_hi, false, a);
  if (!rc && a.o0_lo) rc = encode_out_map(&om.o0_lo, a.o0_lo, false, a);
  if (!rc && a.o1_hi) rc = encode_out_map(&om.o1_hi, a.o1_hi, false, a);
  if (!rc && a.o1_lo) rc = encode_out_map(&om.o1_lo, a.o1_lo, false, a);
  return rc;
}

struct TcPlan {
  ConvArgs args;
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  size_t smem_bytes;
  int grid;
};

constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
int num_sms() {   // per device ordinal (a process may drive several GPUs)
  static std::atomic<int> sms[kMaxDevices];
  const int dev = current_device();
  int v = sms[dev].load(std::memory_order_relaxed);
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: opt in once per (kernel, device),
// raising it when a later launch needs more.  One slot per kernel template instantiation (the `Tag` type).
template <class Tag, class K>
cudaError_t ensure_dyn_smem(K kernel, size_t bytes) {
  static std::atomic<size_t> have[kMaxDevices];
  const int dev = current_device();
  if (have[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have[dev].store(bytes, std::memory_order_release);
  return e;
}
template <int> struct KTag {};

// Encode the 5-D view of an NHWC fp16 plane that makes every filter tap a unit-stride TMA box.
//   stride 1: dims {C,  W,   1, H,   N}
//   stride 2: dims {2C, W/2, 2, H/2, N}   (x parity folded into the channel axis, y parity its own axis)
int encode_act_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int stride, int tw, int th, int tn) {
  // c = channels of the tensor in memory (the kernel may read a slice of them)
  // box = {64 channels, tw, 1, th, tn}
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  cuuint64_t dims[5], strides[4];
  const cuuint64_t es = 2;
  if (stride == 1) {
    dims[0] = c; dims[1] = w; dims[2] = 1; dims[3] = h; dims[4] = n;
    strides[0] = (cuuint64_t)c * es;
    strides[1] = (cuuint64_t)w * c * es;
    strides[2] = (cuuint64_t)w * c * es;
    strides[3] = (cuuint64_t)h * w * c * es;
  } else {
    dims[0] = 2 * c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = n;
    strides[0] = (cuuint64_t)2 * c * es;
    strides[1] = (cuuint64_t)w * c * es;
    strides[2] = (cuuint64_t)2 * w * c * es;
    strides[3] = (cuuint64_t)h * w * c * es;
  }
  cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)tw, 1u, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MCQ_ERR_DRIVER;
}

int encode_weight_map(CUtensorMap* map, const void* ptr, int cout_pad, int ktotal, int bn) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)ktotal, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)ktotal * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MCQ_ERR_DRIVER;
}

int fill_args(const mcq_conv_params* p, ConvArgs& a) {
  MCQ_CHECK_ARG(p && p->a_hi && p->w_hi && p->bias);
  MCQ_CHECK_ARG(p->n > 0 && p->hin > 0 && p->win > 0 && p->cin > 0 && p->cout > 0);
  MCQ_CHECK_ARG(p->ksize == 1 || p->ksize == 3);
  MCQ_CHECK_ARG(p->stride == 1 || p->stride == 2);
  MCQ_CHECK_ARG(p->passes == 1 || p->passes == 3);
  MCQ_CHECK_ARG(p->cin % 4 == 0);
  MCQ_CHECK_ARG(p->cout_pad >= p->cout);
  if (p->passes == 3) MCQ_CHECK_ARG(p->a_lo && p->w_lo);
  if (p->stride == 2) MCQ_CHECK_ARG(p->hin % 2 == 0 && p->win % 2 == 0);
  MCQ_CHECK_ARG(p->mode >= MCQ_EPI_LINEAR && p->mode <= MCQ_EPI_IGDN);
  MCQ_CHECK_ARG(p->store >= MCQ_STORE_NHWC && p->store <= MCQ_STORE_SHUFFLE_NCHW);
  if (p->mode == MCQ_EPI_GATE) MCQ_CHECK_ARG(p->res1 && p->aux);
  if (p->mode == MCQ_EPI_GDN || p->mode == MCQ_EPI_IGDN) MCQ_CHECK_ARG(p->aux);
  if (p->store == MCQ_STORE_SHUFFLE_NCHW) {
    MCQ_CHECK_ARG((p->out_f32 || p->out_u8) && p->cout % 4 == 0 && p->mode == MCQ_EPI_LINEAR && !p->res1 && !p->res2);
    MCQ_CHECK_ARG(!p->out0_hi && !p->out1_hi);
  } else if (p->store == MCQ_STORE_SHUFFLE_NHWC) {
    MCQ_CHECK_ARG(p->cout % 4 == 0 && (p->cout / 4) % 8 == 0);
  } else {
    MCQ_CHECK_ARG(p->cout % 8 == 0);
  }
  MCQ_CHECK_ARG(p->out_f32 || p->out0_hi || p->out1_hi || (p->out_u8 && p->store == MCQ_STORE_SHUFFLE_NCHW));
  if (p->out_u8) MCQ_CHECK_ARG(p->store == MCQ_STORE_SHUFFLE_NCHW);
  std::memset(&a, 0, sizeof(a));
  a.a_hi = (const __half*)p->a_hi; a.a_lo = (const __half*)p->a_lo;
  a.w_hi = (const __half*)p->w_hi; a.w_lo = (const __half*)p->w_lo;
  a.bias = p->bias; a.res1 = p->res1; a.res2 = p->res2; a.aux = p->aux;
  a.out_f32 = p->out_f32;
  a.out_u8 = (unsigned char*)p->out_u8;
  a.o0_hi = (__half*)p->out0_hi; a.o0_lo = (__half*)p->out0_lo;
  a.o1_hi = (__half*)p->out1_hi; a.o1_lo = (__half*)p->out1_lo;
  a.w_scale = p->w_scale; a.res1_scale = p->res1_scale;
  a.dev_scale = p->dev_scale;
  a.n = p->n; a.hin = p->hin; a.win = p->win; a.cin = p->cin;
  a.hout = p->hin / p->stride; a.wout = p->win / p->stride;
  a.cout = p->cout; a.cout_pad = p->cout_pad; a.ksize = p->ksize; a.stride = p->stride;
  a.ktotal = p->ksize * p->ksize * p->cin;
  a.cin_total = p->cin; a.ch_off = 0;
  a.mode = p->mode; a.store = p->store; a.o0_act = p->out0_act; a.o1_act = p->out1_act; a.passes = p->passes;
  const int direct = opt("direct_epi");
  a.direct_epilogue = direct;
  a.gn_ws = nullptr;
  if (p->gn_partials) {
    int32_t rb = 0, unit = 0;
    const int rc = mcq_conv_gn_layout(p, &rb, &unit);
    if (rc) return rc;
    a.gn_ws = (float2*)p->gn_partials;
    a.gn_unit = unit; a.gn_units = p->cout / unit; a.gn_rb = rb;
  }
  const int wait_sleep = opt("wait_sleep_ns");
  a.wait_sleep_ns = wait_sleep;
  return 0;
}

int launch_simt(const ConvArgs& a, cudaStream_t st) {
  const long long M = (long long)a.n * a.hout * a.wout;
  dim3 grid((unsigned)((M + SIMT_TM - 1) / SIMT_TM), (unsigned)((a.cout + SIMT_TN - 1) / SIMT_TN));
  EvScope ev(st);
  conv_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(a);
  g_launches++;
  return cuda_status();
}

// channel counts the 64-channel K chunk does not divide (Neon's 8 / 32-channel nets): stride-1 convs over a whole
// tensor run with a partly zero-filled last chunk (see kchunks in the kernels); fp16 rows must stay 16 B aligned
bool partial_chunk_ok(const ConvArgs& a) {
  return a.cin % 8 == 0 && a.stride == 1 && a.cin_total == a.cin && a.ch_off == 0;
}

bool tc_supported(const ConvArgs& a) {
  if (a.cin % TC_BK != 0 && !partial_chunk_ok(a)) return false;
  if (a.cout_pad % 16 != 0) return false;
  return true;
}

// M tile of the per-tap kernels: a (tw x th x tn) box of output pixels, 128 GEMM rows
void fill_mtile(ConvArgs& a) {
  a.tw = pow2_ceil(a.wout) < 16 ? pow2_ceil(a.wout) : 16;
  const int th_max = TC_BM / a.tw;
  a.th = pow2_ceil(a.hout) < th_max ? pow2_ceil(a.hout) : th_max;
  a.tn = TC_BM / (a.tw * a.th);
  a.tiles_x = (a.wout + a.tw - 1) / a.tw;
  a.tiles_y = (a.hout + a.th - 1) / a.th;
  a.tiles_n = (a.n + a.tn - 1) / a.tn;
}

// per-tap TMA coordinate offsets in the 5-D activation view
void fill_taps(ConvArgs& a) {
  const int pad = a.ksize / 2;
  for (int t = 0; t < a.ksize * a.ksize; ++t) {
    const int r = t / a.ksize - pad, s = t % a.ksize - pad;  // offsets in [-1, 1]
    if (a.stride == 1) {
      a.tap_c[t] = 0; a.tap_dx[t] = s; a.tap_py[t] = 0; a.tap_dy[t] = r;
    } else {
      // input row 2*oy + r: parity (r & 1), coarse row oy + floor(r / 2)
      a.tap_py[t] = r & 1;
      a.tap_dy[t] = (r < 0) ? -1 : 0;
      a.tap_c[t] = (s & 1) * a.cin;
      a.tap_dx[t] = (s < 0) ? -1 : 0;
    }
  }
}

int launch_tc(ConvArgs& a, cudaStream_t st) {
  // ---- N tile: the whole (padded) cout when it fits one MMA, else 128-column tiles
  int bn = (a.cout_pad <= 256 && a.cout_pad % 128 != 0) ? a.cout_pad : 128;
  if (a.cout_pad < 128) bn = a.cout_pad;
  if (a.cout_pad % bn != 0) return MCQ_ERR_UNSUPPORTED;
  if (bn % 16 != 0 || bn > 256) return MCQ_ERR_UNSUPPORTED;
  if (bn > 32 && bn % 32 != 0) return MCQ_ERR_UNSUPPORTED;
  if ((a.passes == 3 ? 2 : 1) * bn > (int)TC_TMEM_COLS) return MCQ_ERR_UNSUPPORTED;
  fill_mtile(a);
  // small feature maps: narrow the N tile so that the few pixel tiles still spread over the SMs
  // (these layers are latency-bound; re-reading A per N tile is free compared with idle SMs)
  const int tiles_m = a.tiles_x * a.tiles_y * a.tiles_n;
  const int spread_pct = opt("tc_spread");
  while (tiles_m * (a.cout_pad / bn) < num_sms() * spread_pct / 100 && bn >= 32 && (bn / 2) % 16 == 0) bn /= 2;
  a.bn = bn;
  a.tiles_c = a.cout_pad / bn;
  fill_taps(a);
  // ---- pipeline depth
  const size_t stage_bytes = (size_t)(TC_A_BYTES + bn * TC_BK * 2) * (a.passes == 3 ? 2 : 1);
  const size_t epi_bytes = (size_t)TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 512 + TC_BIAS_SMEM_FLOATS * 4;
  // 227 KB of dynamic smem per CTA: 1 KB alignment slack, 256 B barriers, epilogue staging, the rest = pipeline
  int stages = (int)((227 * 1024 - 1024 - 256 - epi_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return MCQ_ERR_UNSUPPORTED;
  a.stages = stages;
  const size_t smem = stage_bytes * stages + 8 * (2 * stages + 4) + 16 + 1024 + epi_bytes;
  a.debug_skip_store = opt("epi_skip");

  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  int rc = encode_act_map(&tmA_hi, a.a_hi, a.n, a.hin, a.win, a.cin_total, a.stride, a.tw, a.th, a.tn);
  if (rc) return rc;
  rc = encode_weight_map(&tmB_hi, a.w_hi, a.cout_pad, a.ktotal, bn);
  if (rc) return rc;
  if (a.passes == 3) {
    rc = encode_act_map(&tmA_lo, a.a_lo, a.n, a.hin, a.win, a.cin_total, a.stride, a.tw, a.th, a.tn);
    if (rc) return rc;
    rc = encode_weight_map(&tmB_lo, a.w_lo, a.cout_pad, a.ktotal, bn);
    if (rc) return rc;
  } else {
    tmA_lo = tmA_hi;
    tmB_lo = tmB_hi;
  }
  const int total_tiles = a.tiles_x * a.tiles_y * a.tiles_n * a.tiles_c;
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 1 : 0;
  EvScope ev(st);
  OutMaps om;
  std::memset(&om, 0, sizeof(om));
  const int drain = drain_kind(a, total_tiles / grid);
  if (drain == DRAIN_TMA) {
    rc = encode_out_maps(om, a);
    if (rc) return rc;
  }
#define MCQ_LAUNCH_TC(P, D, TAG)                                                                    \
  do {                                                                                              \
    e = ensure_dyn_smem<KTag<TAG>>(conv_tc_kernel<P, D>, 227 * 1024);                                \
    if (e != cudaSuccess) return (int)e;                                                            \
    e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<P, D>, tmA_hi, tmA_lo, tmB_hi, tmB_lo, a, om);       \
  } while (0)
  if (a.passes == 3) {
    if (drain == DRAIN_TMA) MCQ_LAUNCH_TC(3, DRAIN_TMA, 123);
    else MCQ_LAUNCH_TC(3, DRAIN_ROWS, 103);
  } else {
    if (drain == DRAIN_TMA) MCQ_LAUNCH_TC(1, DRAIN_TMA, 121);
    else MCQ_LAUNCH_TC(1, DRAIN_ROWS, 101);
  }
#undef MCQ_LAUNCH_TC
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

// ---- halo kernel (3x3, stride 1): tap reuse in smem + weight multicast across a cluster
bool halo_supported(const ConvArgs& a) {
  return a.ksize == 3 && a.stride == 1 && (a.cin % TC_BK == 0 || partial_chunk_ok(a)) && a.wout >= HALO_TW &&
         a.hout >= HALO_TH &&
         a.cout_pad % 16 == 0;
}

template <int PASSES, int CL, int DRAIN = DRAIN_ROWS>
int launch_halo_t(ConvArgs& a, HaloArgs& hp, const CUtensorMap* maps, size_t smem, int grid, cudaStream_t st,
                  const OutMaps& om) {
  auto kern = conv_halo_kernel<PASSES, CL, DRAIN>;
  {
    cudaError_t e = ensure_dyn_smem<KTag<200 + PASSES * 10 + CL + DRAIN * 50>>(kern, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 2 : 1;
  EvScope ev(st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], a, hp, om);
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

int launch_halo(ConvArgs& a, cudaStream_t st) {
  int bn = (a.cout_pad <= 256 && a.cout_pad % 128 != 0) ? a.cout_pad : 128;
  if (a.cout_pad < 128) bn = a.cout_pad;
  if (a.cout_pad % bn != 0 || (bn % 32 != 0 && bn != 16)) return MCQ_ERR_UNSUPPORTED;
  const int np = a.passes == 3 ? 2 : 1;
  if (np * bn > (int)TC_TMEM_COLS) return MCQ_ERR_UNSUPPORTED;
  a.bn = bn;
  a.tiles_c = a.cout_pad / bn;
  a.tw = HALO_TW; a.th = HALO_TH; a.tn = 1;
  a.tiles_x = (a.wout + HALO_TW - 1) / HALO_TW;
  a.tiles_y = (a.hout + HALO_TH - 1) / HALO_TH;
  a.tiles_n = a.n;
  HaloArgs hp{};
  hp.pitch = opt("halo_pitch");
  hp.box_w = hp.pitch;
  hp.base_mode = opt("halo_base");
  hp.a_bytes = ((hp.pitch * HALO_ROWS * 128) + 1023) / 1024 * 1024;
  hp.tiles_m = a.tiles_x * a.tiles_y * a.tiles_n;
  int cl = opt("halo_cl");
  if (cl != 1 && cl != 2 && cl != 4) return MCQ_ERR_BAD_ARG;
  while (cl > 1 && ((bn / cl) % 8 != 0 || hp.tiles_m < cl)) cl /= 2;
  hp.groups_m = (hp.tiles_m + cl - 1) / cl;
  // smem budget: A buffers (double/triple) + as many weight stages as fit
  const size_t a_buf = (size_t)hp.a_bytes * np;
  hp.tps = opt("halo_tps") ? opt("halo_tps") : (a.passes == 3 ? 1 : 3);
  if (hp.tps != 1 && hp.tps != 3) return MCQ_ERR_BAD_ARG;
  const size_t b_stage = (size_t)bn * TC_BK * 2 * np * hp.tps;
  const size_t epi_bytes = (size_t)TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 512 + TC_BIAS_SMEM_FLOATS * 4;
  const size_t budget = 227 * 1024 - 1024 - 256 - epi_bytes;
  a.debug_skip_store = opt("epi_skip");
  hp.na = (a.passes == 3) ? 2 : 3;
  int nbs = (int)((budget - a_buf * hp.na) / b_stage);
  if (nbs < 3 && hp.na > 2) { hp.na = 2; nbs = (int)((budget - a_buf * hp.na) / b_stage); }
  if (nbs < 2 && hp.tps > 1) return MCQ_ERR_UNSUPPORTED;
  if (nbs > 8) nbs = 8;
  if (opt("halo_nbs") > 0 && opt("halo_nbs") < nbs) nbs = opt("halo_nbs");
  if (nbs < 2) return MCQ_ERR_UNSUPPORTED;
  {
    // weight-stationary mode: a single N tile whose (cin / 64) * (9 / tps) stages all fit in the ring
    const int all_stages = ((a.cin + TC_BK - 1) / TC_BK) * (9 / hp.tps);
    if (a.tiles_c == 1 && all_stages <= nbs && opt("halo_resident")) {
      hp.resident = 1;
      nbs = all_stages;
    }
  }
  hp.nbs = nbs;
  const size_t smem = a_buf * hp.na + b_stage * nbs + 8 * (2 * hp.na + 2 * nbs + 4) + 16 + 1024 + epi_bytes;

  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  CUtensorMap maps[4];
  auto enc_a = [&](CUtensorMap* m, const void* ptr) {
    cuuint64_t dims[5] = {(cuuint64_t)a.cin, (cuuint64_t)a.win, 1, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    cuuint64_t strides[4] = {(cuuint64_t)a.cin * 2, (cuuint64_t)a.win * a.cin * 2, (cuuint64_t)a.win * a.cin * 2,
                             (cuuint64_t)a.hin * a.win * a.cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)hp.box_w, 1u, (cuuint32_t)HALO_ROWS, 1u};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : MCQ_ERR_DRIVER;
  };
  int rc = enc_a(&maps[0], a.a_hi);
  if (rc) return rc;
  rc = encode_weight_map(&maps[2], a.w_hi, a.cout_pad, a.ktotal, bn / cl);
  if (rc) return rc;
  if (a.passes == 3) {
    rc = enc_a(&maps[1], a.a_lo);
    if (rc) return rc;
    rc = encode_weight_map(&maps[3], a.w_lo, a.cout_pad, a.ktotal, bn / cl);
    if (rc) return rc;
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const int work = hp.groups_m * a.tiles_c;
  int clusters = num_sms() / cl;
  if (work < clusters) {
    const int small_pct = opt("small_grid_pct");
    clusters = work < clusters * small_pct / 100 ? work : clusters * small_pct / 100;
  }
  const int grid = clusters * cl;
  OutMaps om;
  std::memset(&om, 0, sizeof(om));
  // the drain variants are instantiated for the default cluster size only
  const int drain = cl == 2 ? drain_kind(a, work / clusters) : DRAIN_ROWS;
  if (drain == DRAIN_TMA) {
    rc = encode_out_maps(om, a);
    if (rc) return rc;
    if (a.passes == 3) return launch_halo_t<3, 2, DRAIN_TMA>(a, hp, maps, smem, grid, st, om);
    return launch_halo_t<1, 2, DRAIN_TMA>(a, hp, maps, smem, grid, st, om);
  }
  if (a.passes == 3) {
    if (cl == 1) return launch_halo_t<3, 1>(a, hp, maps, smem, grid, st, om);
    if (cl == 2) return launch_halo_t<3, 2>(a, hp, maps, smem, grid, st, om);
    return launch_halo_t<3, 4>(a, hp, maps, smem, grid, st, om);
  }
  if (cl == 1) return launch_halo_t<1, 1>(a, hp, maps, smem, grid, st, om);
  if (cl == 2) return launch_halo_t<1, 2>(a, hp, maps, smem, grid, st, om);
  return launch_halo_t<1, 4>(a, hp, maps, smem, grid, st, om);
}


// ---- CTA-pair kernel (tcgen05.mma.cta_group::2): 3x3 stride-1 convs with a 128-column N tile
template <int PASSES, bool GN = false, int DRAIN = DRAIN_ROWS>
int launch_pair_t(ConvArgs& a, HaloArgs& hp, const CUtensorMap* maps, size_t smem, int grid, cudaStream_t st,
                  const OutMaps& om) {
  auto kern = conv_pair_kernel<PASSES, GN, DRAIN>;
  {
    cudaError_t e = ensure_dyn_smem<KTag<300 + PASSES * 10 + (GN ? 1 : 0) + DRAIN * 2>>(kern, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 2 : 1;
  EvScope ev(st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], a, hp, om);
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

int launch_pair(ConvArgs& a, cudaStream_t st) {
  // N tile: 128 columns, or the whole of a narrow layer (16 / 32 / 64 padded output channels: each CTA of the pair then
  // holds 8 / 16 / 32 weight rows -- whole 8-row swizzle atoms)
  int bn = 128;
  if (a.cout_pad < 128) {
    if (!opt("pair_narrow") || (a.cout_pad != 16 && a.cout_pad != 32 && a.cout_pad != 64)) return MCQ_ERR_UNSUPPORTED;
    bn = a.cout_pad;
  } else if (a.cout_pad % 128 != 0 && a.cout_pad % 96 == 0 && a.passes == 3 && opt("pair_96")) {
    // 192-channel models (qp >= 3: Compressor(192, ...)): two 96-column N tiles -- the 3-pass [hi*hi | hi*lo] MMA is then
    // N = 192 wide, each CTA of the pair holds 96 / 48 weight rows (whole 8-row swizzle atoms)
    bn = 96;
  }
  if (a.cout_pad % bn != 0) return MCQ_ERR_UNSUPPORTED;
  const int np = a.passes == 3 ? 2 : 1;
  a.bn = bn;
  a.tiles_c = a.cout_pad / bn;
  a.tw = HALO_TW; a.th = HALO_TH; a.tn = 1;
  a.tiles_x = (a.wout + HALO_TW - 1) / HALO_TW;
  a.tiles_y = (a.hout + HALO_TH - 1) / HALO_TH;
  a.tiles_n = a.n;
  HaloArgs hp{};
  hp.pitch = 10;
  hp.box_w = 10;
  hp.base_mode = 0;
  hp.a_bytes = ((hp.pitch * HALO_ROWS * 128) + 1023) / 1024 * 1024;
  hp.tiles_m = a.tiles_x * a.tiles_y * a.tiles_n;
  if (hp.tiles_m < 2) return MCQ_ERR_UNSUPPORTED;
  hp.groups_m = (hp.tiles_m + 1) / 2;
  hp.tps = opt("halo_tps") ? opt("halo_tps") : (a.passes == 3 ? 1 : 3);
  if (hp.tps != 1 && hp.tps != 3) return MCQ_ERR_BAD_ARG;
  const size_t a_buf = (size_t)hp.a_bytes * np;
  const size_t b_plane = (size_t)bn * TC_BK * 2;
  const size_t b_stage = (a.passes == 3 ? b_plane + b_plane / 2 : b_plane / 2) * hp.tps;
  const size_t epi_bytes = (size_t)TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 512 + TC_BIAS_SMEM_FLOATS * 4;
  const size_t budget = 227 * 1024 - 1024 - 256 - epi_bytes;
  a.debug_skip_store = opt("epi_skip");
  hp.na = (a.passes == 3) ? 2 : 3;
  int nbs = (int)((budget - a_buf * hp.na) / b_stage);
  if (nbs > 8) nbs = 8;
  if (opt("pair_nbs") > 0 && opt("pair_nbs") < nbs) nbs = opt("pair_nbs");   // A/B knob: fewer weight stages
  // weight-stationary mode: 1-pass, one N tile, and all (cin / 64) * (9 / tps) weight stages fit beside two halo buffers
  const int all_stages = ((a.cin + TC_BK - 1) / TC_BK) * (9 / hp.tps);
  if (a.passes == 1 && a.tiles_c <= num_sms() / 2 && all_stages <= 8 && a_buf * 2 + b_stage * all_stages <= budget &&
      opt("pair_resident")) {
    hp.resident = 1;
    hp.na = 2;
    nbs = all_stages;
  }
  if (nbs < 2) return MCQ_ERR_UNSUPPORTED;
  hp.nbs = nbs;
  const size_t smem = a_buf * hp.na + b_stage * nbs + 8 * (2 * hp.na + 2 * nbs + 4) + 16 + 1024 + epi_bytes;

  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  CUtensorMap maps[5];
  auto enc_a = [&](CUtensorMap* m, const void* ptr) {
    cuuint64_t dims[5] = {(cuuint64_t)a.cin, (cuuint64_t)a.win, 1, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    cuuint64_t strides[4] = {(cuuint64_t)a.cin * 2, (cuuint64_t)a.win * a.cin * 2, (cuuint64_t)a.win * a.cin * 2,
                             (cuuint64_t)a.hin * a.win * a.cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)hp.box_w, 1u, (cuuint32_t)HALO_ROWS, 1u};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : MCQ_ERR_DRIVER;
  };
  int rc = enc_a(&maps[0], a.a_hi);
  if (rc) return rc;
  rc = encode_weight_map(&maps[4], a.w_hi, a.cout_pad, a.ktotal, bn / 2);
  if (rc) return rc;
  if (a.passes == 3) {
    rc = enc_a(&maps[1], a.a_lo);
    if (rc) return rc;
    rc = encode_weight_map(&maps[2], a.w_hi, a.cout_pad, a.ktotal, bn);
    if (rc) return rc;
    rc = encode_weight_map(&maps[3], a.w_lo, a.cout_pad, a.ktotal, bn);
    if (rc) return rc;
  } else {
    maps[1] = maps[0];
    maps[2] = maps[4];
    maps[3] = maps[4];
  }
  const int work = hp.groups_m * a.tiles_c;
  int clusters = num_sms() / 2;
  if (work < clusters) {
    // small layer (<= one work item per cluster): experiment knob -- leave SMs to the sibling branch's launch
    const int small_pct = opt("small_grid_pct");
    clusters = work < clusters * small_pct / 100 ? work : clusters * small_pct / 100;
  }
  if (hp.resident && a.tiles_c > 1) {
    // several N tiles: a cluster keeps ONE of them resident, so the clusters are divided evenly among the N tiles
    hp.per_ct = clusters / a.tiles_c;
    if (hp.per_ct > hp.groups_m) hp.per_ct = hp.groups_m;
    if (hp.per_ct < 1) return MCQ_ERR_UNSUPPORTED;
    clusters = hp.per_ct * a.tiles_c;
  }
  const int grid = clusters * 2;
  OutMaps om;
  std::memset(&om, 0, sizeof(om));
  if (a.gn_ws) {   // GroupNorm partials: the instantiation whose drain also reduces (sum, sum^2) per row block
    if (a.passes == 3) return launch_pair_t<3, true>(a, hp, maps, smem, grid, st, om);
    return launch_pair_t<1, true>(a, hp, maps, smem, grid, st, om);
  }
  const int drain = drain_kind(a, work / clusters);
  if (drain == DRAIN_TMA) {
    rc = encode_out_maps(om, a);
    if (rc) return rc;
    if (a.passes == 3) return launch_pair_t<3, false, DRAIN_TMA>(a, hp, maps, smem, grid, st, om);
    return launch_pair_t<1, false, DRAIN_TMA>(a, hp, maps, smem, grid, st, om);
  }
  if (a.passes == 3) return launch_pair_t<3>(a, hp, maps, smem, grid, st, om);
  return launch_pair_t<1>(a, hp, maps, smem, grid, st, om);
}

// ---- layer chain (conv_chain.cuh): `count` dependent convolutions on small maps in one persistent launch
template <int PASSES>
int launch_chain_t(const ChainParams& cp, size_t smem, int grid, cudaStream_t st) {
  auto kern = conv_chain_kernel<PASSES>;
  {
    cudaError_t e = ensure_dyn_smem<KTag<400 + PASSES>>(kern, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CHAIN_CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 2 : 1;
  EvScope ev(st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, cp);
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

int launch_chain(const mcq_conv_params* params, int count, cudaStream_t st) {
  if (count < 1 || count > CHAIN_MAX_LAYERS) return MCQ_ERR_UNSUPPORTED;
  static thread_local ChainParams cp;   // 27 KB: keep it off the stack
  std::memset(&cp, 0, sizeof(cp));
  const int passes = params[0].passes;
  const int n = params[0].n;
  const int bn_max = passes == 3 ? 64 : 128;
  int ipc = opt("chain_ipc");
  if (ipc <= 0) ipc = (n + 15) / 16;    // <= 16 clusters = 128 CTAs
  const int nclusters = (n + ipc - 1) / ipc;
  int bias_total = 0;
  // buffers written since the last barrier (outputs of the layers after it)
  const void* dirty[CHAIN_MAX_LAYERS * 5];
  int ndirty = 0;
  for (int l = 0; l < count; ++l) {
    const mcq_conv_params* p = &params[l];
    if (p->impl != MCQ_IMPL_TCGEN05 || p->passes != passes || p->n != n) return MCQ_ERR_UNSUPPORTED;
    ChainLayer& L = cp.layers[l];
    ConvArgs& a = L.p;
    int rc = fill_args(p, a);
    if (rc) return rc;
    if (!tc_supported(a) || a.cin % TC_BK != 0) return MCQ_ERR_UNSUPPORTED;
    fill_mtile(a);
    fill_taps(a);
    int bn = a.cout_pad < bn_max ? a.cout_pad : bn_max;
    if (a.cout_pad % bn != 0 || bn % 16 != 0 || (bn > 32 && bn % 32 != 0)) return MCQ_ERR_UNSUPPORTED;
    const int tiles_m = ((ipc + a.tn - 1) / a.tn) * a.tiles_y * a.tiles_x;
    while (tiles_m * (a.cout_pad / bn) < CHAIN_CL && bn >= 32 && (bn / 2) % 16 == 0) bn /= 2;
    a.bn = bn;
    a.tiles_c = a.cout_pad / bn;
    a.debug_skip_store = opt("epi_skip");
    // dependency on anything written since the last barrier?
    const void* ins[5] = {p->a_hi, p->a_lo, p->res1, p->res2, p->aux};
    int dep = 0;
    for (int i = 0; i < 5 && !dep; ++i)
      for (int j = 0; j < ndirty; ++j)
        if (ins[i] && ins[i] == dirty[j]) { dep = 1; break; }
    L.sync_before = (l > 0 && dep) ? 1 : 0;
    if (L.sync_before) ndirty = 0;
    const void* outs[5] = {p->out_f32, p->out0_hi, p->out0_lo, p->out1_hi, p->out1_lo};
    for (int i = 0; i < 5; ++i)
      if (outs[i]) dirty[ndirty++] = outs[i];
    L.bias_off = bias_total;
    bias_total += (a.cout + 3) / 4 * 4;
    rc = encode_act_map(&L.tmA_hi, a.a_hi, a.n, a.hin, a.win, a.cin_total, a.stride, a.tw, a.th, a.tn);
    if (rc) return rc;
    rc = encode_weight_map(&L.tmB_hi, a.w_hi, a.cout_pad, a.ktotal, bn);
    if (rc) return rc;
    if (passes == 3) {
      rc = encode_act_map(&L.tmA_lo, a.a_lo, a.n, a.hin, a.win, a.cin_total, a.stride, a.tw, a.th, a.tn);
      if (rc) return rc;
      rc = encode_weight_map(&L.tmB_lo, a.w_lo, a.cout_pad, a.ktotal, bn);
      if (rc) return rc;
    } else {
      L.tmA_lo = L.tmA_hi;
      L.tmB_lo = L.tmB_hi;
    }
  }
  if (bias_total > CHAIN_BIAS_FLOATS) return MCQ_ERR_UNSUPPORTED;
  const size_t stage_bytes = (size_t)(TC_A_BYTES + bn_max * TC_BK * 2) * (passes == 3 ? 2 : 1);
  const size_t epi_bytes = (size_t)TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 128 + (size_t)bias_total * 4;
  int stages = (int)((227 * 1024 - 1024 - 256 - epi_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return MCQ_ERR_UNSUPPORTED;
  cp.count = count;
  cp.ipc = ipc;
  cp.stages = stages;
  cp.bias_total = bias_total;
  cp.debug = (opt("chain_nosync") ? 1 : 0) | (opt("chain_sync_mode") << 1);
  cp.dbg = g_chain_dbg;
  const size_t smem = stage_bytes * stages + 8 * (2 * stages + 4) + 16 + 1024 + epi_bytes;
  const int grid = nclusters * CHAIN_CL;
  return passes == 3 ? launch_chain_t<3>(cp, smem, grid, st) : launch_chain_t<1>(cp, smem, grid, st);
}

}  // namespace

extern "C" {

int mcq_conv2d(const mcq_conv_params* p, mcq_stream_t stream) {
  ConvArgs a;
  int rc = fill_args(p, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  g_ev_start = (cudaEvent_t)p->ev_start;
  g_ev_stop = (cudaEvent_t)p->ev_stop;
  if (p->impl == MCQ_IMPL_SIMT) rc = launch_simt(a, st);
  else if (p->impl != MCQ_IMPL_TCGEN05) rc = MCQ_ERR_BAD_ARG;
  else if (!tc_supported(a)) rc = MCQ_ERR_UNSUPPORTED;
  else {
    rc = MCQ_ERR_UNSUPPORTED;
    // CTA pairs (cta_group::2) pay off where the tensor pipe is the limiter (3-pass); the 1-pass path is epilogue-bound
    // (streaming weights); with a single 128-column N tile the weights stay resident in the pair's shared memory instead
    const bool pair_resident = a.passes == 1 && ((a.cout_pad % 128 == 0 && a.cout_pad <= 512) || a.cout_pad < 128) &&
                               a.cin == 128 && opt("pair_resident");
    if (a.gn_ws) rc = halo_supported(a) ? launch_pair(a, st) : MCQ_ERR_UNSUPPORTED;   // only the pair kernel has it
    else
    // (streaming-weight 1-pass pairs pay off once K is long enough to amortise the epilogue: cin >= 256, measured on the
    //  a800_16 training step: 696 -> 666 ms)
    if (halo_supported(a) && (opt("pair") >= 0 ? opt("pair") : ((a.passes == 3 || pair_resident || a.cin >= 256) ? 1 : 0)))
      rc = launch_pair(a, st);
    if (rc == MCQ_ERR_UNSUPPORTED && !a.gn_ws && halo_supported(a) && opt("halo")) rc = launch_halo(a, st);
    if (rc == MCQ_ERR_UNSUPPORTED && !a.gn_ws) rc = launch_tc(a, st);
  }
  g_ev_start = g_ev_stop = nullptr;
  return rc;
}

int mcq_conv_chain(const mcq_conv_params* params, int32_t count, mcq_stream_t stream) {
  MCQ_CHECK_ARG(params && count >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  g_ev_start = (cudaEvent_t)params[0].ev_start;
  g_ev_stop = (cudaEvent_t)params[count - 1].ev_stop;
  int rc = opt("chain") ? launch_chain(params, count, st) : MCQ_ERR_UNSUPPORTED;
  g_ev_start = g_ev_stop = nullptr;
  return rc;
}

int32_t mcq_conv_chain_max_layers(void) { return CHAIN_MAX_LAYERS; }

void mcq_debug_timeline(void* device_i64_3072) { g_chain_dbg = (long long*)device_i64_3072; }

int mcq_stem_conv(const void* x, int32_t x_is_u8, int32_t n, int32_t h, int32_t w, int32_t pad_top, int32_t pad_left,
                  int32_t hp, int32_t wp, const float* wgt, const float* bias, int32_t cout, float* out_f32,
                  void* out_hi, void* out_lo, int32_t out_act, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && wgt && bias && (out_f32 || out_hi));
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && hp >= h && wp >= w && hp % 2 == 0 && wp % 2 == 0);
  MCQ_CHECK_ARG(pad_top >= 0 && pad_left >= 0 && pad_top <= hp - h && pad_left <= wp - w);
  MCQ_CHECK_ARG(pad_top < h && (hp - h - pad_top) < h && pad_left < w && (wp - w - pad_left) < w);  // reflect pad limit
  MCQ_CHECK_ARG(cout % 4 == 0 && (27 * cout + 9 * (2 * STEM_PIX + 1)) * 4 <= 48 * 1024);
  StemArgs a;
  a.x = x_is_u8 ? nullptr : (const float*)x;
  a.x_u8 = x_is_u8 ? (const unsigned char*)x : nullptr;
  a.w = wgt; a.bias = bias; a.out_f32 = out_f32; a.o_hi = (__half*)out_hi; a.o_lo = (__half*)out_lo;
  a.o_act = out_act; a.n = n; a.h = h; a.w_ = w; a.pad_top = pad_top; a.pad_left = pad_left; a.hp = hp; a.wp = wp;
  a.cout = cout; a.hout = hp / 2; a.wout = wp / 2;
  const int strips = (a.wout + STEM_PIX - 1) / STEM_PIX;
  const long long blocks = (long long)n * a.hout * strips;
  const size_t smem = (27 * (size_t)cout + 9 * (2 * STEM_PIX + 1)) * sizeof(float);
  stem_conv_kernel<<<(unsigned)blocks, STEM_THREADS, smem, (cudaStream_t)stream>>>(a);
  g_launches++;
  return cuda_status();
}

int mcq_stem_conv_tc(const void* x, int32_t x_is_u8, int32_t n, int32_t h, int32_t w, int32_t pad_top,
                     int32_t pad_left, int32_t hp, int32_t wp, const void* w_lohi, float w_scale, const float* bias,
                     int32_t cout, int32_t cout_pad, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                     mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && w_lohi && bias && (out_f32 || out_hi));
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && hp >= h && wp >= w && hp % 2 == 0 && wp % 2 == 0);
  MCQ_CHECK_ARG(pad_top >= 0 && pad_left >= 0 && pad_top <= hp - h && pad_left <= wp - w);
  MCQ_CHECK_ARG(pad_top < h && (hp - h - pad_top) < h && pad_left < w && (wp - w - pad_left) < w);  // reflect pad limit
  MCQ_CHECK_ARG(cout > 0 && cout_pad >= cout && ((uintptr_t)w_lohi & 15) == 0);
  if (cout % 8 != 0 || cout_pad % 16 != 0 || cout_pad > 128) return MCQ_ERR_UNSUPPORTED;
  if ((long long)n * (hp / 2) * (wp / 2) > 0x7fffffffLL * 64) return MCQ_ERR_UNSUPPORTED;
  ConvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.bias = bias; a.out_f32 = out_f32; a.o0_hi = (__half*)out_hi; a.o0_lo = (__half*)out_lo; a.o0_act = out_act;
  a.w_scale = w_scale; a.res1_scale = 1.f;
  a.n = n; a.hin = hp; a.win = wp; a.cin = 3; a.hout = hp / 2; a.wout = wp / 2;
  a.cout = cout; a.cout_pad = cout_pad; a.ksize = 3; a.stride = 2; a.ktotal = 27;
  a.mode = MCQ_EPI_LINEAR; a.store = MCQ_STORE_NHWC; a.passes = 3; a.bn = cout_pad; a.tiles_c = 1;
  a.direct_epilogue = opt("direct_epi");
  a.debug_skip_store = opt("epi_skip");
  StemTcArgs s;
  s.x_f32 = x_is_u8 ? nullptr : (const float*)x;
  s.x_u8 = x_is_u8 ? (const unsigned char*)x : nullptr;
  s.w_lohi = (const __half*)w_lohi;
  s.h = h; s.w = w; s.pad_top = pad_top; s.pad_left = pad_left; s.hp = hp; s.wp = wp; s.cout_pad = cout_pad;
  const long long total_pix = (long long)n * a.hout * a.wout;
  const long long tiles = (total_pix + STC_BM - 1) / STC_BM;
  const size_t smem = 1024 + 2 * STC_A_BYTES + 16384 + 512 + (size_t)TC_EPI_WARPS * TC_EPI_STAGE_BYTES + 128 * 4 + 256 * 4 + 64;
  const long long per_cta = tiles / (tiles < num_sms() ? tiles : num_sms());
  int drain = drain_kind(a, per_cta);
  if (drain == DRAIN_TMA && total_pix >= 0x7fffffffLL) drain = DRAIN_ROWS;
  OutMaps om;
  std::memset(&om, 0, sizeof(om));
  if (drain == DRAIN_TMA) {
    // the stem's tile is 128 consecutive pixels of the flattened [n, y, x] grid: view the outputs as [pixels, channels]
    ConvArgs v = a;
    v.wout = (int)total_pix; v.hout = 1; v.n = 1; v.tw = 32; v.th = 1; v.tn = 4;
    const int rc = encode_out_maps(om, v);
    if (rc) return rc;
  }
  cudaError_t e = drain == DRAIN_TMA ? ensure_dyn_smem<KTag<802>>(stem_tc_kernel<DRAIN_TMA>, smem)
                                     : ensure_dyn_smem<KTag<800>>(stem_tc_kernel<DRAIN_ROWS>, smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(tiles < num_sms() ? tiles : num_sms()));
  cfg.blockDim = dim3(STC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 1 : 0;
  e = drain == DRAIN_TMA ? cudaLaunchKernelEx(&cfg, stem_tc_kernel<DRAIN_TMA>, a, s, om)
                         : cudaLaunchKernelEx(&cfg, stem_tc_kernel<DRAIN_ROWS>, a, s, om);
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

int mcq_vq_assign(const float* x, const float* codebook, const float* c2, int64_t* codes, float* logits,
                  const float* logit_scale, int32_t* hist, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k,
                  int32_t d, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && codebook && c2 && codes);
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && m > 0 && k > 0 && d > 0 && d % 4 == 0 && d <= 256);
  VqArgs a;
  a.x = x; a.codebook = codebook; a.c2 = c2; a.codes = (long long*)codes; a.logits = logits;
  a.logit_scale = logit_scale; a.hist = hist;
  a.P = n * h * w; a.hw = h * w; a.m = m; a.k = k; a.d = d;
  a.inv_sqrt_k = 1.0f / sqrtf((float)k);
  const size_t smem = ((size_t)d * (VQ_TP + VQ_TK) + VQ_TP) * sizeof(float);
  {
    cudaError_t e = ensure_dyn_smem<KTag<500>>(vq_assign_kernel, smem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid((unsigned)((a.P + VQ_TP - 1) / VQ_TP), (unsigned)m);
  vq_assign_kernel<<<grid, VQ_THREADS, smem, (cudaStream_t)stream>>>(a);
  g_launches++;
  return cuda_status();
}

int mcq_vq_fused_supported(int32_t h, int32_t w, int32_t k, int32_t d) {
  const int hw = h * w;
  return (d == 32 || d == 64 || (d == 128 && opt("vq128_fused"))) && k > 0 && k % VQF_BN == 0 && hw > 0 &&
         (hw % 32 == 0 || 32 % hw == 0);
}

int mcq_vq_assign_fused(const float* x, const void* cb_lohi, float cb_scale, const float* c2, int64_t* codes,
                        float* logits, const float* logit_scale, int32_t* hist, int32_t n, int32_t h, int32_t w,
                        int32_t m, int32_t k, int32_t d, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && cb_lohi && c2 && codes);
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && m > 0 && k > 0 && d > 0);
  if (!mcq_vq_fused_supported(h, w, k, d)) return MCQ_ERR_UNSUPPORTED;
  MCQ_CHECK_ARG(((uintptr_t)cb_lohi & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)c2 & 15) == 0);
  if (logits) MCQ_CHECK_ARG(((uintptr_t)logits & 15) == 0);
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MCQ_ERR_DRIVER;
  const int hw = h * w;
  VqFusedArgs a;
  std::memset(&a, 0, sizeof(a));
  a.x = x; a.c2 = c2; a.codes = (long long*)codes; a.hist = hist; a.hist_on = hist ? 1 : 0;
  a.logit_scale = logit_scale;
  a.P = n * hw; a.hw = hw; a.m = m; a.k = k; a.d = d;
  a.tiles_p = (a.P + VQF_BM - 1) / VQF_BM;
  a.has_logits = logits ? 1 : 0;
  a.cb_scale = cb_scale;
  a.inv_sqrt_k = 1.0f / sqrtf((float)k);
  const int kch = 2 * d / TC_BK;
  const size_t op_bytes = (size_t)kch * VQF_CHUNK_BYTES;
  // d = 128: 64 KB operand rows -> one latent buffer, codebook streamed in half rows (32 KB stages), see vq_fused.cuh
  const int na = d == 128 ? 1 : 2, halves = d == 128 ? 2 : 1;
  const size_t bst_bytes = op_bytes / halves;
  const size_t smem_max = 227 * 1024;
  int nst = 2, nb = 0;
  size_t fixed = 0;
  for (; nst >= 1; --nst) {   // prefer double-buffered logits staging; d >= 64 only has room for one buffer per warp
    const size_t st_bytes = logits ? (size_t)VQF_EPI_WARPS * nst * VQF_STAGE_BYTES : 0;
    fixed = 1024 + na * op_bytes + st_bytes + 2 * VQF_BM * 4 + 6 * VQF_BM * 8 + VQF_C2_SLOTS * VQF_BN * 4 + 64;
    nb = fixed + 8 * 16 < smem_max ? (int)((smem_max - fixed - 8 * 16) / bst_bytes) : 0;
    if (nb >= 2) break;
  }
  if (nb < 2) return MCQ_ERR_UNSUPPORTED;
  if (nb > 4) nb = 4;
  a.nb = nb;
  a.nst = nst;
  a.na = na;
  a.halves = halves;
  const size_t smem = fixed + nb * bst_bytes + 8 * (8 + 2 * nb);

  CUtensorMap tmB, tmL;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(2 * d), (cuuint64_t)m * k};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * d) * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)VQF_BN};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(cb_lohi), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MCQ_ERR_DRIVER;
  }
  if (logits) {
    // logits [n, m, hw, k] fp32; a store box = 32 codewords x 32 consecutive latent points of one codebook
    const int blk_pix = hw >= 32 ? 32 : hw, blk_img = hw >= 32 ? 1 : 32 / hw;
    cuuint64_t dims[4] = {(cuuint64_t)k, (cuuint64_t)hw, (cuuint64_t)m, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)k * 4, (cuuint64_t)hw * k * 4, (cuuint64_t)m * hw * k * 4};
    cuuint32_t box[4] = {32u, (cuuint32_t)blk_pix, 1u, (cuuint32_t)blk_img};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, logits, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
        CUDA_SUCCESS)
      return MCQ_ERR_DRIVER;
  } else {
    tmL = tmB;
  }
  auto kern = logits ? vq_fused_kernel<true> : vq_fused_kernel<false>;
  {
    cudaError_t e = logits ? ensure_dyn_smem<KTag<601>>(kern, smem) : ensure_dyn_smem<KTag<600>>(kern, smem);
    if (e != cudaSuccess) return (int)e;
  }
  int grid = a.tiles_p * m;
  if (grid > num_sms()) grid = num_sms();
  kern<<<grid, VQF_THREADS, smem, (cudaStream_t)stream>>>(tmB, tmL, a);
  g_launches++;
  return cuda_status();
}

int64_t mcq_vq_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t m, int32_t k, int32_t d) {
  (void)k;
  const int64_t P = (int64_t)n * h * w, C = (int64_t)m * d;
  // [hi plane | lo plane] fp16, |x|^2 fp32 [P, m], keys u64 [P, m]; every section 256 B aligned
  auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
  return al(P * C * 2) * 2 + al(P * m * 4) + al(P * m * 8);
}

int mcq_vq_assign_tc(const float* x, const void* cb_hi, const void* cb_lo, float cb_scale, const float* c2,
                     int64_t* codes, int32_t* hist, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k, int32_t d,
                     void* workspace, int64_t workspace_bytes, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && cb_hi && cb_lo && c2 && codes && workspace);
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && m > 0 && k > 0 && d > 0);
  if (d % TC_BK != 0 || k % 32 != 0) return MCQ_ERR_UNSUPPORTED;
  MCQ_CHECK_ARG(workspace_bytes >= mcq_vq_workspace_bytes(n, h, w, m, k, d));
  MCQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)n * h * w, C = (int64_t)m * d;
  auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
  char* ws = (char*)workspace;
  __half* hi = (__half*)ws;
  __half* lo = (__half*)(ws + al(P * C * 2));
  float* x2 = (float*)(ws + 2 * al(P * C * 2));
  unsigned long long* keys = (unsigned long long*)(ws + 2 * al(P * C * 2) + al(P * m * 4));
  vq_prep_kernel<<<(unsigned)((P * m + 127) / 128), 128, 0, st>>>(x, (int)P, m, d, hi, lo, x2, keys);
  g_launches++;
  int rc = cuda_status();
  if (rc) return rc;
  for (int mi = 0; mi < m; ++mi) {
    ConvArgs a;
    std::memset(&a, 0, sizeof(a));
    a.a_hi = hi; a.a_lo = lo;
    a.w_hi = (const __half*)cb_hi + (size_t)mi * k * d;   // packed codebook [m, k, d] fp16, value = c * 2^e
    a.w_lo = (const __half*)cb_lo + (size_t)mi * k * d;
    a.bias = c2 + (size_t)mi * k;                          // |c_k|^2
    a.aux = x2 + mi;                                       // |x|^2, stride m
    a.argmin_keys = keys + mi;
    a.argmin_stride = m;
    a.w_scale = cb_scale; a.res1_scale = 1.f;
    a.n = n; a.hin = h; a.win = w; a.cin = d; a.cin_total = (int)C; a.ch_off = mi * d;
    a.hout = h; a.wout = w;
    a.cout = k; a.cout_pad = k; a.ksize = 1; a.stride = 1; a.ktotal = d;
    a.mode = EPI_ARGMIN; a.store = MCQ_STORE_NHWC; a.passes = 3;
    rc = launch_tc(a, st);
    if (rc) return rc;
  }
  vq_finalize_kernel<<<(unsigned)((P * m + 127) / 128), 128, 0, st>>>(keys, (int)P, h * w, m, k, (long long*)codes,
                                                                       hist);
  g_launches++;
  return cuda_status();
}

int mcq_vq_dequant(const int64_t* codes, const float* codebook, int32_t n, int32_t h, int32_t w, int32_t m, int32_t k,
                   int32_t d, float* out_f32, void* out0_hi, void* out0_lo, int32_t out0_act, void* out1_hi,
                   void* out1_lo, int32_t out1_act, int32_t* status, mcq_stream_t stream) {
  MCQ_CHECK_ARG(codes && codebook && (out_f32 || out0_hi || out1_hi));
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && m > 0 && k > 0 && d > 0 && d % 4 == 0);
  DequantArgs a;
  a.codes = (const long long*)codes; a.codebook = codebook; a.out_f32 = out_f32;
  a.o0_hi = (__half*)out0_hi; a.o0_lo = (__half*)out0_lo; a.o1_hi = (__half*)out1_hi; a.o1_lo = (__half*)out1_lo;
  a.o0_act = out0_act; a.o1_act = out1_act; a.status = status;
  a.P = n * h * w; a.hw = h * w; a.m = m; a.k = k; a.d = d;
  const long long total = (long long)a.P * (m * d / 4);
  vq_dequant_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  g_launches++;
  return cuda_status();
}

int mcq_code_histogram(const int64_t* codes, int32_t n, int32_t m, int32_t hw, int32_t k, int32_t* hist,
                       mcq_stream_t stream) {
  MCQ_CHECK_ARG(codes && hist && n > 0 && m > 0 && hw > 0 && k > 0);
  const long long total = (long long)n * m * hw;
  code_histogram_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const long long*)codes, n,
                                                                                          m, hw, k, hist);
  g_launches++;
  return cuda_status();
}

int mcq_groupnorm(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t groups, const float* gamma,
                  const float* beta, float eps, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                  mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && gamma && beta && (out_f32 || out_hi));
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && groups > 0 && c % groups == 0 && eps >= 0.f);
  MCQ_CHECK_ARG((long long)h * w <= 0x7fffffffLL / c);
  if (c % 4 != 0 || c > GN_MAX_C) return MCQ_ERR_UNSUPPORTED;
  GroupNormArgs a;
  a.x = x; a.gamma = gamma; a.beta = beta; a.out_f32 = out_f32; a.o_hi = (__half*)out_hi; a.o_lo = (__half*)out_lo;
  a.o_act = out_act; a.n = n; a.hw = h * w; a.c = c; a.groups = groups; a.eps = eps;
  // a cluster per image in flight; as many CTAs per cluster as leave every CTA >= 32 pixels (tiny maps: a single CTA)
  int slices = GN_MAX_CLUSTER;
  while (slices > 1 && a.hw < slices * 32) slices >>= 1;
  const int threads = opt("gn_threads");
  const int ctas_per_sm = opt("gn_ctas_per_sm");
  // persistent clusters, ctas_per_sm CTAs per SM's worth of them (never more than images)
  long long clusters = ((long long)ctas_per_sm * num_sms()) / slices;
  if (clusters < 1) clusters = 1;
  if (clusters > n) clusters = n;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * slices));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)gn_smem_bytes(threads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)slices;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (threads == 1024) {
    const cudaError_t attr = ensure_dyn_smem<KTag<700>>(groupnorm_kernel<1024>, (size_t)gn_smem_bytes(1024));
    e = attr != cudaSuccess ? attr : cudaLaunchKernelEx(&cfg, groupnorm_kernel<1024>, a);
  } else if (threads == 512) {
    e = cudaLaunchKernelEx(&cfg, groupnorm_kernel<512>, a);
  } else if (threads == 256) {
    e = cudaLaunchKernelEx(&cfg, groupnorm_kernel<256>, a);
  } else {
    return MCQ_ERR_BAD_ARG;
  }
  g_launches++;
  return e == cudaSuccess ? cuda_status() : (int)e;
}

int mcq_conv_gn_layout(const mcq_conv_params* p, int32_t* rowblocks_per_image, int32_t* unit) {
  MCQ_CHECK_ARG(p && rowblocks_per_image && unit);
  MCQ_CHECK_ARG(p->gn_groups > 0 && p->cout > 0 && p->cout % p->gn_groups == 0);
  const int cg = p->cout / p->gn_groups;
  const int hout = p->hin / (p->stride > 0 ? p->stride : 1), wout = p->win / (p->stride > 0 ? p->stride : 1);
  if (p->impl != MCQ_IMPL_TCGEN05 || p->ksize != 3 || p->stride != 1 || p->store != MCQ_STORE_NHWC ||
      p->mode != MCQ_EPI_LINEAR || p->cin % TC_BK != 0 || p->cout % 128 != 0 || p->cout_pad != p->cout ||
      wout < HALO_TW || hout < HALO_TH || !p->out_f32 || !opt("direct_epi") || !opt("gn_fused"))
    return MCQ_ERR_UNSUPPORTED;
  if ((long long)p->n * ((wout + HALO_TW - 1) / HALO_TW) * ((hout + HALO_TH - 1) / HALO_TH) < 2) return MCQ_ERR_UNSUPPORTED;
  if (cg % 4 != 0 || (cg > 16 && cg % 16 != 0) || (cg < 16 && 16 % cg != 0)) return MCQ_ERR_UNSUPPORTED;
  *unit = cg < 16 ? cg : 16;
  // a drain warp owns 32 consecutive tile rows = (32 / HALO_TW) image rows of one HALO_TW-wide tile column
  *rowblocks_per_image = ((hout + 32 / HALO_TW - 1) / (32 / HALO_TW)) * ((wout + HALO_TW - 1) / HALO_TW);
  return 0;
}

int mcq_groupnorm_apply(const float* x, const void* partials, int32_t rowblocks_per_image, int32_t unit, int32_t n,
                        int32_t h, int32_t w, int32_t c, int32_t groups, const float* gamma, const float* beta,
                        float eps, void* stats, float* out_f32, void* out_hi, void* out_lo, int32_t out_act,
                        mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && partials && stats && gamma && beta && (out_f32 || out_hi));
  MCQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && groups > 0 && c % groups == 0 && eps >= 0.f && n <= 65535);
  MCQ_CHECK_ARG(rowblocks_per_image > 0 && unit >= 4 && unit % 4 == 0 && (c / groups) % unit == 0);
  GroupNormApplyArgs a;
  a.x = x; a.partials = (const float2*)partials; a.stats = (float2*)stats; a.gamma = gamma; a.beta = beta;
  a.out_f32 = out_f32; a.o_hi = (__half*)out_hi; a.o_lo = (__half*)out_lo; a.o_act = out_act;
  a.n = n; a.hw = h * w; a.c = c; a.groups = groups; a.rb = rowblocks_per_image; a.unit = unit; a.eps = eps;
  if (c % 4 != 0 || c > 4 * GNA_THREADS) return MCQ_ERR_UNSUPPORTED;
  gn_finalize_kernel<<<dim3((unsigned)groups, (unsigned)n), GNF_THREADS, 0, (cudaStream_t)stream>>>(a);
  g_launches++;
  const int rows = GNA_THREADS / (c / 4);
  const int iters = opt("gn_apply_iters");
  const int ppb = rows * iters;
  gn_apply_kernel<<<dim3((unsigned)((a.hw + ppb - 1) / ppb), (unsigned)n), GNA_THREADS, 0, (cudaStream_t)stream>>>(a, ppb);
  g_launches++;
  return cuda_status();
}

int mcq_add_scaled(const float* x, const float* y, float alpha, int64_t count, float* out_f32, void* out_hi,
                   void* out_lo, int32_t out_act, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && y && (out_f32 || out_hi) && count > 0 && count % 4 == 0);
  const long long c4 = count / 4;
  add_scaled_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, alpha, c4, out_f32, out_act,
                                                                                    (__half*)out_hi, (__half*)out_lo);
  g_launches++;
  return cuda_status();
}

int mcq_split_planes(const float* x, int64_t count, int32_t act, void* out_hi, void* out_lo, const float* dev_scale,
                     mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && out_hi && count > 0 && count % 4 == 0);
  const long long c4 = count / 4;
  split_planes_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, c4, act, (__half*)out_hi,
                                                                                      (__half*)out_lo, dev_scale);
  g_launches++;
  return cuda_status();
}

int mcq_nchw_to_nhwc(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, float* out_f32, void* out0_hi,
                     void* out0_lo, int32_t out0_act, void* out1_hi, void* out1_lo, int32_t out1_act,
                     mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && (out_f32 || out0_hi || out1_hi) && n > 0 && c > 0 && h > 0 && w > 0 && n <= 65535);
  LayoutArgs a;
  a.x = x; a.out_f32 = out_f32; a.o0_hi = (__half*)out0_hi; a.o0_lo = (__half*)out0_lo; a.o1_hi = (__half*)out1_hi;
  a.o1_lo = (__half*)out1_lo; a.o0_act = out0_act; a.o1_act = out1_act; a.n = n; a.c = c; a.hw = h * w;
  dim3 grid((unsigned)((a.hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(a);
  g_launches++;
  return cuda_status();
}

int mcq_nhwc_to_nchw(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, float* out, mcq_stream_t stream) {
  MCQ_CHECK_ARG(x && out && n > 0 && c > 0 && h > 0 && w > 0 && n <= 65535);
  dim3 grid((unsigned)((h * w + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, c, h * w, out);
  g_launches++;
  return cuda_status();
}

// ---- weight gradient (conv_wgrad.cuh)
namespace {
int plan_wgrad(const mcq_wgrad_params* p, WgradArgs& w) {
  MCQ_CHECK_ARG(p && p->n > 0 && p->hin > 0 && p->win > 0 && p->cin > 0 && p->cout > 0);
  MCQ_CHECK_ARG(p->ksize == 1 || p->ksize == 3);
  MCQ_CHECK_ARG(p->stride == 1 || p->stride == 2);
  if (p->cin % 8 != 0 || p->cout % 8 != 0) return MCQ_ERR_UNSUPPORTED;
  if (p->stride == 2 && (p->cin % TC_BK != 0 || p->hin % 2 != 0 || p->win % 2 != 0)) return MCQ_ERR_UNSUPPORTED;
  std::memset(&w, 0, sizeof(w));
  ConvArgs g;
  std::memset(&g, 0, sizeof(g));
  g.n = p->n; g.hin = p->hin; g.win = p->win; g.cin = p->cin;
  g.hout = p->hin / p->stride; g.wout = p->win / p->stride;
  g.ksize = p->ksize; g.stride = p->stride;
  fill_mtile(g);
  fill_taps(g);
  w.n = p->n; w.hout = g.hout; w.wout = g.wout; w.cin = p->cin; w.cout = p->cout;
  w.ksize = p->ksize; w.ntaps = p->ksize * p->ksize;
  w.tw = g.tw; w.th = g.th; w.tn = g.tn; w.tiles_x = g.tiles_x; w.tiles_y = g.tiles_y; w.tiles_n = g.tiles_n;
  w.tiles_ci = (p->cin + 127) / 128; w.tiles_co = (p->cout + 127) / 128;
  w.taps_per_group = w.ntaps < WG_MAX_TAPS ? w.ntaps : WG_MAX_TAPS;
  w.tap_groups = (w.ntaps + w.taps_per_group - 1) / w.taps_per_group;
  const int items = w.tiles_ci * w.tiles_co * w.tap_groups;
  const int tiles_pix = w.tiles_x * w.tiles_y * w.tiles_n;
  int splits = num_sms() / items;
  if (splits < 1) splits = 1;
  if (splits > tiles_pix) splits = tiles_pix;
  w.splits = splits;
  for (int t = 0; t < 9; ++t) { w.tap_c[t] = g.tap_c[t]; w.tap_dx[t] = g.tap_dx[t]; w.tap_py[t] = g.tap_py[t]; w.tap_dy[t] = g.tap_dy[t]; }
  return 0;
}
}  // namespace

int64_t mcq_conv_wgrad_workspace_bytes(const mcq_wgrad_params* p) {
  WgradArgs w;
  if (plan_wgrad(p, w)) return -1;
  return (int64_t)w.tiles_ci * w.tiles_co * w.tap_groups * w.splits * WG_MAX_TAPS * 128 * 128 * 4;
}

int mcq_conv_wgrad(const mcq_wgrad_params* p, mcq_stream_t stream) {
  WgradArgs w;
  int rc = plan_wgrad(p, w);
  if (rc) return rc;
  MCQ_CHECK_ARG(p->x_hi && p->dy_hi && p->dw && p->workspace && ((uintptr_t)p->workspace & 255) == 0);
  const int64_t need = mcq_conv_wgrad_workspace_bytes(p);
  MCQ_CHECK_ARG(p->workspace_bytes >= need);
  w.partial = (float*)p->workspace;
  CUtensorMap tmDY, tmX;
  rc = encode_act_map(&tmDY, p->dy_hi, p->n, w.hout, w.wout, p->cout, 1, w.tw, w.th, w.tn);
  if (rc) return rc;
  rc = encode_act_map(&tmX, p->x_hi, p->n, p->hin, p->win, p->cin, p->stride, w.tw, w.th, w.tn);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)WG_NA * WG_A_BYTES + (size_t)WG_NB * WG_B_BYTES + 8 * (2 * WG_NA + 2 * WG_NB + 2) + 64;
  cudaError_t e = ensure_dyn_smem<KTag<900>>(conv_wgrad_kernel, smem);
  if (e != cudaSuccess) return (int)e;
  cudaStream_t st = (cudaStream_t)stream;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(w.tiles_ci * w.tiles_co * w.tap_groups * w.splits));
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = opt("pdl") ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, conv_wgrad_kernel, tmDY, tmX, w);
  g_launches++;
  if (e != cudaSuccess) return (int)e;
  const long long total = (long long)p->cout * p->cin * w.ntaps;
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.partial, p->dw, w, p->scale, p->dev_scale,
                                                                        p->accumulate);
  g_launches++;
  return cuda_status();
}

const char* mcq_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case MCQ_ERR_BAD_ARG: return "mcquic_b200: bad argument (shape/pointer/alignment contract violated)";
    case MCQ_ERR_UNSUPPORTED: return "mcquic_b200: configuration not supported by the tcgen05 kernel";
    case MCQ_ERR_DRIVER: return "mcquic_b200: CUDA driver call failed (cuTensorMapEncodeTiled)";
    case MCQ_ERR_CODE_RANGE: return "mcquic_b200: code index outside [0, k)";
    case MCQ_ERR_WATCHDOG: return "mcquic_b200: device-side pipeline watchdog fired";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "mcquic_b200: unknown error";
  }
}

int mcq_version(void) { return 2; }

int mcq_set_option(const char* name, int32_t value) {
  int* v = name ? find_option(name) : nullptr;
  if (!v) return MCQ_ERR_BAD_ARG;
  *v = value;
  return 0;
}

int32_t mcq_get_option(const char* name) {
  const int* v = name ? find_option(name) : nullptr;
  return v ? *v : INT32_MIN;
}

int mcq_device_error_flag(void) {
  int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_watchdog_flag, sizeof(int)) != cudaSuccess) return MCQ_ERR_WATCHDOG;
  if (v) {
    int z = 0;
    cudaMemcpyToSymbol(g_watchdog_flag, &z, sizeof(int));
  }
  return v;
}

int mcq_kernel_launch_count(void) { return g_launches.load(); }

}  // extern "C"
