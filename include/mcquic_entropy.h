/*
 * mcquic_entropy -- host-side (CPU, multi-threaded) rANS entropy coder for McQuic code maps, C ABI.
 * SURVEY.md section 8(f) NEXT-1: the rANS coder stays on the host (north_star); this library replaces the
 * reference's pybind11 module `mcquic.rans` (third_party/CompressAI/cpp_exts/*.cpp over ryg_rans/rans64.h) with a
 * batched interface and produces BIT-IDENTICAL byte streams.
 *
 * Stream format (= the reference's): 64-bit rANS state, 32-bit renormalisation words, probability precision 16 bits,
 * symbols encoded in reverse order so that the decoder reads forward; the stream starts with the flushed state
 * (low word, high word).  One stream per (image, level); symbol j of a stream uses CDF number j / hw (the reference
 * passes indexes = arange(m) expanded over [h, w], cdfSizes = k + 2, offsets = 0: mcquic/modules/entropyCoder.py:114-124).
 * Values outside [0, k) take the reference's 4-bit bypass path (buffered_rans_encoder.cpp:122-160).
 */
#ifndef MCQUIC_ENTROPY_H_
#define MCQUIC_ENTROPY_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MCQ_ENT_PRECISION 16

/* Replaces pmfToQuantizedCDF(pmf, 16) (cpp_exts/ops.cpp:42-111).  pmf: float[k]; cdf_out: uint32[k + 1].
 * Returns 0, or -1 for a negative / non-finite entry or an all-zero pmf (the reference throws std::domain_error). */
int mcq_pmf_to_quantized_cdf(const float* pmf, int32_t k, uint32_t* cdf_out);

/* Upper bound of the bytes one stream of `count` symbols can take. */
int64_t mcq_rans_stream_capacity(int64_t count);

/* Replaces RansEncoder.encodeWithIndexes (cpp_exts/rans_encoder.cpp:49-60 -> buffered_rans_encoder.cpp:104-196) for a
 * whole level at once.  codes: int64 [n, m, hw] (host); cdfs: uint32 [m, k + 1]; out: n slots of `capacity` bytes;
 * out_sizes[n] receives the stream lengths.  n_threads <= 0: automatic -- up to one thread per
 * hardware core and image, but only as many as get 16 Ki symbols each (a level smaller than that is coded by the caller's
 * thread: starting a thread costs more than coding 4 k symbols); the streams do not depend on the thread count.  Returns 0 or -1/-2. */
int mcq_rans_encode_level(const int64_t* codes, int32_t n, int32_t m, int32_t hw, int32_t k, const uint32_t* cdfs,
                          uint8_t* out, int64_t capacity, int32_t* out_sizes, int32_t n_threads);

/* Replaces RansDecoder.decodeWithIndexes (cpp_exts/rans_decoder.cpp:104-173); the per-symbol linear CDF search is a
 * 65536-entry lookup table per CDF (kept across calls, keyed by the CDF's contents; thread-safe).  in: n slots of `stride` bytes holding streams of in_sizes[i] bytes.
 * Returns 0; -1 bad argument / CDF table that does not cover [0, 2^16]; -2 stream length < 8, not a multiple of 4 or
 * > stride; -3 a stream ended before m * hw symbols were decoded (truncated stream, or a header that claims more symbols
 * than the stream holds) -- the decoder never reads past in_sizes[i] bytes of a stream. */
int mcq_rans_decode_level(const uint8_t* in, const int32_t* in_sizes, int64_t stride, int32_t n, int32_t m, int32_t hw,
                          int32_t k, const uint32_t* cdfs, int64_t* codes_out, int32_t n_threads);

int mcq_entropy_version(void);

#ifdef __cplusplus
}
#endif
#endif
