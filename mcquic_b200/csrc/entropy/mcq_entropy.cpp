// Host rANS coder for McQuic code maps (see include/mcquic_entropy.h).  Written from the published rANS scheme
// (Duda 2013; 64-bit state / 32-bit word variant with lower bound L = 2^31) and the reference's stream conventions;
// bit-compatibility with the reference coder is enforced by tests/test_entropy.py against oracle/_ref.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/mcquic_entropy.h"

namespace {

constexpr uint32_t kPrec = MCQ_ENT_PRECISION;
constexpr uint64_t kLow = 1ull << 31;          // normalisation interval [L, L * 2^32)
constexpr uint32_t kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;

struct Slot {        // one coding step: an interval [start, start + range) of 2^16, or `bits` raw bits
  uint16_t start, range;
  uint8_t raw;       // 1: bypass step of kBypassBits bits carrying `start`
};

// state transition x -> C(s, x); emits one 32-bit word (towards lower addresses) when x would overflow
inline void put_interval(uint64_t& x, uint32_t*& w, uint32_t start, uint32_t range, uint32_t bits) {
  const uint64_t limit = ((kLow >> bits) << 32) * range;
  if (x >= limit) {
    *--w = (uint32_t)x;
    x >>= 32;
  }
  x = ((x / range) << bits) + (x % range) + start;
}

inline void put_raw(uint64_t& x, uint32_t*& w, uint32_t value) {
  const uint64_t limit = ((kLow >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits));
  if (x >= limit) {
    *--w = (uint32_t)x;
    x >>= 32;
  }
  x = (x << kBypassBits) | value;
}

// Every stream word is read through next_word(): a truncated or crafted stream (symbol count from an untrusted header)
// runs out of words instead of reading past the buffer -- `bad` is set and zeros are fed from then on.
struct Reader {
  const uint32_t* r;
  const uint32_t* end;
  bool bad = false;
  uint32_t next_word() {
    if (r < end) return *r++;
    bad = true;
    return 0u;
  }
};

inline uint32_t get_raw(uint64_t& x, Reader& rd) {
  const uint32_t v = (uint32_t)x & kBypassMax;
  x >>= kBypassBits;
  if (x < kLow) x = (x << 32) | rd.next_word();
  return v;
}

// encodes one stream; returns the number of bytes written at the END of [buf, buf + cap_words)
int64_t encode_stream(const int64_t* sym, int m, int hw, int k, const uint32_t* cdfs, uint32_t* buf, int64_t cap_words,
                      std::vector<Slot>& slots) {
  slots.clear();
  const int cdf_len = k + 1;
  const int64_t sentinel = k;   // the reference's max_value = cdfSizes - 2 with cdfSizes = k + 2
  for (int mi = 0; mi < m; ++mi) {
    const uint32_t* cdf = cdfs + (size_t)mi * cdf_len;
    for (int j = 0; j < hw; ++j) {
      int64_t v = sym[(size_t)mi * hw + j];
      uint32_t raw = 0;
      if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = sentinel; }
      else if (v >= sentinel) { raw = (uint32_t)(2 * (v - sentinel)); v = sentinel; }
      // v == sentinel indexes cdf[k + 1], one past the table, exactly as the reference does under -DNDEBUG; guard it
      const uint32_t lo = v < cdf_len ? cdf[v] : (1u << kPrec);
      const uint32_t hi = v + 1 < cdf_len ? cdf[v + 1] : (1u << kPrec);
      slots.push_back({(uint16_t)lo, (uint16_t)(hi - lo), 0});
      if (v == sentinel) {
        int nb = 0;
        while ((raw >> (nb * kBypassBits)) != 0) ++nb;
        int left = nb;
        while (left >= (int)kBypassMax) { slots.push_back({(uint16_t)kBypassMax, 0, 1}); left -= kBypassMax; }
        slots.push_back({(uint16_t)left, 0, 1});
        for (int b = 0; b < nb; ++b) slots.push_back({(uint16_t)((raw >> (b * kBypassBits)) & kBypassMax), 0, 1});
      }
    }
  }
  if ((int64_t)slots.size() + 2 > cap_words) return -1;
  uint64_t x = kLow;
  uint32_t* w = buf + cap_words;
  for (size_t i = slots.size(); i-- > 0;) {
    const Slot& s = slots[i];
    if (s.raw) put_raw(x, w, s.start);
    else {
      if (s.range == 0) return -2;   // zero-probability symbol: not encodable
      put_interval(x, w, s.start, s.range, kPrec);
    }
  }
  w -= 2;
  w[0] = (uint32_t)x;
  w[1] = (uint32_t)(x >> 32);
  return (int64_t)((buf + cap_words) - w) * 4;
}

// returns 0, or -3 when the stream ends before m * hw symbols were decoded (truncated / header does not match the stream)
int decode_stream(const uint32_t* r, const uint32_t* end, int m, int hw, int k, const uint32_t* cdfs,
                  const uint16_t* const* luts, int64_t* out) {
  if (end - r < 2) return -3;
  uint64_t x = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
  Reader rd{r + 2, end};
  const int cdf_len = k + 1;
  const uint64_t mask = (1ull << kPrec) - 1;
  for (int mi = 0; mi < m; ++mi) {
    const uint32_t* cdf = cdfs + (size_t)mi * cdf_len;
    const uint16_t* lut = luts[mi];
    for (int j = 0; j < hw; ++j) {
      const uint32_t cum = (uint32_t)(x & mask);
      const uint32_t s = lut[cum];
      const uint32_t lo = cdf[s], range = cdf[s + 1] - lo;
      x = (uint64_t)range * (x >> kPrec) + (x & mask) - lo;
      if (x < kLow) x = (x << 32) | rd.next_word();
      int64_t v = s;
      if ((int)s == k) {   // bypass (never produced for codes in [0, k))
        uint32_t val = get_raw(x, rd);
        int nb = (int)val;
        while (val == kBypassMax && !rd.bad) { val = get_raw(x, rd); nb += (int)val; }
        uint32_t raw = 0;
        for (int b = 0; b < nb && b < 8; ++b) raw |= get_raw(x, rd) << (b * kBypassBits);
        v = raw >> 1;
        v = (raw & 1) ? -v - 1 : v + k;
      }
      if (rd.bad) return -3;
      out[(size_t)mi * hw + j] = v;
    }
  }
  return 0;
}

// n_threads > 0: that many threads.  0 = automatic: as many as pay for their own start-up -- creating and joining a
// std::thread costs 50-100 us, about what coding 4 k symbols takes, so a thread has to get kSymbolsPerThread of them
// (measured on the `Validator.speed` shapes, 10 x 3 x 768 x 512: the three levels of a batch hold 20 k symbols;
// one thread codes them in 0.46 ms, eight threads in 1.2 ms)
constexpr int64_t kSymbolsPerThread = 16384;
// cum_freq -> symbol table of one codebook (the reference scans the CDF linearly for every symbol).  Filling its 2^16
// entries costs as much as decoding a few thousand symbols, and a model decodes with the same handful of CDFs call after
// call: the tables are kept, keyed by the CDF's contents (a changed CDF -- frequency EMA update -- simply misses).
struct Lut {
  std::vector<uint32_t> cdf;
  std::vector<uint16_t> table;
};
constexpr size_t kLutCacheEntries = 32;      // 32 x (k + 1 words + 128 KB)

// nullptr: the CDF does not cover [0, 2^16) monotonically
std::shared_ptr<const Lut> lut_for(const uint32_t* cdf, int k) {
  static std::mutex mu;
  static std::vector<std::shared_ptr<const Lut>> cache;
  static size_t next = 0;
  {
    std::lock_guard<std::mutex> g(mu);
    for (const auto& e : cache)
      if ((int)e->cdf.size() == k + 1 && std::memcmp(e->cdf.data(), cdf, sizeof(uint32_t) * (size_t)(k + 1)) == 0) return e;
  }
  if (cdf[0] != 0 || cdf[k] != (1u << kPrec)) return nullptr;     // the table must cover [0, 2^16) completely
  auto lut = std::make_shared<Lut>();
  lut->cdf.assign(cdf, cdf + k + 1);
  lut->table.resize((size_t)1 << kPrec);
  for (int s = 0; s < k; ++s) {
    if (cdf[s + 1] < cdf[s] || cdf[s + 1] > (1u << kPrec)) return nullptr;
    for (uint32_t c = cdf[s]; c < cdf[s + 1]; ++c) lut->table[c] = (uint16_t)s;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() < kLutCacheEntries) cache.push_back(lut);
  else { cache[next] = lut; next = (next + 1) % kLutCacheEntries; }
  return lut;
}

template <class F>
void parallel_for(int n, int n_threads, int64_t symbols_per_item, F fn) {
  int t = n_threads;
  if (t <= 0) {
    t = (int)std::thread::hardware_concurrency();
    const int64_t worth = (int64_t)n * symbols_per_item / kSymbolsPerThread;
    if (worth < t) t = (int)worth;
  }
  t = std::max(1, std::min(t, n));
  if (t == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::vector<std::thread> pool;
  for (int k = 0; k < t; ++k)
    pool.emplace_back([=]() { for (int i = k; i < n; i += t) fn(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int mcq_pmf_to_quantized_cdf(const float* pmf, int32_t k, uint32_t* cdf) {
  if (!pmf || !cdf || k <= 0) return -1;
  for (int i = 0; i < k; ++i)
    if (pmf[i] < 0 || !std::isfinite(pmf[i])) return -1;
  const uint32_t one = 1u << kPrec;
  cdf[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < k; ++i) {
    cdf[i + 1] = (uint32_t)std::round(pmf[i] * (float)one);   // float arithmetic, like the reference's lambda
    total += cdf[i + 1];
  }
  if (total == 0) return -1;
  for (int i = 0; i <= k; ++i) cdf[i] = (uint32_t)(((uint64_t)one * cdf[i]) / total);
  for (int i = 1; i <= k; ++i) cdf[i] += cdf[i - 1];
  cdf[k] = one;
  // every symbol needs a non-empty interval: take one count from the smallest interval that can spare it
  for (int i = 0; i < k; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best = ~0u;
    int donor = -1;
    for (int j = 0; j < k; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best) { best = f; donor = j; }
    }
    if (donor < 0) return -1;
    if (donor < i) { for (int j = donor + 1; j <= i; ++j) cdf[j]--; }
    else { for (int j = i + 1; j <= donor; ++j) cdf[j]++; }
  }
  return 0;
}

int64_t mcq_rans_stream_capacity(int64_t count) { return (count + 4) * 4; }

int mcq_rans_encode_level(const int64_t* codes, int32_t n, int32_t m, int32_t hw, int32_t k, const uint32_t* cdfs,
                          uint8_t* out, int64_t capacity, int32_t* out_sizes, int32_t n_threads) {
  if (!codes || !cdfs || !out || !out_sizes || n <= 0 || m <= 0 || hw <= 0 || k <= 0 || k > 65535) return -1;
  if (capacity % 4 != 0 || capacity < 16) return -1;
  std::vector<int> status(n, 0);
  parallel_for(n, n_threads, (int64_t)m * hw, [&](int i) {
    thread_local std::vector<Slot> slots;
    thread_local std::vector<uint32_t> words;
    const int64_t cap_words = capacity / 4;
    words.resize((size_t)cap_words);
    const int64_t bytes = encode_stream(codes + (size_t)i * m * hw, m, hw, k, cdfs, words.data(), cap_words, slots);
    if (bytes < 0) { status[i] = (int)bytes; out_sizes[i] = 0; return; }
    std::memcpy(out + (size_t)i * capacity, (const uint8_t*)(words.data() + cap_words) - bytes, (size_t)bytes);
    out_sizes[i] = (int32_t)bytes;
  });
  for (int s : status) if (s) return s;
  return 0;
}

int mcq_rans_decode_level(const uint8_t* in, const int32_t* in_sizes, int64_t stride, int32_t n, int32_t m, int32_t hw,
                          int32_t k, const uint32_t* cdfs, int64_t* codes_out, int32_t n_threads) {
  if (!in || !in_sizes || !cdfs || !codes_out || n <= 0 || m <= 0 || hw <= 0 || k <= 0 || k > 65535) return -1;
  std::vector<std::shared_ptr<const Lut>> held((size_t)m);     // keeps the tables alive while other calls recycle the cache
  std::vector<const uint16_t*> luts((size_t)m);
  for (int mi = 0; mi < m; ++mi) {
    held[mi] = lut_for(cdfs + (size_t)mi * (k + 1), k);
    if (!held[mi]) return -1;
    luts[mi] = held[mi]->table.data();
  }
  for (int i = 0; i < n; ++i)
    if (in_sizes[i] < 8 || in_sizes[i] % 4 != 0 || in_sizes[i] > stride) return -2;
  std::vector<int> status((size_t)n, 0);
  parallel_for(n, n_threads, (int64_t)m * hw, [&](int i) {
    const size_t nwords = (size_t)in_sizes[i] / 4;
    std::vector<uint32_t> words(nwords, 0u);   // aligned copy
    std::memcpy(words.data(), in + (size_t)i * stride, (size_t)in_sizes[i]);
    status[i] = decode_stream(words.data(), words.data() + nwords, m, hw, k, cdfs, luts.data(),
                              codes_out + (size_t)i * m * hw);
  });
  for (int s : status) if (s) return s;
  return 0;
}

int mcq_entropy_version(void) { return 1; }

}  // extern "C"
