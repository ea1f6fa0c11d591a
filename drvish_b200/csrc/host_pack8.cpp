// Lossless 8-bit transport of count tiles for the host->device copy (host side; plain C++ so that the
// vector intrinsics are compiled by g++, not nvcc's front end).
//
// The reference ships every minibatch as a dense fp32 tensor (`xs = (x.cuda() for x in xs)`,
// src/drvish/train/__init__.py:42-43): 82 MB per 4096 x 5000 tile, 1.5 ms at the measured PCIe rate - more than
// twice the whole ELBO step.  scRNA counts are small non-negative integers, so a tile crosses the bus as ONE
// BYTE per entry: an integer 0..254 is its own byte; anything else (a count >= 255, and for that matter a
// non-integer or negative value) becomes the byte 255 plus an (index, fp32 value) pair in an escape list.  The
// device widens the bytes back to the fp32 tile the step reads and patches the escapes (drv_unpack_counts_u8):
// bit-exact for ANY fp32 tile.  A tile with more than `max_escapes` escapes is refused (it then travels in
// another format).
//
// drv_host_pack_counts_u8 runs on a persistent pool of worker threads (created on first use; spawning
// threads per call cost more than the packing), each packing a 64-byte-aligned slice with AVX-512 / AVX2 /
// scalar code chosen at run time.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Escapes {
  std::vector<uint32_t> idx;
  std::vector<float> val;
};

inline bool byte_ok(float v) { return v >= 0.f && v <= 254.f && (float)(int)v == v; }

void pack_scalar(const float* src, uint8_t* dst, int64_t base, int64_t n, Escapes& e) {
  for (int64_t i = 0; i < n; ++i) {
    const float v = src[i];
    if (byte_ok(v)) {
      dst[i] = (uint8_t)(int)v;
    } else {
      dst[i] = 255;
      e.idx.push_back((uint32_t)(base + i));
      e.val.push_back(v);
    }
  }
}

__attribute__((target("avx2"))) void pack_avx2(const float* src, uint8_t* dst, int64_t base, int64_t n, Escapes& e) {
  const __m256i lim = _mm256_set1_epi32(254);
  const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  int64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i iv[4];
    unsigned bad = 0;
    for (int k = 0; k < 4; ++k) {
      const __m256 v = _mm256_loadu_ps(src + i + 8 * k);
      iv[k] = _mm256_cvttps_epi32(v);
      const __m256 back = _mm256_cvtepi32_ps(iv[k]);
      const __m256i eq = _mm256_castps_si256(_mm256_cmp_ps(back, v, _CMP_EQ_OQ));
      // 0 <= iv <= 254 as one unsigned test: max_epu32(iv, 254) == 254
      const __m256i in = _mm256_cmpeq_epi32(_mm256_max_epu32(iv[k], lim), lim);
      const unsigned okm = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_and_si256(eq, in)));
      bad |= (~okm & 0xffu) << (8 * k);
      iv[k] = _mm256_blendv_epi8(_mm256_set1_epi32(255), iv[k], _mm256_and_si256(eq, in));
    }
    // 32 x int32 -> 32 x uint8 (values are 0..255): packus twice, then undo the per-lane interleave
    const __m256i w01 = _mm256_packus_epi32(iv[0], iv[1]);
    const __m256i w23 = _mm256_packus_epi32(iv[2], iv[3]);
    const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(w01, w23), perm);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), b);
    while (bad) {
      const int j = __builtin_ctz(bad);
      bad &= bad - 1;
      e.idx.push_back((uint32_t)(base + i + j));
      e.val.push_back(src[i + j]);
    }
  }
  pack_scalar(src + i, dst + i, base + i, n - i, e);
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) void pack_avx512(const float* src, uint8_t* dst, int64_t base, int64_t n,
                                                                     Escapes& e) {
  const __m512i lim = _mm512_set1_epi32(254);
  const bool stream = (reinterpret_cast<uintptr_t>(dst) & 63) == 0;
  int64_t i = 0;
  for (; i + 64 <= n; i += 64) {
    __m128i q[4];
    uint64_t bad = 0;
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 512), _MM_HINT_NTA);  // 2 KB ahead: the source is streamed once
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 528), _MM_HINT_NTA);
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 544), _MM_HINT_NTA);
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 560), _MM_HINT_NTA);
    for (int k = 0; k < 4; ++k) {
      const __m512 v = _mm512_loadu_ps(src + i + 16 * k);
      const __m512i iv = _mm512_cvttps_epi32(v);
      const __mmask16 ok = _mm512_cmp_ps_mask(_mm512_cvtepi32_ps(iv), v, _CMP_EQ_OQ) & _mm512_cmple_epu32_mask(iv, lim);
      q[k] = _mm512_cvtusepi32_epi8(_mm512_mask_mov_epi32(_mm512_set1_epi32(255), ok, iv));
      bad |= (uint64_t)(uint16_t)~ok << (16 * k);
    }
    __m512i out = _mm512_castsi128_si512(q[0]);
    out = _mm512_inserti32x4(out, q[1], 1);
    out = _mm512_inserti32x4(out, q[2], 2);
    out = _mm512_inserti32x4(out, q[3], 3);
    if (stream) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), out);  // the packed tile is only read by the DMA engine
    else _mm512_storeu_si512(dst + i, out);
    while (bad) {
      const int j = __builtin_ctzll(bad);
      bad &= bad - 1;
      e.idx.push_back((uint32_t)(base + i + j));
      e.val.push_back(src[i + j]);
    }
  }
  if (stream) _mm_sfence();
  pack_scalar(src + i, dst + i, base + i, n - i, e);
}

using PackFn = void (*)(const float*, uint8_t*, int64_t, int64_t, Escapes&);

PackFn pick() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) return pack_avx512;
  if (__builtin_cpu_supports("avx2")) return pack_avx2;
  return pack_scalar;
}

// Persistent fork-join pool: run(n_tasks, fn) executes fn(0..n_tasks-1) on the workers and returns when all are done.
class Pool {
 public:
  explicit Pool(int n) {
    for (int t = 0; t < n; ++t) th_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  // Tasks are PULLED (atomic counter), not dealt out: a worker whose core is taken by another thread of the process
  // (the Python main thread polls the GPU, the prefetcher's worker issues copies) does not hold its share back, the
  // others take it over; the calling thread works too.
  void run(int n_tasks, const std::function<void(int)>& fn) {
    {
      std::unique_lock<std::mutex> g(m_);
      fn_ = &fn;
      n_tasks_ = n_tasks;
      next_.store(0, std::memory_order_relaxed);
      pending_ = (int)th_.size();
      ++epoch_;
    }
    cv_.notify_all();
    for (int k; (k = next_.fetch_add(1, std::memory_order_relaxed)) < n_tasks;) fn(k);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      int n_tasks;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        fn = fn_;
        n_tasks = n_tasks_;
      }
      (void)t;
      for (int k; (k = next_.fetch_add(1, std::memory_order_relaxed)) < n_tasks;) (*fn)(k);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_tasks_ = 0, pending_ = 0;
  std::atomic<int> next_{0};
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

std::mutex g_pool_mutex;
Pool* g_pool = nullptr;  // leaked on purpose: worker threads must not be joined from a static destructor at exit

}  // namespace

// h_src: n fp32 values; h_dst: n bytes (64-byte aligned for the streaming stores); h_esc_idx / h_esc_val: room for
// max_escapes entries.  Returns the number of escapes (>= 0), or -1 for invalid arguments, or -2 when the tile has
// more than max_escapes escapes (h_dst is then incomplete).  n must be < 2^32.  Calls are serialised.
extern "C" int64_t drv_host_pack_counts_u8(const float* h_src, uint8_t* h_dst, int64_t n, uint32_t* h_esc_idx, float* h_esc_val,
                                           int64_t max_escapes, int n_threads) {
  if (!h_src || !h_dst || n < 0 || n >= ((int64_t)1 << 32) || max_escapes < 0 || (max_escapes > 0 && (!h_esc_idx || !h_esc_val)))
    return -1;
  static const PackFn fn = pick();
  if (n_threads < 1) n_threads = 1;
  const int64_t min_chunk = 1 << 15;
  int n_tasks = (int)((n + min_chunk - 1) / min_chunk);
  if (n_tasks < 1) n_tasks = 1;
  std::lock_guard<std::mutex> serial(g_pool_mutex);
  if (n_threads > 1 && (!g_pool || g_pool->size() != n_threads)) {
    delete g_pool;
    g_pool = new Pool(n_threads);
  }
  // slices of whole 64-byte output lines, a few per thread so that a slow core does not hold the others up
  const int want = n_threads > 1 ? 8 * n_threads : 1;
  if (n_tasks > want) n_tasks = want;
  const int64_t chunk = ((n + n_tasks - 1) / n_tasks + 63) / 64 * 64;
  std::vector<Escapes> esc((size_t)n_tasks);
  const std::function<void(int)> task = [&](int k) {
    const int64_t a = (int64_t)k * chunk, b = a + chunk < n ? a + chunk : n;
    if (a < b) fn(h_src + a, h_dst + a, a, b - a, esc[(size_t)k]);
  };
  if (n_threads > 1 && n_tasks > 1) g_pool->run(n_tasks, task);
  else
    for (int k = 0; k < n_tasks; ++k) task(k);
  int64_t total = 0;
  for (const auto& e : esc) total += (int64_t)e.idx.size();
  if (total > max_escapes) return -2;
  int64_t o = 0;
  for (const auto& e : esc) {
    if (e.idx.empty()) continue;
    memcpy(h_esc_idx + o, e.idx.data(), e.idx.size() * sizeof(uint32_t));
    memcpy(h_esc_val + o, e.val.data(), e.val.size() * sizeof(float));
    o += (int64_t)e.idx.size();
  }
  return total;
}

// "avx512" / "avx2" / "scalar": which packer this host runs (bench.py prints it)
extern "C" const char* drv_host_pack_isa(void) {
  const PackFn fn = pick();
  return fn == pack_avx512 ? "avx512" : fn == pack_avx2 ? "avx2" : "scalar";
}
