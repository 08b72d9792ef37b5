// Host side of the reference-exact mini-batch sampler: numpy's legacy MT19937 stream restated so that a seeded
// reconstruct() run draws the SAME rows as the reference (core/data.py:15-73,122-176, called every Adam iteration at
// reconstruct.py:114-118), at ~1 ms instead of ~60 ms (numpy, reference) / ~20 ms (round 1) per iteration.
// Plain C++ (g++), no CUDA: linked into libdeepjoin_b200.so next to dj_abi.cu.
//
// numpy pieces restated (numpy/random/mtrand.pyx + _legacy-distributions, stable by numpy's stream-compatibility policy):
//   RandomState.shuffle / permutation : for i = n-1 .. 1: j = random_interval(i); swap(x[i], x[j]);
//       random_interval(max): 32-bit outputs masked to the smallest 2^k - 1 >= max, rejected while > max
//   RandomState.randint(n) / choice(a, size) with replacement : off + masked rejection of 32-bit outputs with
//       rng = n - 1 (random_bounded_uint64_fill, use_masked = True; no output consumed when rng == 0)
//   RandomState.choice(n, size=k, replace=False) : permutation(n)[:k]
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#if defined(__x86_64__)
#include <immintrin.h>
#define DJ_X86 1
#else
#define DJ_X86 0
#endif

using namespace dj;

namespace {

// ---- MT19937, block-wise: one refill produces the next 624 tempered outputs ---------------------------------
struct Stream {
  uint32_t* key;  // numpy's state array (np.random.get_state()[1]), advanced in place
  int pos;        // index of the next output within the current block (numpy's `pos`)
  alignas(64) uint32_t buf[624 + 8];
};

static inline uint32_t temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

#define DJ_MT_REFILL_BODY                                                                                   \
  uint32_t* __restrict__ k = key;                                                                           \
  const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAT = 0x9908b0dfu;                                     \
  int i;                                                                                                    \
  uint32_t y;                                                                                               \
  for (i = 0; i < 624 - 397; i++) { y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT); } \
  for (; i < 623; i++) { y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT); }     \
  y = (k[623] & UP) | (k[0] & LO);                                                                          \
  k[623] = k[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);                                                     \
  for (i = 0; i < 624; i++) out[i] = temper(k[i]);

static void refill_generic(uint32_t* key, uint32_t* __restrict__ out) { DJ_MT_REFILL_BODY }
#if DJ_X86
__attribute__((target("avx2"))) static void refill_avx2(uint32_t* key, uint32_t* __restrict__ out) { DJ_MT_REFILL_BODY }
#endif

static bool has_avx2() {
#if DJ_X86
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
#else
  return false;
#endif
}

static inline void refill(Stream& s) {
#if DJ_X86
  if (has_avx2()) refill_avx2(s.key, s.buf);
  else
#endif
    refill_generic(s.key, s.buf);
  s.pos = 0;
}
static void stream_init(Stream& s, uint32_t* key, int pos) {
  s.key = key;
  s.pos = pos;
  for (int i = pos; i < 624; i++) s.buf[i] = temper(key[i]);
}
static inline uint32_t next32(Stream& s) {
  if (s.pos == 624) refill(s);
  return s.buf[s.pos++];
}
static inline uint32_t mask_of(uint32_t v) {
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v;
}
// numpy: off + buffered_bounded_masked_uint32(rng) (legacy generator: masked rejection)
static inline uint32_t bounded(Stream& s, uint32_t rng) {
  if (rng == 0) return 0;
  if (rng == 0xFFFFFFFFu) return next32(s);
  const uint32_t mask = mask_of(rng);
  uint32_t v;
  while ((v = (next32(s) & mask)) > rng) {}
  return v;
}

// ---- swap partners of RandomState.shuffle on n items ----------------------------------------------------------
// jr[q], q = (n-1) - i, is random_interval(i) for i = n-1 .. 1 (stored in the order they are drawn).
// The acceptance test `candidate <= i` only depends on earlier acceptances through i, which moves by at most one per
// candidate: a block of B candidates none of which lies in (i - (B-1), i] has acceptances that do not depend on each
// other, so the compares leave the sequential chain (scalar: B = 4; AVX2: B = 8 with a compress-store).
static void draw_scalar(Stream& s, uint32_t& i, uint32_t lo, uint32_t mask, uint32_t*& jr) {
  // candidates of the current block while i stays in [lo, mask]
  int p = s.pos;
  while (p + 4 <= 624 && i >= lo + 4) {
    const uint32_t m0 = s.buf[p] & mask, m1 = s.buf[p + 1] & mask, m2 = s.buf[p + 2] & mask, m3 = s.buf[p + 3] & mask;
    const uint32_t i3 = i - 3;
    const uint32_t a0 = m0 <= i, a1 = m1 <= i, a2 = m2 <= i, a3 = m3 <= i;
    if ((uint32_t)(m0 <= i3) + (m1 <= i3) + (m2 <= i3) + (m3 <= i3) != a0 + a1 + a2 + a3) break;
    *jr = m0; jr += a0;
    *jr = m1; jr += a1;
    *jr = m2; jr += a2;
    *jr = m3; jr += a3;
    i -= a0 + a1 + a2 + a3;
    p += 4;
  }
  const int stop = std::min(p + 4, 624);
  for (; p < stop && i >= lo; p++) {
    const uint32_t v = s.buf[p] & mask;
    const uint32_t a = v <= i;
    *jr = v;
    jr += a;
    i -= a;
  }
  s.pos = p;
}

#if DJ_X86
alignas(32) static uint32_t g_compress[256][8];
static void init_compress() {
  static bool done = false;
  static std::mutex m;
  std::lock_guard<std::mutex> lk(m);
  if (done) return;
  for (int a = 0; a < 256; a++) {
    int c = 0;
    for (int b = 0; b < 8; b++)
      if ((a >> b) & 1) g_compress[a][c++] = (uint32_t)b;
    for (; c < 8; c++) g_compress[a][c] = 0;
  }
  done = true;
}
__attribute__((target("avx2,popcnt"))) static void draw_avx2(Stream& s, uint32_t& i, uint32_t lo, uint32_t mask, uint32_t*& jr) {
  int p = s.pos;
  const __m256i vmask = _mm256_set1_epi32((int)mask);
  // all values are < 2^31 (n <= 2^31 - 1): signed compares are exact
  while (p + 8 <= 624 && i >= lo + 8) {
    const __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s.buf + p)), vmask);
    const int rej = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v, _mm256_set1_epi32((int)i))));
    const int rej7 = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v, _mm256_set1_epi32((int)(i - 7)))));
    if (rej != rej7) break;  // a candidate in (i - 7, i]: exact scalar steps below
    const int acc = (~rej) & 0xff;
    const __m256i idx = _mm256_load_si256(reinterpret_cast<const __m256i*>(g_compress[acc]));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(jr), _mm256_permutevar8x32_epi32(v, idx));  // 8 lanes written, cnt valid
    const uint32_t cnt = (uint32_t)_mm_popcnt_u32((unsigned)acc);
    jr += cnt;
    i -= cnt;
    p += 8;
  }
  const int stop = std::min(p + 8, 624);
  for (; p < stop && i >= lo; p++) {
    const uint32_t v = s.buf[p] & mask;
    const uint32_t a = v <= i;
    *jr = v;
    jr += a;
    i -= a;
  }
  s.pos = p;
}
#endif

// jr must have room for n + 8 entries
static void draw_shuffle(Stream& s, int64_t n, uint32_t* jr) {
  if (n < 2) return;
  uint32_t i = (uint32_t)(n - 1);
  const bool avx2 = has_avx2();
  while (i >= 1) {
    const uint32_t mask = mask_of(i);
    const uint32_t lo = (mask >> 1) + 1;  // every i in [lo, mask] draws with this mask
    while (i >= lo) {
      if (s.pos == 624) refill(s);
#if DJ_X86
      if (avx2) draw_avx2(s, i, lo, mask, jr);
      else
#endif
        draw_scalar(s, i, lo, mask, jr);
    }
  }
  (void)avx2;
}

template <class T>
static void apply_shuffle(T* w, int64_t n, const uint32_t* jr) {
  constexpr int B = 16;
  for (int64_t i = n - 1, q = 0; i >= 1; i--, q++) {
    if (i - B >= 1) __builtin_prefetch(&w[jr[q + B]], 1);
    const uint32_t j = jr[q];
    const T t = w[j];
    w[j] = w[i];
    w[i] = t;
  }
}

struct Job {
  int e = 0;
  std::vector<int32_t> rows;      // rows before the second shuffle
  std::vector<uint32_t> jr;       // its swap partners
  int64_t n_pos = 0, n_neg = 0;
  int64_t start_pos = -1, start_neg = -1;        // window mode (core/data.py:53-55, 63-65)
  std::vector<uint32_t> pick_pos, pick_neg;      // replacement mode (:57, :67): indices into the class lists
};

static void finish(Job& J, const float* sdf, int num_samples, int32_t* out) {
  int32_t* w = J.rows.data();
  const int64_t n = (int64_t)J.rows.size();
  apply_shuffle(w, n, J.jr.data());
  const int64_t pos_num = num_samples / 2, neg_num = num_samples - pos_num;
  int32_t* o_pos = out;
  int32_t* o_neg = out + pos_num;
  if (J.start_pos >= 0 && J.start_neg >= 0) {  // both classes by window: one pass, stop when both are full
    int64_t cp = 0, cn = 0;
    const int64_t p_end = J.start_pos + pos_num, n_end = J.start_neg + neg_num;
    for (int64_t p = 0; p < n && (cp < p_end || cn < n_end); p++) {
      const float v = sdf[w[p]];
      if (v > 0.f) { if (cp >= J.start_pos && cp < p_end) o_pos[cp - J.start_pos] = w[p]; cp++; }
      else if (v <= 0.f) { if (cn >= J.start_neg && cn < n_end) o_neg[cn - J.start_neg] = w[p]; cn++; }
    }
    return;
  }
  std::vector<int32_t> lp, ln;  // a class drawn with replacement: the class lists are needed
  lp.reserve((size_t)J.n_pos);
  ln.reserve((size_t)J.n_neg);
  for (int64_t p = 0; p < n; p++) {
    const float v = sdf[w[p]];
    if (v > 0.f) lp.push_back(w[p]);
    else if (v <= 0.f) ln.push_back(w[p]);
  }
  if (J.start_pos >= 0) for (int64_t i = 0; i < pos_num; i++) o_pos[i] = lp[(size_t)(J.start_pos + i)];
  else for (int64_t i = 0; i < pos_num; i++) o_pos[i] = lp[J.pick_pos[(size_t)i]];
  if (J.start_neg >= 0) for (int64_t i = 0; i < neg_num; i++) o_neg[i] = ln[(size_t)(J.start_neg + i)];
  else for (int64_t i = 0; i < neg_num; i++) o_neg[i] = ln[J.pick_neg[(size_t)i]];
}

}  // namespace

extern "C" int dj_host_mt19937_permutation(uint32_t* key624, int32_t* pos, int64_t n, int64_t* out) {
  if (!key624 || !pos || !out || n < 0 || *pos < 0 || *pos > 624) return fail(DJ_ERR_INVALID, "dj_host_mt19937_permutation: bad argument");
  if (n > 0x7fffffffll) return fail(DJ_ERR_UNSUPPORTED, "dj_host_mt19937_permutation: n > 2^31 - 1");
#if DJ_X86
  init_compress();
#endif
  Stream s;
  stream_init(s, key624, *pos);
  std::vector<uint32_t> jr((size_t)n + 8);
  draw_shuffle(s, n, jr.data());
  for (int64_t i = 0; i < n; i++) out[i] = i;
  apply_shuffle(out, n, jr.data());
  *pos = s.pos;
  return DJ_OK;
}

// The reference's per-iteration sampler for ALL iterations of one reconstruct() call.  Per iteration numpy draws
//   permutation(half)[:k]                (np.random.choice(half, k, replace=False), core/data.py:141/:153)
//   permutation(len(rows))               (:175)
//   randint(len(neg) - neg_num), randint(len(pos) - pos_num), or choice(.., n) with replacement   (:53-71)
// The only thing that serialises the iterations is the generator: how many words a bounded draw consumes depends on
// its range, and the range of the window draws depends on how many of the selected uniform rows are inside the
// shape.  So the calling thread walks the stream -- the swap partners of both shuffles, the first shuffle itself,
// the class counts, the window starts -- and hands everything that is independent from there (second shuffle,
// locating the windows, writing the rows) to worker threads.
// progress (optional): *progress = number of LEADING iterations whose rows are final, stored with release order as
// the workers finish -- a consumer thread can hand rows [0, *progress) to the device while the rest is still drawn.
extern "C" int dj_host_sample_schedule(uint32_t* key624, int32_t* pos, const float* sdf, int64_t n_pts, double uniform_ratio,
                                       int num_iterations, int num_samples, int n_threads, int32_t* out, int32_t* progress) {
  if (!key624 || !pos || !sdf || !out || n_pts < 2 || num_iterations < 0 || num_samples < 1 || *pos < 0 || *pos > 624)
    return fail(DJ_ERR_INVALID, "dj_host_sample_schedule: bad argument");
  if (n_pts > 0x7fffffffll) return fail(DJ_ERR_UNSUPPORTED, "dj_host_sample_schedule: more than 2^31 - 1 rows");
  if (!(uniform_ratio > 0.0 && uniform_ratio < 1.0)) return fail(DJ_ERR_INVALID, "dj_host_sample_schedule: uniform_ratio must be in (0, 1)");
#if DJ_X86
  init_compress();
#endif
  Stream S;
  stream_init(S, key624, *pos);
  const int64_t half = (int64_t)((double)n_pts / 2.0);  // int(pts.shape[0] / 2), core/data.py:137
  const int64_t pos_num = num_samples / 2, neg_num = num_samples - pos_num;
  // the fixed part of the rows: the surface half for ratio < 0.5, the uniform half above it, every row at 0.5
  int64_t k = 0, fixed_lo = 0, fixed_hi = n_pts, sel_off = 0;
  const bool ratio_step = uniform_ratio != 0.5;
  if (ratio_step) {
    if (uniform_ratio > 0.5) { k = (int64_t)(((double)half * (1.0 - uniform_ratio)) / uniform_ratio); fixed_lo = half; fixed_hi = 2 * half; sel_off = 0; }
    else { k = (int64_t)(((double)half * uniform_ratio) / (1.0 - uniform_ratio)); fixed_lo = 0; fixed_hi = half; sel_off = half; }
    if (k > half) return fail(DJ_ERR_INVALID, "Cannot take a larger sample than population when 'replace=False'");
  }
  int64_t fixed_pos = 0, fixed_neg = 0;
  for (int64_t i = fixed_lo; i < fixed_hi; i++) { fixed_pos += sdf[i] > 0.f; fixed_neg += sdf[i] <= 0.f; }
  const int64_t n_rows = ratio_step ? (half + k) : n_pts;

  const int hw = (int)std::thread::hardware_concurrency();
  const int nw = std::max(1, std::min(n_threads > 0 ? n_threads : hw - 1, 16));
  std::mutex mu;
  std::condition_variable cv_job, cv_room;
  std::deque<std::unique_ptr<Job>> queue, spare;  // spare: finished jobs whose buffers the next iterations reuse
  bool done = false;
  std::vector<char> finished((size_t)std::max(num_iterations, 1), 0);
  int lead = 0;  // iterations [0, lead) are complete (under mu)
  if (progress) __atomic_store_n(progress, 0, __ATOMIC_RELEASE);
  std::vector<std::thread> workers;
  for (int t = 0; t < nw; t++)
    workers.emplace_back([&] {
      for (;;) {
        std::unique_ptr<Job> J;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv_job.wait(lk, [&] { return done || !queue.empty(); });
          if (queue.empty()) return;
          J = std::move(queue.front());
          queue.pop_front();
        }
        finish(*J, sdf, num_samples, out + (size_t)J->e * num_samples);
        {
          std::lock_guard<std::mutex> lk(mu);
          finished[(size_t)J->e] = 1;
          while (lead < num_iterations && finished[(size_t)lead]) lead++;
          if (progress) __atomic_store_n(progress, lead, __ATOMIC_RELEASE);
          spare.push_back(std::move(J));
        }
        cv_room.notify_one();
      }
    });
  int rc = DJ_OK;
  size_t in_flight = 0;
  std::vector<int32_t> perm1((size_t)std::max<int64_t>(half, 1));
  std::vector<uint32_t> jr1((size_t)half + 8);
  for (int e = 0; e < num_iterations && rc == DJ_OK; e++) {
    std::unique_ptr<Job> J;
    {
      std::unique_lock<std::mutex> lk(mu);
      if (!spare.empty()) { J = std::move(spare.front()); spare.pop_front(); }
      else if ((int)in_flight >= 2 * nw + 2) {  // bound the memory: wait for a finished job
        cv_room.wait(lk, [&] { return !spare.empty(); });
        J = std::move(spare.front());
        spare.pop_front();
      }
    }
    if (!J) { J.reset(new Job()); in_flight++; }
    J->e = e;
    J->start_pos = J->start_neg = -1;
    J->pick_pos.clear();
    J->pick_neg.clear();
    J->rows.resize((size_t)n_rows);
    int64_t n_pos = fixed_pos, n_neg = fixed_neg;
    int32_t* r = J->rows.data();
    if (ratio_step) {
      draw_shuffle(S, half, jr1.data());
      for (int64_t i = 0; i < half; i++) perm1[(size_t)i] = (int32_t)i;
      apply_shuffle(perm1.data(), half, jr1.data());
      const int32_t* sel;
      if (uniform_ratio > 0.5) {  // [selected surface rows | uniform half]
        for (int64_t i = 0; i < k; i++) r[i] = perm1[(size_t)i];
        for (int64_t i = 0; i < half; i++) r[k + i] = (int32_t)(half + i);
        sel = r;
      } else {                    // [surface half | selected uniform rows]
        for (int64_t i = 0; i < half; i++) r[i] = (int32_t)i;
        for (int64_t i = 0; i < k; i++) r[half + i] = (int32_t)(perm1[(size_t)i] + sel_off);
        sel = r + half;
      }
      for (int64_t i = 0; i < k; i++) { const float v = sdf[sel[i]]; n_pos += v > 0.f; n_neg += v <= 0.f; }
    } else {
      for (int64_t i = 0; i < n_rows; i++) r[i] = (int32_t)i;
    }
    J->jr.resize((size_t)n_rows + 8);
    draw_shuffle(S, n_rows, J->jr.data());
    J->n_pos = n_pos;
    J->n_neg = n_neg;
    // negatives first (core/data.py:53-61), then positives (:63-71)
    if (n_neg > neg_num) J->start_neg = bounded(S, (uint32_t)(n_neg - neg_num - 1));
    else if (n_neg == 0) rc = fail(DJ_ERR_NO_SAMPLES, "NoSamplesError: no sample with sdf <= 0");
    else { J->pick_neg.resize((size_t)neg_num); for (auto& q : J->pick_neg) q = bounded(S, (uint32_t)(n_neg - 1)); }
    if (rc == DJ_OK) {
      if (n_pos > pos_num) J->start_pos = bounded(S, (uint32_t)(n_pos - pos_num - 1));
      else if (n_pos == 0 && pos_num > 0) rc = fail(DJ_ERR_NO_SAMPLES, "NoSamplesError: no sample with sdf > 0");
      else { J->pick_pos.resize((size_t)pos_num); for (auto& q : J->pick_pos) q = bounded(S, (uint32_t)(n_pos - 1)); }
    }
    if (rc != DJ_OK) break;
    {
      std::lock_guard<std::mutex> lk(mu);
      queue.push_back(std::move(J));
    }
    cv_job.notify_one();
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    done = true;
  }
  cv_job.notify_all();
  for (auto& t : workers) t.join();
  *pos = S.pos;
  return rc;
}
