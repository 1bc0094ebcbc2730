// Host-side genotype feeder (feeder.hpp): worker pool, pinned slot cache, narrowing loops.  Plain C++ (g++), no device code.
#include "feeder.hpp"

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <functional>
#include <thread>

namespace crm {

size_t host_dtype_size(int dtype) {
    switch (dtype) {
        case HD_F64: case HD_I64: return 8;
        case HD_I8: case HD_U8: return 1;
        case HD_I16: return 2;
        case HD_I32: case HD_F32: return 4;
        default: return 0;
    }
}

int host_threads() {
    static const int n = [] {
        const char* env = getenv("CRM_HOST_THREADS");
        if (env && atoi(env) > 0) return std::min(atoi(env), 256);
        cpu_set_t set;
        int cpus = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
        if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(cpus, 16));
    }();
    return n;
}

// ------------------------------------------------------------------------------------------------
// narrowing loops: rows [r0, r1) x columns [c0, c0 + b) of a row-major host matrix -> int8, with the integrality / range check
// ------------------------------------------------------------------------------------------------
struct NarrowFlags { int bad; int gmax; };

#if defined(__x86_64__)
#define CRM_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define CRM_CLONES
#endif

template <class T>
static inline void narrow_row_float(const T* s, int8_t* d, long long b, int& bad, int& gmax) {
    int lb = 0, lm = 0;
    for (long long j = 0; j < b; j++) {
        const T v = s[j];
        const bool in = (v >= (T)-127) & (v <= (T)127);      // false for NaN
        const T vs = in ? v : (T)0;
        const int q = (int)vs;
        lb |= (!in) | ((T)q != vs);
        const int a = q < 0 ? -q : q;
        lm = a > lm ? a : lm;
        d[j] = (int8_t)q;
    }
    bad |= lb; gmax = lm > gmax ? lm : gmax;
}
template <class T>
static inline void narrow_row_int(const T* s, int8_t* d, long long b, int& bad, int& gmax) {
    int lb = 0, lm = 0;
    for (long long j = 0; j < b; j++) {
        const long long v = (long long)s[j];
        const bool in = (v >= -127) & (v <= 127);
        const int q = in ? (int)v : 0;
        lb |= !in;
        const int a = q < 0 ? -q : q;
        lm = a > lm ? a : lm;
        d[j] = (int8_t)q;
    }
    bad |= lb; gmax = lm > gmax ? lm : gmax;
}

#if defined(__x86_64__)
// float64 -> int8, 16 entries per iteration: truncating conversion, exactness by converting back (NaN and out-of-range entries convert to
// INT_MIN and fail the comparison), range by the running maximum of |q|.  The auto-vectoriser leaves the scalar loop alone; this is twice
// as fast per thread and the conversion of 8 GB of dosages is what the end-to-end scan waits for.
__attribute__((target("avx2"))) static void narrow_row_f64_avx2(const double* s, int8_t* d, long long b, int& bad, int& gmax) {
    __m256d ok = _mm256_castsi256_pd(_mm256_set1_epi32(-1));
    __m128i mx = _mm_setzero_si128();
    long long j = 0;
    for (; j + 16 <= b; j += 16) {
        const __m256d v0 = _mm256_loadu_pd(s + j), v1 = _mm256_loadu_pd(s + j + 4), v2 = _mm256_loadu_pd(s + j + 8), v3 = _mm256_loadu_pd(s + j + 12);
        const __m128i q0 = _mm256_cvttpd_epi32(v0), q1 = _mm256_cvttpd_epi32(v1), q2 = _mm256_cvttpd_epi32(v2), q3 = _mm256_cvttpd_epi32(v3);
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(q0), v0, _CMP_EQ_OQ), _mm256_cmp_pd(_mm256_cvtepi32_pd(q1), v1, _CMP_EQ_OQ)),
                                              _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(q2), v2, _CMP_EQ_OQ), _mm256_cmp_pd(_mm256_cvtepi32_pd(q3), v3, _CMP_EQ_OQ))));
        mx = _mm_max_epu32(mx, _mm_max_epu32(_mm_max_epu32(_mm_abs_epi32(q0), _mm_abs_epi32(q1)), _mm_max_epu32(_mm_abs_epi32(q2), _mm_abs_epi32(q3))));
        const __m128i p01 = _mm_packs_epi32(q0, q1), p23 = _mm_packs_epi32(q2, q3);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + j), _mm_packs_epi16(p01, p23));
    }
    int lb = _mm256_movemask_pd(ok) != 0xF;
    alignas(16) unsigned m4[4];
    _mm_store_si128(reinterpret_cast<__m128i*>(m4), mx);
    unsigned lm = 0;
    for (int u = 0; u < 4; u++) lm = m4[u] > lm ? m4[u] : lm;
    int tb = 0, tm = 0;
    narrow_row_float<double>(s + j, d + j, b - j, tb, tm);
    lb |= tb;
    lm = (unsigned)tm > lm ? (unsigned)tm : lm;
    if (lm > 127u) { lb = 1; lm = 127u; }
    bad |= lb; gmax = (int)lm > gmax ? (int)lm : gmax;
}
#endif

static NarrowFlags narrow_f64(const double* src, long long ld, long long r0, long long r1, long long c0, long long b, int8_t* dst, long long ldd) {
    NarrowFlags f{0, 0};
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !(getenv("CRM_NARROW_SCALAR") && atoi(getenv("CRM_NARROW_SCALAR")) != 0);
    if (avx2) {
        for (long long i = r0; i < r1; i++) narrow_row_f64_avx2(src + i * ld + c0, dst + i * ldd, b, f.bad, f.gmax);
        return f;
    }
#endif
    for (long long i = r0; i < r1; i++) narrow_row_float<double>(src + i * ld + c0, dst + i * ldd, b, f.bad, f.gmax);
    return f;
}
CRM_CLONES static NarrowFlags narrow_f32(const float* src, long long ld, long long r0, long long r1, long long c0, long long b, int8_t* dst, long long ldd) {
    NarrowFlags f{0, 0};
    for (long long i = r0; i < r1; i++) narrow_row_float<float>(src + i * ld + c0, dst + i * ldd, b, f.bad, f.gmax);
    return f;
}
template <class T>
static NarrowFlags narrow_int(const T* src, long long ld, long long r0, long long r1, long long c0, long long b, int8_t* dst, long long ldd) {
    NarrowFlags f{0, 0};
    for (long long i = r0; i < r1; i++) narrow_row_int<T>(src + i * ld + c0, dst + i * ldd, b, f.bad, f.gmax);
    return f;
}

static NarrowFlags narrow_any(const void* src, int dtype, long long ld, long long r0, long long r1, long long c0, long long b, int8_t* dst, long long ldd) {
    switch (dtype) {
        case HD_F64: return narrow_f64((const double*)src, ld, r0, r1, c0, b, dst, ldd);
        case HD_F32: return narrow_f32((const float*)src, ld, r0, r1, c0, b, dst, ldd);
        case HD_I8: return narrow_int<int8_t>((const int8_t*)src, ld, r0, r1, c0, b, dst, ldd);       // copy + range check (-128)
        case HD_U8: return narrow_int<uint8_t>((const uint8_t*)src, ld, r0, r1, c0, b, dst, ldd);
        case HD_I16: return narrow_int<int16_t>((const int16_t*)src, ld, r0, r1, c0, b, dst, ldd);
        case HD_I32: return narrow_int<int32_t>((const int32_t*)src, ld, r0, r1, c0, b, dst, ldd);
        case HD_I64: return narrow_int<int64_t>((const int64_t*)src, ld, r0, r1, c0, b, dst, ldd);
        default: return NarrowFlags{1, 0};
    }
}

template <class T>
static void widen_rows(const T* src, long long ld, long long r0, long long r1, long long c0, long long cols, double* dst, long long ldd) {
    for (long long i = r0; i < r1; i++) {
        const T* s = src + i * ld + c0;
        double* d = dst + i * ldd;
        for (long long j = 0; j < cols; j++) d[j] = (double)s[j];
    }
}

// ------------------------------------------------------------------------------------------------
// worker pool
// ------------------------------------------------------------------------------------------------
struct PoolJob {
    virtual ~PoolJob() {}
    virtual bool step() = 0;          // claim and run one unit; false when nothing is left to claim
};

struct Pool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::shared_ptr<PoolJob>> jobs;
    std::vector<std::thread> threads;
    bool started = false;

    void ensure_started() {
        if (started) return;
        started = true;
        const int n = host_threads();
        for (int t = 0; t < n; t++) {
            threads.emplace_back([this] { this->worker(); });
            threads.back().detach();      // workers live for the process; they only touch jobs through shared_ptr
        }
    }
    void submit(const std::shared_ptr<PoolJob>& job) {
        std::lock_guard<std::mutex> lock(mu);
        ensure_started();
        jobs.push_back(job);
        cv.notify_all();
    }
    void worker() {
        for (;;) {
            std::shared_ptr<PoolJob> job;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [this] { return !jobs.empty(); });
                job = jobs.front();
            }
            while (job->step()) {}
            std::lock_guard<std::mutex> lock(mu);
            if (!jobs.empty() && jobs.front() == job) jobs.pop_front();      // exhausted: the next job becomes visible
        }
    }
};
static Pool& pool() { static Pool* p = new Pool(); return *p; }

// ---- asynchronous narrowing job ----
struct FeedPoolJob : PoolJob {
    std::shared_ptr<FeedJob> job;
    explicit FeedPoolJob(const std::shared_ptr<FeedJob>& j) : job(j) {}
    bool step() override {
        FeedJob& J = *job;
        J.workers_inside.fetch_add(1);       // before the cancelled check: feeder_cancel either sees this worker or this worker sees the flag
        const long long u = J.cancelled.load() ? J.total_units : J.next_unit.fetch_add(1);
        if (u >= J.total_units) {
            J.workers_inside.fetch_sub(1);
            std::lock_guard<std::mutex> lock(J.mu);
            J.cv.notify_all();
            return false;
        }
        const long long b = u / J.units_per_block, rc = u % J.units_per_block;
        if (b >= J.nslots) {          // the ring slot must have been copied out by the consumer
            std::unique_lock<std::mutex> lock(J.mu);
            J.cv.wait(lock, [&] { return J.released[b - J.nslots].load() != 0 || J.cancelled.load(); });
        }
        if (!J.cancelled.load()) {
            const long long r0 = rc * J.row_chunk, r1 = std::min(J.rows, r0 + J.row_chunk);
            const long long c0 = J.starts[b], w = J.starts[b + 1] - c0;
            const NarrowFlags f = narrow_any(J.src, J.dtype, J.ld, r0, r1, c0, w, J.slots[b % J.nslots], J.block_ld(b));
            if (f.bad) J.bad[b].store(1);
            int cur = J.gmax[b].load();
            while (f.gmax > cur && !J.gmax[b].compare_exchange_weak(cur, f.gmax)) {}
        }
        if (J.remaining[b].fetch_sub(1) == 1) { std::lock_guard<std::mutex> lock(J.mu); J.cv.notify_all(); }
        J.workers_inside.fetch_sub(1);
        { std::lock_guard<std::mutex> lock(J.mu); J.cv.notify_all(); }
        return true;
    }
};

void feeder_submit(const std::shared_ptr<FeedJob>& job) {
    FeedJob& J = *job;
    const long long nb = J.nblocks();
    const int threads = host_threads();
    // units of ~rows / (2 threads) rows: every block keeps all workers busy, and a unit streams whole rows of the block
    J.row_chunk = std::max<long long>(64, (J.rows + 2 * threads - 1) / (2 * threads));
    J.units_per_block = (J.rows + J.row_chunk - 1) / J.row_chunk;
    J.total_units = nb * J.units_per_block;
    J.remaining.reset(new std::atomic<int>[nb]); J.bad.reset(new std::atomic<int>[nb]); J.gmax.reset(new std::atomic<int>[nb]); J.released.reset(new std::atomic<int>[nb]);
    for (long long b = 0; b < nb; b++) { J.remaining[b].store((int)J.units_per_block); J.bad[b].store(0); J.gmax[b].store(0); J.released[b].store(0); }
    J.next_unit.store(0);
    pool().submit(std::make_shared<FeedPoolJob>(job));
}

void feeder_wait_block(FeedJob& J, long long b, int* bad, int* gmax) {
    std::unique_lock<std::mutex> lock(J.mu);
    J.cv.wait(lock, [&] { return J.remaining[b].load() <= 0; });
    if (bad) *bad = J.bad[b].load();
    if (gmax) *gmax = J.gmax[b].load();
}

void feeder_release_now(FeedJob& J, long long b) {
    J.released[b].store(1);
    std::lock_guard<std::mutex> lock(J.mu);
    J.cv.notify_all();
}

struct ReleaseNote { std::shared_ptr<FeedJob> job; long long b; };
static void CUDART_CB release_callback(void* p) {
    ReleaseNote* note = static_cast<ReleaseNote*>(p);
    feeder_release_now(*note->job, note->b);
    delete note;
}
cudaError_t feeder_release_after(const std::shared_ptr<FeedJob>& job, long long b, cudaStream_t copy_stream) {
    ReleaseNote* note = new ReleaseNote{job, b};
    const cudaError_t e = cudaLaunchHostFunc(copy_stream, release_callback, note);
    if (e != cudaSuccess) { delete note; feeder_release_now(*job, b); }
    return e;
}

void feeder_cancel(FeedJob& J) {
    J.cancelled.store(true);
    std::unique_lock<std::mutex> lock(J.mu);
    J.cv.notify_all();
    J.cv.wait(lock, [&] { return J.workers_inside.load() == 0; });
}

// ---- synchronous parallel-for ----
struct ForJob : PoolJob {
    std::function<void(long long)> fn;
    long long count = 0;
    std::atomic<long long> next{0}, done{0};
    std::mutex mu; std::condition_variable cv;
    bool step() override {
        const long long i = next.fetch_add(1);
        if (i >= count) return false;
        fn(i);
        if (done.fetch_add(1) + 1 == count) { std::lock_guard<std::mutex> lock(mu); cv.notify_all(); }
        return true;
    }
};
static void parallel_for(long long count, const std::function<void(long long)>& fn) {
    if (count <= 0) return;
    auto job = std::make_shared<ForJob>();
    job->fn = fn; job->count = count;
    pool().submit(job);
    while (job->step()) {}             // the caller works too
    std::unique_lock<std::mutex> lock(job->mu);
    job->cv.wait(lock, [&] { return job->done.load() >= job->count; });
}

void host_parallel_widen(const void* src, int dtype, long long ld, long long rows, long long c0, long long cols, double* dst, long long ldd) {
    const int threads = host_threads();
    const long long chunk = std::max<long long>(64, (rows + 4 * threads - 1) / (4 * threads));
    const long long units = (rows + chunk - 1) / chunk;
    parallel_for(units, [&](long long u) {
        const long long r0 = u * chunk, r1 = std::min(rows, r0 + chunk);
        switch (dtype) {
            case HD_F64: for (long long i = r0; i < r1; i++) memcpy(dst + i * ldd, (const double*)src + i * ld + c0, (size_t)cols * 8); break;
            case HD_F32: widen_rows((const float*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            case HD_I8: widen_rows((const int8_t*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            case HD_U8: widen_rows((const uint8_t*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            case HD_I16: widen_rows((const int16_t*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            case HD_I32: widen_rows((const int32_t*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            case HD_I64: widen_rows((const int64_t*)src, ld, r0, r1, c0, cols, dst, ldd); break;
            default: break;
        }
    });
}

void host_parallel_narrow(const void* src, int dtype, long long ld, long long rows, long long c0, long long cols, int8_t* dst, long long ldd, int* bad, int* gmax) {
    const int threads = host_threads();
    const long long chunk = std::max<long long>(64, (rows + 4 * threads - 1) / (4 * threads));
    const long long units = (rows + chunk - 1) / chunk;
    std::atomic<int> any_bad{0}, top{0};
    parallel_for(units, [&](long long u) {
        const NarrowFlags f = narrow_any(src, dtype, ld, u * chunk, std::min(rows, (u + 1) * chunk), c0, cols, dst, ldd);
        if (f.bad) any_bad.store(1);
        int cur = top.load();
        while (f.gmax > cur && !top.compare_exchange_weak(cur, f.gmax)) {}
    });
    if (bad) *bad = any_bad.load();
    if (gmax) *gmax = top.load();
}

// ------------------------------------------------------------------------------------------------
// pinned slot cache
// ------------------------------------------------------------------------------------------------
static std::mutex g_slot_mu;
static int8_t* g_slot_ptr[8] = {};
static size_t g_slot_cap[8] = {};

int8_t* pinned_slot(int index, size_t bytes) {
    if (index < 0 || index >= 8) return nullptr;
    std::lock_guard<std::mutex> lock(g_slot_mu);
    if (g_slot_cap[index] >= bytes) return g_slot_ptr[index];
    if (g_slot_ptr[index]) { cudaFreeHost(g_slot_ptr[index]); g_slot_ptr[index] = nullptr; g_slot_cap[index] = 0; }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_slot_ptr[index] = static_cast<int8_t*>(p);
    g_slot_cap[index] = bytes;
    return g_slot_ptr[index];
}

void release_pinned_slots() {
    std::lock_guard<std::mutex> lock(g_slot_mu);
    for (int i = 0; i < 8; i++) {
        if (g_slot_ptr[i]) cudaFreeHost(g_slot_ptr[i]);
        g_slot_ptr[i] = nullptr; g_slot_cap[i] = 0;
    }
}

}  // namespace crm
