// libcrm_b200.so -- C ABI (include/crm_b200.h) over the sm_100a kernels of this directory.
// Host orchestration only: set-up of the per-gene state, batching of SNPs, kernel launches.
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/crm_b200.h"
#include "common.cuh"
#include "launch.cuh"
#include "feeder.hpp"
#include "trace.cuh"

namespace crm {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define CRM_SOLVER(expr)                                                                     \
    do {                                                                                     \
        cusolverStatus_t _s = (expr);                                                        \
        if (_s != CUSOLVER_STATUS_SUCCESS) {                                                 \
            crm::set_error("%s failed with cusolverStatus %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
            return crm::CRM_ERR_SOLVER;                                                      \
        }                                                                                    \
    } while (0)

static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// Grow-only device buffer.  Allocation is stream-ordered on the stream of the ABI call in flight (AllocScope, set by every entry
// point that takes a stream): memory is obtained and returned in the same order as the kernels that use it, so a buffer that grows
// while earlier launches on that stream still read the old one is released only after them.  (Calls on one handle must be ordered by
// the caller anyway -- a scan reads what the set-up wrote.)  The memory comes from a private per-device pool with an unlimited release
// threshold, so that creating and destroying model objects in a loop (one per gene) reuses the same physical memory instead of paying
// cudaMalloc/cudaFree of tens of GB every time; the device's default pool is left untouched.  crm_trim_pool() hands cached memory back.
static thread_local cudaStream_t g_alloc_stream = (cudaStream_t)0;
struct AllocScope {
    cudaStream_t saved;
    explicit AllocScope(cudaStream_t st) : saved(g_alloc_stream) { g_alloc_stream = st; }
    ~AllocScope() { g_alloc_stream = saved; }
};
static cudaMemPool_t g_pools[32] = {};
static std::mutex g_pool_mu;
static cudaMemPool_t device_pool(int device) {
    if (device < 0 || device >= 32) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pools[device]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        unsigned long long threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        g_pools[device] = pool;
    }
    return g_pools[device];
}
// bytes held by the pool beyond what is in use (reusable by the next reserve)
static size_t pool_cached_bytes(int device) {
    cudaMemPool_t pool = device_pool(device);
    unsigned long long reserved = 0, used = 0;
    if (pool && cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used) return (size_t)(reserved - used);
    return 0;
}
// stream-ordered scratch of the kernel launchers (split-K partial sums ...): from the same private pool -- the device's default pool
// releases its memory at every synchronisation (release threshold 0), which turns each such allocation into a trip to the OS
cudaError_t pool_alloc_async(void** ptr, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool = device_pool(dev);
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, st) : cudaMallocAsync(ptr, bytes, st);
}
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return CRM_OK;
        release();
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool = device_pool(dev);
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&ptr, bytes, pool, g_alloc_stream) : cudaMallocAsync(&ptr, bytes, g_alloc_stream);
        if (e != cudaSuccess) { set_error("stream-ordered allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); ptr = nullptr; cudaGetLastError(); return CRM_ERR_CUDA; }
        cap = bytes;
        return CRM_OK;
    }
    void release() { if (ptr) cudaFreeAsync(ptr, g_alloc_stream); ptr = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// Creating a cuSOLVER handle costs tens of milliseconds, so one context per device lives in a process-wide pool and is
// reused by every model object (set-up calls are serialised by the pool mutex).
struct EigCtx { cusolverDnHandle_t solver = nullptr; cublasHandle_t blas = nullptr; DevBuf mat, val, work, vec, ws, quality, blas_ws; std::vector<char> host_work; };
// cuBLAS handle of the set-up with an explicit workspace from the private pool (no allocation inside the library calls)
static int ensure_blas(EigCtx& e) {
    if (e.blas) return CRM_OK;
    if (cublasCreate(&e.blas) != CUBLAS_STATUS_SUCCESS) { e.blas = nullptr; set_error("cublasCreate failed"); return CRM_ERR_SOLVER; }
    const size_t bytes = (size_t)64 << 20;
    if (e.blas_ws.reserve(bytes) == CRM_OK) cublasSetWorkspace(e.blas, e.blas_ws.ptr, bytes);
    return CRM_OK;
}
struct EigPool { std::mutex mu; std::vector<EigCtx> ctx; };
static EigPool g_eig_pool[16];

struct Handle {
    int device = 0;
    bool ready = false;
    long long n = 0;
    int c = 0, k0 = 0, k1 = 0, R = 0;
    long long mL = 0;
    int m = 0, mp = 0, Mx = 0, ldH = 0, kexp = 0, epitch = 0, M2 = 0, ld2 = 0, max_rank = 0;
    std::vector<double> rho;
    // per-gene state
    // Operands that are contracted against genotype rows.  `cells`: one row per cell (the reference's expanded G).
    // `donors`: the same operands summed over the cells of each donor, for donor-level genotypes (G_cells = G_donors[donor]).
    struct GenoSpace { long long K = 0; const double* HxE = nullptr; long long ldE = 0; const double* Hx = nullptr; long long ldHx = 0;
                       const double* A2 = nullptr; int ld2 = 0; };
    GenoSpace cells, donors;
    const GenoSpace* gs = nullptr;   // space of the scan in flight
    bool donors_set = false;
    DevBuf HxE_D, A2_D, dperm, doff;
    DevBuf HxE;           // optional n x (kexp * ldH) pre-expanded basis [Hx | Hx.E0_1 | ... | Hx.E0_k] (see launch_rotation)
    bool use_hxe = false, hxe_built = false, hxe_decided = false;
    // exact int8 split of the rotation for integer dosages (ozaki.cuh): digit planes of HxE, built by the first rotation that uses them
    int rotation_mode = 0;        // 0 auto (int8 split when the genotype block is integer), 1 fp64 DMMA only, 2 int8 split required
    bool oz_built = false, oz_built_a2 = false;
    DevBuf A8, a8expo, Gt8, G2t8, D32, ozflags, A28, a28expo;   // A28: digit planes of A2 = [1 | E0 | pairs] for the g^2 Grams
    bool oz_block_valid = false;   // Gt8 / G2t8 hold the int8 image of the genotype block of the rotation just done
    int oz_block_gmax = 0;
    // affine-integer genotype columns g = a d + b (ozaki.cuh): Gt8 holds d; aff = [a | b | tolerance] per column of the block;
    // colsum / colsum2 = column sums of [Hx | Hx.E0_j] and of A2 (per gene, built on demand)
    bool oz_block_affine = false, colsum_valid = false;
    DevBuf aff, affscratch, colsum, colsum2, sq1;
    double prof_oz_gemm_ops = 0.0; std::vector<cudaEvent_t> prof_oz_events;
    // Khatri-Rao structure of the background basis (crm_set_background_factors): L[:, i q + c] = (E0 M)[:, i] * hK[:, c], so that the
    // products L.E0_j are combinations of the symmetric triple products hK_c.E0_l.E0_j -- the rotation contracts the compact set of
    // distinct columns (kr_rows of them, layout in kr_layout) and kr_expand_kernel rebuilds the full C from it
    bool kr = false, oz_built_kr = false, hxe_built_kr = false;
    int kr_q = 0, kr_r = 0;
    long long kr_R1 = 0, kr_R2 = 0, kr_R3 = 0, kr_rows = 0;
    DevBuf krK, krM, Cc, fittab;
    // declaration made before crm_setup: applied (and verified, result read at the end of the set-up) inside it
    bool kr_pending = false, kr_unverified = false;
    const double* kr_pending_hK = nullptr; long long kr_pending_ld = 0; int kr_pending_q = 0, kr_pending_r = 0;
    const double* kr_pending_M = nullptr;
    // digit planes built during the set-up on a side stream (crm_hint_integer_genotypes): the first int8 rotation waits for planes_ev
    bool early_planes = false, planes_ev_pending = false;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ev = nullptr, planes_ev = nullptr;
    int hxe_blocks = 0;   // context blocks j held by HxE at a time: kexp = whole basis resident, fewer = streamed in groups
    DevBuf Hx, Eext, A2, gram, S, yr, Wr, Tt, stats, eigwork, eigmat, eigval, devinfo;
    // null-model state for the association scans
    // scan workspaces
    DevBuf C, sq, Hg, gr, Vg, GEr, fit_lml, fit_delta, fit_scale, fit_beta, fit_x, fit_nfev, fit_flags;
    DevBuf Ys, sgram, HY, Zs, lin, Zp, ucoef, coef;
    DevBuf YW, ywgram;    // [R][1 + c][mp] rotated [y | W] per rho and the (1 + c)^2 plain Gram of [y | W] (wide-design fits)
    DevBuf rho_idx, best_lml, v0, v1, perm, offsets, Q, lam, nlam, sflags, liu, ifault, conv, gchunk[2], gtchunk[2], scratch;
    // optional timing of the rotation kernel (K1, EXPAND mode) with events on the launching stream
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_events;
    double prof_flops = 0.0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t last_stream = nullptr;   // stream of the latest ABI call on this handle: the buffers are released in its order
    // genotypes staged ahead of the scan (crm_stage_genotypes): the whole host matrix in device memory, copied in column chunks on
    // the copy stream while the set-up runs; one event per chunk
    DevBuf gstage;
    const double* stage_src = nullptr; long long stage_ldg = 0, stage_p = 0, stage_rows = 0, stage_ld = 0; int stage_chunk = 0;
    std::vector<cudaEvent_t> stage_events;          // one per chunk
    bool stage_valid = false;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // host genotypes converted to int8 blocks by the feeder (feeder.hpp): job in flight, the matrix it describes, device landing buffers
    std::shared_ptr<FeedJob> feed;
    const void* feed_src = nullptr; long long feed_ld = 0, feed_p = 0, feed_rows = 0; int feed_dtype = 0;
    DevBuf g8dev[2], gwide, gwide2;
    // buffers of a model object recycled from an earlier one (crm_destroy keeps them, see BufCache): work on them is ordered after this event
    cudaEvent_t adopt_ev = nullptr;
    template <class F> void for_each_buf(F&& fn) {
        DevBuf* all[] = {&A8, &a8expo, &Gt8, &G2t8, &D32, &ozflags, &A28, &a28expo, &HxE_D, &A2_D, &dperm, &doff, &HxE, &Hx, &Eext, &A2, &gram, &S, &yr, &Wr, &Tt, &stats, &eigwork, &eigmat, &eigval, &devinfo, &C, &sq, &Hg, &gr,
                         &Vg, &GEr, &fit_lml, &fit_delta, &fit_scale, &fit_beta, &fit_x, &fit_nfev, &fit_flags, &rho_idx, &best_lml,
                         &v0, &v1, &perm, &offsets, &Q, &lam, &nlam, &sflags, &liu, &ifault, &conv, &gchunk[0], &gchunk[1],
                         &gtchunk[0], &gtchunk[1], &gstage, &g8dev[0], &g8dev[1], &gwide, &gwide2, &aff, &affscratch, &colsum, &colsum2, &sq1, &scratch, &Ys, &sgram, &HY, &Zs, &lin, &Zp, &ucoef, &coef, &YW, &ywgram, &krK, &krM, &Cc, &fittab};
        for (DevBuf* b : all) fn(*b);
    }
    void free_all() { for_each_buf([](DevBuf& b) { b.release(); }); }
};

// Device buffers of destroyed model objects, kept for the next crm_create on the same device.  A scan of many genes creates and drops one
// model per gene; handing the (grow-only) buffers over instead of returning them to the memory pool means that a model of the same shape
// allocates nothing at all -- pool allocations of the large operands (17 GB of digit planes) were measured to take 0.1-2.5 s every few
// steps when the pool had split its blocks for smaller requests (profiles/r02_step_trace.txt).  At most two sets per device are kept;
// crm_trim_pool releases them.
struct BufCache { std::vector<DevBuf> bufs; cudaEvent_t ev = nullptr; };
static std::mutex g_cache_mu;
static std::vector<BufCache> g_cache[32];
static const size_t BUF_CACHE_DEPTH = 2;

// first use of recycled buffers on the stream of this call: order it after the last work of their previous owner
static int adopt_buffers(Handle* h, cudaStream_t st) {
    if (!h->adopt_ev) return CRM_OK;
    CRM_CUDA(cudaStreamWaitEvent(st, h->adopt_ev, 0));
    cudaEventDestroy(h->adopt_ev);
    h->adopt_ev = nullptr;
    return CRM_OK;
}

// ------------------------------------------------------------------------------------------------
// small set-up kernels
// ------------------------------------------------------------------------------------------------
__global__ void build_eext_kernel(const double* E0, long long lde0, long long n, int k0, double* Eext, int epitch) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * epitch) return;
    const long long i = idx / epitch; const int j = (int)(idx - i * epitch);
    Eext[idx] = (j == 0) ? 1.0 : (j <= k0 ? E0[i * lde0 + (j - 1)] : 0.0);
}
// out[i][(j - j0) * ldH + a] = Eext[i][j] * Hx[i][a]  for j in [j0, j0 + nj)   (j = 0: Hx itself)
__global__ void build_hxe_kernel(const double* Hx, int ldH, const double* Eext, int epitch, int j0, int nj, long long n, double* out) {
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {          // one cell per block iteration, threads over basis columns
        const double* hrow = Hx + i * ldH;
        const double* erow = Eext + i * epitch + j0;
        double* orow = out + i * (long long)nj * ldH;
        for (int a = threadIdx.x; a < ldH; a += blockDim.x) {
            const double hv = hrow[a];
            for (int j = 0; j < nj; j++) orow[(long long)j * ldH + a] = erow[j] * hv;
        }
    }
}
// ---- Khatri-Rao structured background (Handle::kr) ----
// Compact row layout of the rotation when L[:, i q + c] = (E0 M)[:, i] * hK[:, c]:
//   [0, ldH)                      Hx[:, a]                               (block j = 0 of the full layout)
//   R1 + (j-1) k1 + a             E1[:, a] * E0[:, j-1]                   j = 1..k0, a < k1
//   R2 + (j-1) (1+c) + t          [y | W][:, t] * E0[:, j-1]              t < 1 + c
//   R3 + pair(l, j) q + cc        hK[:, cc] * E0[:, l] * E0[:, j]         l <= j, pair = j (j + 1) / 2 + l  (the pair columns of A2)
// Full layout (row = j ldH + a): block j >= 1, column k1 + i q + cc  =  sum_l M[l][i] * compact[R3 + pair(l, j-1) q + cc].
// does the declared structure reproduce the L columns of Hx?  flag |= 1 on the first entry that does not
constexpr int KR_CHECK_CELLS = 8;
__global__ void __launch_bounds__(256) kr_check_kernel(const double* Hx, int ldH, int k1, const double* Eext, int epitch, int k0, const double* hK, int q, const double* M,
                                                       int r, long long n, int* flag) {
    extern __shared__ double kc_smem[];               // us[cell][ii], mag[cell][ii] of KR_CHECK_CELLS cells
    double* us = kc_smem; double* mag = kc_smem + KR_CHECK_CELLS * r;
    const long long i0 = (long long)blockIdx.x * KR_CHECK_CELLS;
    for (int t = threadIdx.x; t < KR_CHECK_CELLS * r; t += blockDim.x) {
        const int ci = t / r, ii = t - ci * r; const long long i = i0 + ci;
        double u = 0.0, a = 0.0;
        if (i < n) for (int l = 0; l < k0; l++) { const double v = Eext[i * epitch + 1 + l] * M[l * r + ii]; u += v; a += fabs(v); }
        us[t] = u; mag[t] = a;
    }
    __syncthreads();
    const int per = r * q;
    bool bad = false;
    for (int t = threadIdx.x; t < KR_CHECK_CELLS * per; t += blockDim.x) {
        const int ci = t / per, col = t - ci * per, ii = col / q, cc = col - ii * q; const long long i = i0 + ci;
        if (i >= n) break;
        const double hk = hK[i * q + cc];
        const double want = us[ci * r + ii] * hk, got = Hx[i * ldH + k1 + col];
        if (!(fabs(got - want) <= 1e-11 * mag[ci * r + ii] * fabs(hk) + 1e-290)) bad = true;
    }
    if (bad) atomicOr(flag, 1);
}
// full rotation output C[(s kexp + j)][a] from the compact one S[s][row] (one block per SNP; M in shared memory, k0 x rp)
constexpr int KR_IT = 8;
__global__ void __launch_bounds__(256) kr_expand_kernel(const double* __restrict__ S, long long lds, const double* __restrict__ M, int k0, int r, int q, int k1, int m,
                                                        int c, int ldH, long long R1, long long R2, long long R3, long long B, double* __restrict__ C) {
    extern __shared__ double Msh[];
    const int rp = (r + KR_IT - 1) / KR_IT * KR_IT;
    for (int t = threadIdx.x; t < k0 * rp; t += blockDim.x) { const int l = t / rp, i = t - l * rp; Msh[t] = i < r ? M[l * r + i] : 0.0; }
    __syncthreads();
    const int kexp = 1 + k0, small = k1 + (ldH - m);
    for (long long s = blockIdx.x; s < B; s += gridDim.x) {
        const double* Ss = S + s * lds;
        double* Cs = C + s * (long long)kexp * ldH;
        for (int a = threadIdx.x; a < ldH; a += blockDim.x) Cs[a] = Ss[a];
        for (int t = threadIdx.x; t < k0 * small; t += blockDim.x) {
            const int j = t / small, u = t - j * small;
            double v; int a;
            if (u < k1) { a = u; v = Ss[R1 + (long long)j * k1 + u]; }
            else { const int w = u - k1; a = m + w; v = w < 1 + c ? Ss[R2 + (long long)j * (1 + c) + w] : 0.0; }
            Cs[(long long)(1 + j) * ldH + a] = v;
        }
        for (int t = threadIdx.x; t < k0 * q; t += blockDim.x) {
            const int j = t / q, cc = t - j * q;
            double* out = Cs + (long long)(1 + j) * ldH + k1 + cc;
            for (int i0 = 0; i0 < r; i0 += KR_IT) {
                double acc[KR_IT];
#pragma unroll
                for (int u = 0; u < KR_IT; u++) acc[u] = 0.0;
                for (int l = 0; l < k0; l++) {
                    const int lo = min(l, j), hi = max(l, j);
                    const double sv = Ss[R3 + (long long)(hi * (hi + 1) / 2 + lo) * q + cc];
                    const double* mrow = Msh + l * rp + i0;
#pragma unroll
                    for (int u = 0; u < KR_IT; u++) acc[u] = fma(mrow[u], sv, acc[u]);
                }
#pragma unroll
                for (int u = 0; u < KR_IT; u++) if (i0 + u < r) out[(long long)(i0 + u) * q] = acc[u];
            }
        }
    }
}
// fp64 image of the compact basis (same columns and the same products as the digit planes of the int8 route): out[i][row]
__global__ void build_hxe_compact_kernel(const double* Hx, int ldH, int k1, int m, int c, const double* Eext, int epitch, int k0, const double* A2, int ld2,
                                         const double* hK, int q, long long R1, long long R2, long long R3, long long rows, long long n, double* out) {
    const long long sym_end = R3 + (long long)(k0 * (k0 + 1) / 2) * q;
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        const double* hrow = Hx + i * ldH; const double* erow = Eext + i * epitch + 1; const double* prow = A2 + i * ld2 + 1 + k0; const double* krow = hK + i * q;
        double* orow = out + i * rows;
        for (long long col = threadIdx.x; col < rows; col += blockDim.x) {
            double v = 0.0;
            if (col < R1) v = hrow[col];
            else if (col < R2) { const int t = (int)(col - R1), j = t / k1, a = t - j * k1; v = erow[j] * hrow[a]; }
            else if (col < R3) { const int t = (int)(col - R2), j = t / (1 + c), w = t - j * (1 + c); v = erow[j] * hrow[m + w]; }
            else if (col < sym_end) { const long long t = col - R3; const int pi = (int)(t / q), cc = (int)(t - (long long)pi * q); v = prow[pi] * krow[cc]; }
            orow[col] = v;
        }
    }
}
// donor-level pre-expanded basis: out[dn][j * ldH + a] = sum over the cells t of donor dn of Eext[t][j] * Hx[t][a]
__global__ void aggregate_hxe_kernel(const double* Hx, int ldH, const double* Eext, int epitch, int kexp, const int* perm, const int* off, double* out) {
    const int dn = blockIdx.x, j = blockIdx.y, a = blockIdx.z * blockDim.x + threadIdx.x;
    if (a >= ldH) return;
    double s = 0.0;
    for (int t = off[dn]; t < off[dn + 1]; t++) { const long long i = perm[t]; s += Eext[i * epitch + j] * Hx[i * ldH + a]; }
    out[(long long)dn * kexp * ldH + (long long)j * ldH + a] = s;
}
// out[dn][col] = sum over the cells of donor dn of src[cell][col]
__global__ void aggregate_rows_kernel(const double* src, long long ld, int cols, const int* perm, const int* off, double* out, long long ldo) {
    const int dn = blockIdx.x, col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    double s = 0.0;
    for (int t = off[dn]; t < off[dn + 1]; t++) s += src[(long long)perm[t] * ld + col];
    out[(long long)dn * ldo + col] = s;
}
// A2 = [1 | E0 | E0_j * E0_l (j >= l, packed)]
__global__ void build_a2_kernel(const double* E0, long long lde0, long long n, int k0, double* A2, int ld2) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * ld2) return;
    const long long i = idx / ld2; const int col = (int)(idx - i * ld2);
    const int npair = k0 * (k0 + 1) / 2;
    double v = 0.0;
    if (col == 0) v = 1.0;
    else if (col <= k0) v = E0[i * lde0 + (col - 1)];
    else if (col < 1 + k0 + npair) {
        const int e = col - 1 - k0;
        int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (j * (j + 1) / 2 > e) j--;
        while ((j + 1) * (j + 2) / 2 <= e) j++;
        const int l = e - j * (j + 1) / 2;
        v = E0[i * lde0 + j] * E0[i * lde0 + l];
    }
    A2[idx] = v;
}
// C_rho = D^(1/2) (H'H) D^(1/2), D = diag(rho I_k1, (1-rho) I), restricted to the columns [a0, a0 + ms) that D keeps
// (rho = 1 keeps only the E1 block, rho = 0 only the L block: the other block is exactly zero)
__global__ void scale_gram_kernel(const double* gram, int ldg, int a0, int ms, int k1, double rho, double* out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ms * ms) return;
    const int a = a0 + idx / ms, b = a0 + idx % ms;
    const double da = a < k1 ? sqrt(rho) : sqrt(1.0 - rho), db = b < k1 ? sqrt(rho) : sqrt(1.0 - rho);
    out[idx] = da * db * gram[(long long)a * ldg + b];
}
// eigenpairs of the kept block [a0, a0 + ms) (ascending, eigenvector i in row i of V, length ms) ->
// S[i], Tt[a][rho*mp + i] = d_a V[i][a - a0] / sqrt(S_i); entries i >= ms and rows outside the block are zero
__global__ void build_basis_kernel(const double* V, const double* ev, int m, int mp, int a0, int ms, int k1, double rho, int tall,
                                   double* S, double* Tt, long long ldt, int rho_index, int* rank_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const double evmax = ev[ms - 1];
    const double thr_rel = 1e-12 * evmax;
    if (idx < mp) {
        double s = 0.0;
        if (idx < ms) { const double e = ev[idx]; const bool keep = (e > thr_rel) && (tall || e >= CRM_EPS_SMALL) && e > 0.0; s = keep ? e : 0.0; }
        S[idx] = s;
    }
    if (idx == 0) { int r = 0; for (int i = 0; i < ms; i++) { const double e = ev[i]; if ((e > thr_rel) && (tall || e >= CRM_EPS_SMALL) && e > 0.0) r++; } rank_out[rho_index] = r; }
    for (long long t = idx; t < (long long)m * mp; t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t / mp), i = (int)(t - (long long)a * mp);
        double v = 0.0;
        if (i < ms && a >= a0 && a < a0 + ms) {
            const double e = ev[i];
            const bool keep = (e > thr_rel) && (tall || e >= CRM_EPS_SMALL) && e > 0.0;
            if (keep) { const double da = a < k1 ? sqrt(rho) : sqrt(1.0 - rho); v = da * V[(long long)i * ms + (a - a0)] / sqrt(e); }
        }
        Tt[(long long)a * ldt + (long long)rho_index * mp + i] = v;
    }
}
// yr[rho][i] = sum_a Tt[a][rho*mp+i] gram[a][m],  Wr[rho][cc][i] likewise with gram[a][m+1+cc]
__global__ void rotate_null_kernel(const double* Tt, long long ldt, const double* gram, int ldg, int m, int mp, int R, int c,
                                   double* yr, double* Wr) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)R * mp) return;
    const int rho = (int)(idx / mp), i = (int)(idx - (long long)rho * mp);
    for (int col = 0; col <= c; col++) {
        double s = 0.0;
        for (int a = 0; a < m; a++) s += Tt[(long long)a * ldt + idx] * gram[(long long)a * ldg + m + col];
        if (col == 0) yr[idx] = s; else Wr[((long long)rho * c + (col - 1)) * mp + i] = s;
    }
}
// refresh the phenotype-dependent column a = m of the (pre-expanded) bases after Hx[:, m] changed
__global__ void refresh_y_hxe_kernel(const double* Hx, int ldH, int m, const double* Eext, int epitch, int kexp, long long n, double* HxE) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * kexp) return;
    const long long i = idx / kexp; const int j = (int)(idx - i * kexp);
    HxE[i * (long long)kexp * ldH + (long long)j * ldH + m] = Eext[i * epitch + j] * Hx[i * ldH + m];
}
__global__ void refresh_y_donor_kernel(const double* Hx, int ldH, int m, const double* Eext, int epitch, int kexp, const int* perm, const int* off, long long d,
                                       double* HxE_D) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d * kexp) return;
    const long long dn = idx / kexp; const int j = (int)(idx - dn * kexp);
    double s = 0.0;
    for (int t = off[dn]; t < off[dn + 1]; t++) { const long long i = perm[t]; s += Eext[i * epitch + j] * Hx[i * ldH + m]; }
    HxE_D[dn * (long long)kexp * ldH + (long long)j * ldH + m] = s;
}
__global__ void symmetrise_row_kernel(double* gram, int ldg, int row, int count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < count) gram[(long long)a * ldg + row] = gram[(long long)row * ldg + a];
}
// YW[rho][0][i] = yr[rho][i], YW[rho][1 + a][i] = Wr[rho][a][i];  ywgram = plain Gram of [y | W] from stats
__global__ void build_yw_kernel(const double* yr, const double* Wr, const double* stats, int R, int c, int mp, double* YW, double* ywgram) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)R * (1 + c) * mp;
    if (idx < total) {
        const int i = (int)(idx % mp); const long long rc = idx / mp; const int col = (int)(rc % (1 + c)), rho = (int)(rc / (1 + c));
        YW[idx] = col == 0 ? yr[(long long)rho * mp + i] : Wr[((long long)rho * c + (col - 1)) * mp + i];
    }
    if (idx < (long long)(1 + c) * (1 + c)) {
        const int a = (int)(idx / (1 + c)), b = (int)(idx % (1 + c));
        double v;
        if (a == 0 && b == 0) v = stats[0];
        else if (a == 0) v = stats[b];
        else if (b == 0) v = stats[a];
        else v = stats[1 + c + (a - 1) * c + (b - 1)];
        ywgram[idx] = v;
    }
}
__global__ void extract_stats_kernel(const double* gram, int ldg, int m, int c, double* stats) {
    if (threadIdx.x == 0) stats[0] = gram[(long long)m * ldg + m];
    for (int t = threadIdx.x; t < c; t += blockDim.x) stats[1 + t] = gram[(long long)m * ldg + m + 1 + t];
    for (int t = threadIdx.x; t < c * c; t += blockDim.x) { const int a = t / c, b = t - a * c; stats[1 + c + t] = gram[(long long)(m + 1 + a) * ldg + m + 1 + b]; }
}
// a handful of host doubles -> device memory through the kernel-argument path: a cudaMemcpyAsync from host memory would queue behind
// the genotype transfer on the host-to-device copy engine (crm_stage_genotypes) and stall the compute stream until it ends
struct SmallValues { double v[64]; };
__global__ void set_values_kernel(double* dst, SmallValues vals, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = vals.v[i];
}
static int upload_small(double* dst, const double* src, int count, cudaStream_t st) {
    if (count > 64) { set_error("upload_small: %d values", count); return CRM_ERR_UNSUPPORTED; }
    SmallValues sv{};
    for (int i = 0; i < count; i++) sv.v[i] = src[i];
    set_values_kernel<<<1, 64, 0, st>>>(dst, sv, count);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
__global__ void finalize_interaction_kernel(const int* rho_idx, const double* v0, const double* v1, const double* grid, long long p,
                                            double* rho1, double* e2, double* g2, double* eps2) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p) return;
    const double r = grid[rho_idx[s]];
    rho1[s] = r; e2[s] = v0[s] * r; g2[s] = v0[s] * (1.0 - r); eps2[s] = v1[s];
}
__global__ void override_select_kernel(const int* ov_idx, const double* ov_v0, const double* ov_v1, long long p, int* rho_idx, double* v0, double* v1) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p) return;
    if (ov_idx) rho_idx[s] = ov_idx[s];
    if (ov_v0) v0[s] = ov_v0[s];
    if (ov_v1) v1[s] = ov_v1[s];
}
__global__ void or_flags_kernel(const int* fit_flags, const int* rho_idx, const int* sflags, int R, long long p, int* out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p) return;
    const int f = fit_flags[s * R + rho_idx[s]];
    out[s] = (sflags[s] & 1) | ((f & 1) << 1) | ((f & 2) << 1);
}

// Ys = [y | W | E0] (n x ld) from Hx (columns m.. ) and Eext (columns 1..k0)
__global__ void build_ys_kernel(const double* Hx, int ldH, int m, int c, const double* Eext, int epitch, int k0, long long n, double* Ys, int ld) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * ld) return;
    const long long i = idx / ld; const int j = (int)(idx - i * ld);
    double v = 0.0;
    if (j <= c) v = Hx[i * ldH + m + j];
    else if (j < 1 + c + k0) v = Eext[i * epitch + 1 + (j - 1 - c)];
    Ys[idx] = v;
}
// out[col][i] = sum_a Tt[a][off + i] HY[col][a]
__global__ void apply_basis_kernel(const double* Tt, long long ldt, long long off, const double* HY, long long ldhy, int m, int mp, int ncol, double* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncol * mp) return;
    const int col = (int)(idx / mp), i = (int)(idx - (long long)col * mp);
    double s = 0.0;
    for (int a = 0; a < m; a++) s += Tt[(long long)a * ldt + off + i] * HY[(long long)col * ldhy + a];
    out[idx] = s;
}
// best rho per SNP -> persistent effect and BLUP coefficients: beta_g = beta[c], coef = v0 rho ucoef / sqrt(2 maf (1 - maf))
__global__ void finalize_betas_kernel(const int* rho_idx, const double* v0, const double* grid, const double* beta, const double* ucoef,
                                      const double* maf, int R, int P, int c, int k0, long long p, double* beta_g, double* coef) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p) return;
    const int r = rho_idx[s];
    const long long o = s * R + r;
    beta_g[s] = beta[o * P + c];
    const double f = maf[s], norm = 1.0 / sqrt(2.0 * f * (1.0 - f)), w = v0[s] * grid[r] * norm;
    for (int j = 0; j < k0; j++) coef[s * k0 + j] = w * ucoef[o * k0 + j];
}

static inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------
// set-up
// ------------------------------------------------------------------------------------------------
// donor-level operands from the current Hx / Eext / A2 and the stored cell -> donor grouping
static int aggregate_donors(Handle* h, cudaStream_t st) {
    const long long d = h->donors.K, ldE = (long long)h->kexp * h->ldH;
    CRM_CHECK(h->HxE_D.reserve((size_t)d * ldE * 8));
    CRM_CHECK(h->A2_D.reserve((size_t)d * h->ld2 * 8));
    dim3 g1((unsigned)d, (unsigned)h->kexp, (unsigned)((h->ldH + 127) / 128));
    aggregate_hxe_kernel<<<g1, 128, 0, st>>>(h->Hx.as<double>(), h->ldH, h->Eext.as<double>(), h->epitch, h->kexp, h->dperm.as<int>(), h->doff.as<int>(), h->HxE_D.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    dim3 g2((unsigned)d, (unsigned)((h->ld2 + 127) / 128), 1);
    aggregate_rows_kernel<<<g2, 128, 0, st>>>(h->A2.as<double>(), h->ld2, h->ld2, h->dperm.as<int>(), h->doff.as<int>(), h->A2_D.as<double>(), h->ld2);
    CRM_CUDA(cudaGetLastError()); count_launch();
    h->donors.HxE = h->HxE_D.as<double>(); h->donors.ldE = ldE; h->donors.Hx = h->HxE_D.as<double>(); h->donors.ldHx = ldE;
    h->donors.A2 = h->A2_D.as<double>(); h->donors.ld2 = h->ld2;
    return CRM_OK;
}

static int build_test_contexts(Handle* h, const double* E0, long long lde0, cudaStream_t st) {
    build_eext_kernel<<<blocks_for(h->n * h->epitch, 256), 256, 0, st>>>(E0, lde0, h->n, h->k0, h->Eext.as<double>(), h->epitch);
    CRM_CUDA(cudaGetLastError()); count_launch();
    build_a2_kernel<<<blocks_for(h->n * h->ld2, 256), 256, 0, st>>>(E0, lde0, h->n, h->k0, h->A2.as<double>(), h->ld2);
    CRM_CUDA(cudaGetLastError()); count_launch();
    h->hxe_built = false;    // the pre-expanded basis is (re)built by the first cell-level rotation that needs it
    h->kr = false;           // a declared structure of the background refers to the contexts it was declared with
    h->oz_built = false;
    h->oz_built_a2 = false;
    h->colsum_valid = false;
    h->cells.K = h->n; h->cells.HxE = nullptr; h->cells.ldE = (long long)h->kexp * h->ldH;   // cells.HxE unused: see launch_rotation
    h->cells.Hx = h->Hx.as<double>(); h->cells.ldHx = h->ldH; h->cells.A2 = h->A2.as<double>(); h->cells.ld2 = h->ld2;
    if (h->donors_set) CRM_CHECK(aggregate_donors(h, st));
    return CRM_OK;
}

// Rotation of [g, g.E0_1, ..., g.E0_k] onto [H | y | W] for B SNP columns: C[(s*kexp + j)][a].
// Two equivalent routes through K1: with the pre-expanded basis HxE = [Hx | Hx.E0_1 | ...] the Hadamard factor sits on the
// basis side and the loop is a plain DMMA contraction (98% of the FP64 tensor peak).  HxE stays resident when its
// n*kexp*ldH*8 bytes fit comfortably; otherwise it is streamed through a smaller buffer in groups of context blocks, rebuilt
// per rotation call (one elementwise pass per group, negligible against the contraction).  Without it (CRM_NO_HXE=1) the
// factor is applied to the genotype fragments on the fly (EXPAND mode, ~90% of peak, no extra memory).
// Rotation through the exact int8 split (ozaki.cuh): *used = 1 when the block was integer-valued and the route was taken.
// CRM_INT8_GEMM=lt routes the int8 contraction through cuBLASLt + oz_combine_kernel instead of the fused tcgen05 kernel
// A genotype matrix as the ABI receives it: float64 or int8 on the device, any HostDtype on the host.
struct GSource { const void* ptr; long long ld; int dtype; int on_host; };
// A block of SNP columns as the kernels see it.  Either image may be missing: `G` (float64, device pointer, leading dimension, number
// of addressable columns) is produced on demand from `G8` (int8 dosages, row-major, leading dimension in bytes) by block_f64();
// gmax = largest |dosage| of the int8 image when the host already knows it (-1: unknown).
struct GBlock { const double* G; long long ld; long long cols; const double* G2; long long ld2; long long b; long long s0;
                const int8_t* G8; long long ld8; int gmax; };

// float64 image of a block that arrived as int8 (consumers without an int8 route: DMMA contractions, permuted designs)
static int block_f64(Handle* h, GBlock& k, cudaStream_t st) {
    if (k.G) return CRM_OK;
    if (!k.G8) { set_error("genotype block without data"); return CRM_ERR_INVALID; }
    const long long rows = h->gs->K, ld = round_up(k.b, 2);
    CRM_CHECK(h->gwide.reserve((size_t)rows * ld * 8));
    CRM_CHECK(oz_launch_widen_i8(k.G8, k.ld8, rows, k.b, h->gwide.as<double>(), ld, st));
    k.G = h->gwide.as<double>(); k.ld = ld; k.cols = k.b;
    return CRM_OK;
}

static bool int8_route_library() {
    static const bool lt = [] { const char* v = getenv("CRM_INT8_GEMM"); return v && !strcmp(v, "lt"); }();
    return lt;
}
// C[s][col] (ldc) = recombined int8 contraction of the digit planes P8 [8][Mp][Kp] against Gt8 [Bp][Kp]
static int int8_split_contract(Handle* h, const int8_t* P8, long long Mp, long long Mtot, const int* expo, const int8_t* Gt8, long long Bp, long long B,
                               long long Kp, double* C, long long ldc, cudaStream_t st) {
    if (!int8_route_library()) return oz_launch_mma(P8, Mp, Mtot, expo, Gt8, Bp, B, Kp, C, ldc, st);
    CRM_CHECK(h->D32.reserve((size_t)OZAKI_SLICES * Mp * Bp * sizeof(int)));
    CRM_CHECK(oz_int8_gemm(P8, (long long)OZAKI_SLICES * Mp, Gt8, Bp, Kp, h->D32.as<int>(), Bp, st));
    return oz_launch_combine(h->D32.as<int>(), Mp, Bp, expo, Mtot, B, C, ldc, st);
}

// column sums of the expanded basis, colsum[j * ldH + a] = sum_i Eext[i][j] Hx[i][a], and of A2: the images of a constant genotype
// column, which map contractions of the integer part d of an affine column g = a d + b back to contractions of g
static int ensure_column_sums(Handle* h, cudaStream_t st) {
    if (h->colsum_valid) return CRM_OK;
    CRM_CHECK(h->colsum.reserve((size_t)h->kexp * h->ldH * 8));
    CRM_CHECK(h->colsum2.reserve((size_t)h->ld2 * 8));
    CRM_CUDA(cudaMemsetAsync(h->colsum.ptr, 0, (size_t)h->kexp * h->ldH * 8, st));
    GemmOperands op{};
    op.A = h->Hx.as<double>(); op.lda = h->ldH; op.a_cols = h->Mx;
    op.B = h->Eext.as<double>(); op.ldb = h->epitch; op.b_cols = h->epitch; op.B2 = op.B; op.ldb2 = op.ldb; op.b2_cols = op.b_cols;
    gemm_set_free_split(true);
    int status = launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, h->Mx, 0, h->kexp, h->colsum.as<double>(), h->ldH, 1, st);
    GemmOperands o2{};
    o2.A = h->A2.as<double>(); o2.lda = h->ld2; o2.a_cols = h->M2;
    o2.B = h->Eext.as<double>(); o2.ldb = h->epitch; o2.b_cols = h->epitch; o2.B2 = o2.B; o2.ldb2 = o2.ldb; o2.b2_cols = o2.b_cols;
    if (status == CRM_OK) status = launch_gemm(GEMM_PLAIN, o2, (int)h->n, 0, h->M2, 0, 1, h->colsum2.as<double>(), h->ld2, 1, st);     // Eext column 0 = ones
    gemm_set_free_split(false);
    CRM_CHECK(status);
    h->colsum_valid = true;
    return CRM_OK;
}

// room for the digit planes of [Hx | Hx.E0_j]?  (cudaMemGetInfo costs milliseconds: only asked for large requests); extra = further
// bytes the caller needs next to them
// Applies a declared structure of the background (crm_set_background_factors): copies the factors, starts the verification against the
// L columns of Hx (result in devinfo[2 R], non-zero = the claim does not hold) and lays out the compact rows.  h->kr is set tentatively:
// the caller reads the flag (at once, or with the read-back that ends the set-up).  *applicable = false: shapes the structure cannot
// describe or the expansion kernel does not take -- nothing is changed, the full basis is used.
static int kr_apply(Handle* h, const double* hK, long long ldhk, int q, const double* M, int r, cudaStream_t st, bool* applicable) {
    *applicable = false;
    h->kr = false;
    const int rp = (r + KR_IT - 1) / KR_IT * KR_IT;
    if ((long long)r * q != h->mL || r > h->k0 || (size_t)h->k0 * rp * sizeof(double) > 48 * 1024) return CRM_OK;
    const long long R1 = h->ldH, R2 = R1 + (long long)h->k0 * h->k1, R3 = R2 + (long long)h->k0 * (1 + h->c);
    const long long rows = round_up(R3 + (long long)(h->k0 * (h->k0 + 1) / 2) * q, 2);     // even: leading dimension of fp64 TMA operands
    if (rows >= (long long)h->kexp * h->ldH) return CRM_OK;                                   // nothing to gain (tiny k0)
    CRM_CHECK(h->krK.reserve((size_t)h->n * q * 8));
    CRM_CHECK(h->krM.reserve((size_t)h->k0 * r * 8));
    int* flag = h->devinfo.as<int>() + 2 * h->R;
    CRM_CUDA(cudaMemcpy2DAsync(h->krK.ptr, (size_t)q * 8, hK, (size_t)ldhk * 8, (size_t)q * 8, (size_t)h->n, cudaMemcpyDeviceToDevice, st));
    CRM_CUDA(cudaMemcpyAsync(h->krM.ptr, M, (size_t)h->k0 * r * 8, cudaMemcpyHostToDevice, st));
    CRM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    kr_check_kernel<<<(unsigned)((h->n + KR_CHECK_CELLS - 1) / KR_CHECK_CELLS), 256, (size_t)2 * KR_CHECK_CELLS * r * sizeof(double), st>>>(
        h->Hx.as<double>(), h->ldH, h->k1, h->Eext.as<double>(), h->epitch, h->k0, h->krK.as<double>(), q, h->krM.as<double>(), r, h->n, flag);
    CRM_CUDA(cudaGetLastError()); count_launch();
    h->kr = true; h->kr_unverified = true; h->kr_q = q; h->kr_r = r;
    h->kr_R1 = R1; h->kr_R2 = R2; h->kr_R3 = R3; h->kr_rows = rows;
    *applicable = true;
    return CRM_OK;
}
static int kr_expand(Handle* h, const double* S, long long lds, long long B, double* C, cudaStream_t st) {
    const int rp = (h->kr_r + KR_IT - 1) / KR_IT * KR_IT;
    const size_t smem = (size_t)h->k0 * rp * sizeof(double);
    if (smem > 48 * 1024) { set_error("structured background: %d x %d context map does not fit shared memory", h->k0, h->kr_r); return CRM_ERR_UNSUPPORTED; }
    kr_expand_kernel<<<(unsigned)std::min<long long>(B, 148 * 64), 256, smem, st>>>(S, lds, h->krM.as<double>(), h->k0, h->kr_r, h->kr_q, h->k1, h->m, h->c, h->ldH,
                                                                                   h->kr_R1, h->kr_R2, h->kr_R3, B, C);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
// CRM_KR=0 (read per call: tests, A/B) ignores a declared Khatri-Rao structure
static bool kr_active(const Handle* h) {
    if (!h->kr) return false;
    const char* v = getenv("CRM_KR");
    return !(v && atoi(v) == 0);
}
// rows of the K-major digit-plane operand: the full expanded basis, or the compact set of a structured background
static long long plane_rows(const Handle* h) { return kr_active(h) ? h->kr_rows : (long long)h->kexp * h->ldH; }
static int planes_fit(Handle* h, double extra, bool* fits) {
    const long long Mtot = plane_rows(h), Mp = round_up(Mtot, 16), Kp = round_up(h->n, 16);
    const size_t a8_bytes = (size_t)OZAKI_SLICES * Mp * Kp;
    *fits = true;
    if ((h->oz_built && h->oz_built_kr == kr_active(h)) || h->A8.cap >= a8_bytes) return CRM_OK;
    static size_t total_mem[16] = {0};
    if (!total_mem[h->device]) { size_t f = 0; CRM_CUDA(cudaMemGetInfo(&f, &total_mem[h->device])); }
    const double need = (double)a8_bytes + extra;
    if (need > 0.2 * (double)total_mem[h->device]) {
        size_t free_b = 0, total_b = 0;
        CRM_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += pool_cached_bytes(h->device);
        if (need > 0.5 * (double)free_b) *fits = false;
    }
    return CRM_OK;
}
// exponents + digit planes of [Hx | Hx.E0_j] (or of its compact form, kr_layout) on stream `on` (buffers reserved by the caller's stream order)
static int build_planes(Handle* h, cudaStream_t on) {
    const long long n = h->n, Mtot = plane_rows(h), Mp = round_up(Mtot, 16), Kp = round_up(n, 16);
    int* expo = h->a8expo.as<int>();
    int8_t* A8 = h->A8.as<int8_t>();
    if (!kr_active(h)) {
        CRM_CHECK(oz_launch_exponents(h->Hx.as<double>(), h->ldH, h->Eext.as<double>(), h->epitch, h->kexp, n, expo, on));
        return oz_launch_slices(h->Hx.as<double>(), h->ldH, h->Eext.as<double>(), h->epitch, 0, h->kexp, n, expo, A8, Mp, Kp, on);
    }
    const double* Hx = h->Hx.as<double>(); const double* Ee = h->Eext.as<double>(); const double* A2 = h->A2.as<double>(); const double* hK = h->krK.as<double>();
    const int ldH = h->ldH, k0 = h->k0, k1 = h->k1, yw = 1 + h->c, q = h->kr_q, npair = k0 * (k0 + 1) / 2;
    CRM_CHECK(oz_launch_fill_exponents(expo, Mtot, on));
    CRM_CHECK(oz_launch_product_exponents(Hx, ldH, ldH, Ee, h->epitch, 0, 1, n, expo, 0, ldH, on));
    CRM_CHECK(oz_launch_product_exponents(Hx, ldH, k1, Ee, h->epitch, 1, k0, n, expo, h->kr_R1, k1, on));
    CRM_CHECK(oz_launch_product_exponents(Hx + h->m, ldH, yw, Ee, h->epitch, 1, k0, n, expo, h->kr_R2, yw, on));
    CRM_CHECK(oz_launch_product_exponents(hK, q, q, A2, h->ld2, 1 + k0, npair, n, expo, h->kr_R3, q, on));
    CRM_CHECK(oz_launch_product_slices(Hx, ldH, ldH, Ee, h->epitch, 0, 1, n, expo, A8, Mp, Kp, 0, ldH, on));
    CRM_CHECK(oz_launch_product_slices(Hx, ldH, k1, Ee, h->epitch, 1, k0, n, expo, A8, Mp, Kp, h->kr_R1, k1, on));
    CRM_CHECK(oz_launch_product_slices(Hx + h->m, ldH, yw, Ee, h->epitch, 1, k0, n, expo, A8, Mp, Kp, h->kr_R2, yw, on));
    return oz_launch_product_slices(hK, q, q, A2, h->ld2, 1 + k0, npair, n, expo, A8, Mp, Kp, h->kr_R3, q, on);
}
// work of the side stream on the digit planes must be over before `st` touches them (or hands their memory on)
static int wait_early_planes(Handle* h, cudaStream_t st) {
    if (!h->planes_ev_pending) return CRM_OK;
    CRM_CUDA(cudaStreamWaitEvent(st, h->planes_ev, 0));
    h->planes_ev_pending = false;
    return CRM_OK;
}
// digit planes on the side stream, ordered after everything enqueued on `st` so far (crm_hint_integer_genotypes; called from the set-up)
static int start_early_planes(Handle* h, cudaStream_t st) {
    static const bool off = [] { const char* v = getenv("CRM_EARLY_PLANES"); return v && atoi(v) == 0; }();
    if (off || h->rotation_mode == 1 || h->oz_built) return CRM_OK;
    const long long Mtot = plane_rows(h), Mp = round_up(Mtot, 16), Kp = round_up(h->n, 16);
    bool fits = true;
    CRM_CHECK(planes_fit(h, 0.0, &fits));
    if (!fits) return CRM_OK;
    CRM_CHECK(wait_early_planes(h, st));
    if (h->A8.reserve((size_t)OZAKI_SLICES * Mp * Kp) != CRM_OK) { cudaGetLastError(); return CRM_OK; }
    CRM_CHECK(h->a8expo.reserve((size_t)Mtot * sizeof(int)));
    if (!h->side_stream) {
        CRM_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        CRM_CUDA(cudaEventCreateWithFlags(&h->side_ev, cudaEventDisableTiming));
        CRM_CUDA(cudaEventCreateWithFlags(&h->planes_ev, cudaEventDisableTiming));
    }
    CRM_CUDA(cudaEventRecord(h->side_ev, st));
    CRM_CUDA(cudaStreamWaitEvent(h->side_stream, h->side_ev, 0));
    CRM_CHECK(build_planes(h, h->side_stream));
    CRM_CUDA(cudaEventRecord(h->planes_ev, h->side_stream));
    h->planes_ev_pending = true;
    h->oz_built = true; h->oz_built_kr = kr_active(h);
    return CRM_OK;
}
// the y column of Hx changed (crm_update_phenotype): only the plane rows that contain it are rebuilt -- y itself and y.E0_j
static int refresh_y_planes(Handle* h, cudaStream_t st) {
    if (!h->oz_built) return CRM_OK;
    CRM_CHECK(wait_early_planes(h, st));
    const bool compact = kr_active(h);
    if (h->oz_built_kr != compact) { h->oz_built = false; return CRM_OK; }
    const long long n = h->n, Mtot = plane_rows(h), Mp = round_up(Mtot, 16), Kp = round_up(n, 16);
    const double* ycol = h->Hx.as<double>() + h->m; const double* Ee = h->Eext.as<double>();
    int* expo = h->a8expo.as<int>(); int8_t* A8 = h->A8.as<int8_t>();
    struct Part { int j0, nj; long long row0, rstride; } parts[2];
    int np = 0;
    if (compact) { parts[np++] = {0, 1, (long long)h->m, (long long)h->ldH}; parts[np++] = {1, h->k0, h->kr_R2, (long long)(1 + h->c)}; }
    else parts[np++] = {0, h->kexp, (long long)h->m, (long long)h->ldH};
    for (int i = 0; i < np; i++) {
        const Part& pt = parts[i];
        CRM_CHECK(oz_launch_fill_exponents_strided(expo, pt.row0, pt.rstride, pt.nj, st));
        CRM_CHECK(oz_launch_product_exponents(ycol, h->ldH, 1, Ee, h->epitch, pt.j0, pt.nj, n, expo, pt.row0, pt.rstride, st));
        CRM_CHECK(oz_launch_product_slices(ycol, h->ldH, 1, Ee, h->epitch, pt.j0, pt.nj, n, expo, A8, Mp, Kp, pt.row0, pt.rstride, st));
    }
    return CRM_OK;
}
static int rotation_int8_split(Handle* h, const GBlock& blk, double* C, cudaStream_t st, int* used) {
    *used = 0;
    const long long B = blk.b;
    const bool compact = kr_active(h);
    const long long Mfull = (long long)h->kexp * h->ldH;
    const long long n = h->n, Mtot = plane_rows(h), Mp = round_up(Mtot, 16), Kp = round_up(n, 16), Bp = round_up(B, 16);
    const size_t a8_bytes = (size_t)OZAKI_SLICES * Mp * Kp;
    {
        bool fits = true;
        CRM_CHECK(planes_fit(h, int8_route_library() ? (double)OZAKI_SLICES * Mp * Bp * 4.0 : 0.0, &fits));
        if (!fits) return CRM_OK;
    }
    PhaseTrace tr(st);
    h->oz_block_valid = false;
    CRM_CHECK(h->Gt8.reserve((size_t)Bp * Kp));
    CRM_CHECK(h->G2t8.reserve((size_t)Bp * Kp));
    CRM_CHECK(h->ozflags.reserve(64));
    int flags[4] = {0, 0, 0, 0};
    bool affine = false;
    if (blk.G8 && blk.gmax >= 0) {
        // int8 dosages whose range the host already knows (converted by the feeder): no flag read-back, no host synchronisation
        if ((double)Kp * 64.0 * (double)std::max(blk.gmax, 1) >= 2147483648.0) return CRM_OK;
        CRM_CHECK(oz_launch_transpose_i8(blk.G8, blk.ld8, n, B, h->Gt8.as<int8_t>(), h->G2t8.as<int8_t>(), Bp, Kp, nullptr, st));
        flags[1] = blk.gmax;
    } else {
        if (blk.G8) CRM_CHECK(oz_launch_transpose_i8(blk.G8, blk.ld8, n, B, h->Gt8.as<int8_t>(), h->G2t8.as<int8_t>(), Bp, Kp, h->ozflags.as<int>(), st));
        else CRM_CHECK(oz_launch_genotypes(blk.G, blk.ld, n, B, h->Gt8.as<int8_t>(), h->G2t8.as<int8_t>(), Bp, Kp, h->ozflags.as<int>(), st));
        CRM_CUDA(cudaMemcpyAsync(flags, h->ozflags.ptr, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CRM_CUDA(cudaStreamSynchronize(st));
        if (flags[2] != 0) { set_error("There are non-finite values in the genotype matrix (SNP columns %lld..%lld).", blk.s0, blk.s0 + B - 1); return CRM_ERR_NONFINITE; }
        if (flags[0] != 0) {
            // not integer dosages: every column an affine image a d + b of small integers (standardised / centred dosages)?
            const char* aff_env = getenv("CRM_AFFINE");          // CRM_AFFINE=0: float64 route for every non-integer block (tests, A/B)
            const bool affine_on = !(aff_env && atoi(aff_env) == 0);
            if (!affine_on || !blk.G) return CRM_OK;
            const long long lda = round_up(B, 2);
            CRM_CHECK(h->aff.reserve((size_t)3 * lda * 8));
            CRM_CHECK(h->affscratch.reserve(oz_affine_scratch_bytes(B)));
            CRM_CHECK(oz_launch_affine_genotypes(blk.G, blk.ld, n, B, h->affscratch.ptr, h->aff.as<double>(), lda, h->Gt8.as<int8_t>(), h->G2t8.as<int8_t>(), Bp, Kp,
                                                 h->ozflags.as<int>(), st));
            CRM_CUDA(cudaMemcpyAsync(flags, h->ozflags.ptr, sizeof(flags), cudaMemcpyDeviceToHost, st));
            CRM_CUDA(cudaStreamSynchronize(st));
            if (flags[0] != 0) return CRM_OK;                                          // real-valued genotypes: the fp64 route
            affine = true;
        }
        if ((double)Kp * 64.0 * (double)std::max(flags[1], 1) >= 2147483648.0) return CRM_OK;   // int32 accumulation could overflow
    }
    tr.mark("genotypes->int8");
    CRM_CHECK(wait_early_planes(h, st));
    if (!h->oz_built || h->oz_built_kr != compact) {
        if (h->A8.reserve(a8_bytes) != CRM_OK) return CRM_OK;      // no room after all: the fp64 route takes over
        CRM_CHECK(h->a8expo.reserve((size_t)Mtot * sizeof(int)));
        CRM_CHECK(build_planes(h, st));
        h->oz_built = true; h->oz_built_kr = compact;
        tr.mark("digit planes");
    }
    double* Crot = C;                 // where the contraction writes: the full layout, or the compact one (expanded below)
    if (compact) { CRM_CHECK(h->Cc.reserve((size_t)B * Mtot * 8)); Crot = h->Cc.as<double>(); }
    if (h->prof_on) {
        cudaEvent_t e0, e1;
        CRM_CUDA(cudaEventCreate(&e0)); CRM_CUDA(cudaEventCreate(&e1));
        CRM_CUDA(cudaEventRecord(e0, st));
        h->prof_oz_events.push_back(e0); h->prof_oz_events.push_back(e1);
        h->prof_oz_gemm_ops += 2.0 * (double)OZAKI_SLICES * (double)Mp * (double)Kp * (double)Bp;
    }
    CRM_CHECK(int8_split_contract(h, h->A8.as<int8_t>(), Mp, Mtot, h->a8expo.as<int>(), h->Gt8.as<int8_t>(), Bp, B, Kp, Crot, Mtot, st));
    if (h->prof_on) CRM_CUDA(cudaEventRecord(h->prof_oz_events.back(), st));
    tr.mark("int8 contraction");
    if (compact) {
        CRM_CHECK(kr_expand(h, Crot, Mtot, B, C, st));
        tr.mark("expand");
    }
    if (affine) {
        CRM_CHECK(ensure_column_sums(h, st));
        CRM_CHECK(oz_launch_affine_fix(C, Mfull, B, Mfull, h->aff.as<double>(), round_up(B, 2), h->colsum.as<double>(), st));
        tr.mark("affine map");
    }
    tr.report("int8 rotation");
    h->oz_block_valid = true;
    h->oz_block_gmax = flags[1];
    h->oz_block_affine = affine;
    *used = 1;
    return CRM_OK;
}

// non-finite genotypes make every later stage meaningless (the reference's LMM raises on them): checked where no other kernel does
static int check_block_finite(Handle* h, const GBlock& blk, cudaStream_t st) {
    if (!blk.G) return CRM_OK;          // int8 images are finite by construction
    CRM_CHECK(h->ozflags.reserve(64));
    int flags[4] = {0, 0, 0, 0};
    CRM_CUDA(cudaMemsetAsync(h->ozflags.ptr, 0, sizeof(flags), st));
    CRM_CHECK(oz_launch_finite_check(blk.G, blk.ld, h->gs->K, blk.b, h->ozflags.as<int>(), st));
    CRM_CUDA(cudaMemcpyAsync(flags, h->ozflags.ptr, sizeof(flags), cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    if (flags[2] != 0) { set_error("There are non-finite values in the genotype matrix (SNP columns %lld..%lld).", blk.s0, blk.s0 + blk.b - 1); return CRM_ERR_NONFINITE; }
    return CRM_OK;
}

// Pre-expanded basis of the fp64 route when it fits comfortably (CRM_NO_HXE=1 forces the on-the-fly route).  Decided by the first
// float64 rotation of a model, not by the set-up: the memory query costs milliseconds and integer dosages never get here.
static int decide_hxe(Handle* h) {
    if (h->hxe_decided) return CRM_OK;
    size_t free_b = 0, total_b = 0;
    CRM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    free_b += pool_cached_bytes(h->device);     // memory cached in the allocation pool is reusable as well
    const char* env = getenv("CRM_NO_HXE");
    const char* envb = getenv("CRM_HXE_BLOCKS");       // tests: force streaming in groups of this many context blocks
    const double budget = 0.45 * (double)(free_b + h->HxE.cap);
    const double block_bytes = (double)h->n * h->ldH * 8.0;
    int blocks = (int)std::min<double>(h->kexp, std::floor(budget / block_bytes));
    if (envb && atoi(envb) > 0) blocks = std::min(blocks, atoi(envb));
    h->use_hxe = !(env && atoi(env) != 0) && blocks >= 1 && (double)blocks * h->ldH < 2.0e9;
    h->hxe_blocks = h->use_hxe ? blocks : 0;
    if (!h->use_hxe) h->HxE.release();
    h->hxe_decided = true;
    return CRM_OK;
}

static int launch_rotation(Handle* h, GBlock& blk, double* C, cudaStream_t st) {
    const Handle::GenoSpace& gs = *h->gs;
    const long long ldE = (long long)h->kexp * h->ldH, B = blk.b;
    GemmOperands op{};
    if (h->gs == &h->donors) {                       // donor-level operands are always fully expanded (d rows only)
        CRM_CHECK(block_f64(h, blk, st));
        CRM_CHECK(check_block_finite(h, blk, st));
        op.B = blk.G; op.ldb = blk.ld; op.b_cols = blk.cols; op.B2 = blk.G; op.ldb2 = blk.ld; op.b2_cols = blk.cols;
        op.A = gs.HxE; op.lda = gs.ldE; op.a_cols = gs.ldE;
        return launch_gemm(GEMM_PLAIN, op, (int)gs.K, 0, (int)gs.ldE, 0, (int)B, C, gs.ldE, 1, st);
    }
    if (h->rotation_mode != 1) {
        int used = 0;
        CRM_CHECK(rotation_int8_split(h, blk, C, st, &used));
        if (used) return CRM_OK;
        if (h->rotation_mode == 2) { set_error("CRM_ROTATION=int8 but the genotype block is not integer-valued in [-127, 127] (or memory is short)"); return CRM_ERR_UNSUPPORTED; }
    }
    CRM_CHECK(block_f64(h, blk, st));
    if (h->rotation_mode == 1) CRM_CHECK(check_block_finite(h, blk, st));     // (the int8 conversion kernel reported it otherwise)
    op.B = blk.G; op.ldb = blk.ld; op.b_cols = blk.cols; op.B2 = blk.G; op.ldb2 = blk.ld; op.b2_cols = blk.cols;
    CRM_CHECK(decide_hxe(h));
    if (h->use_hxe && h->hxe_blocks == h->kexp && kr_active(h)) {
        // structured background, basis resident: plain contraction against the compact basis, then the expansion (see kr_layout)
        const long long rows = h->kr_rows;
        CRM_CHECK(h->HxE.reserve((size_t)h->n * rows * 8));
        if (!(h->hxe_built && h->hxe_built_kr)) {
            build_hxe_compact_kernel<<<(unsigned)std::min<long long>(h->n, 148 * 16), 256, 0, st>>>(h->Hx.as<double>(), h->ldH, h->k1, h->m, h->c, h->Eext.as<double>(), h->epitch,
                                                                                                h->k0, h->A2.as<double>(), h->ld2, h->krK.as<double>(), h->kr_q, h->kr_R1,
                                                                                                h->kr_R2, h->kr_R3, rows, h->n, h->HxE.as<double>());
            CRM_CUDA(cudaGetLastError()); count_launch();
        }
        CRM_CHECK(h->Cc.reserve((size_t)B * rows * 8));
        op.A = h->HxE.as<double>(); op.lda = rows; op.a_cols = rows;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, (int)rows, 0, (int)B, h->Cc.as<double>(), rows, 1, st));
        h->hxe_built = true; h->hxe_built_kr = true;
        return kr_expand(h, h->Cc.as<double>(), rows, B, C, st);
    }
    if (h->use_hxe) {
        const int nb = h->hxe_blocks;
        CRM_CHECK(h->HxE.reserve((size_t)h->n * nb * h->ldH * 8));
        for (int j0 = 0; j0 < h->kexp; j0 += nb) {
            const int nj = std::min(nb, h->kexp - j0);
            if (!(nb == h->kexp && h->hxe_built && !h->hxe_built_kr)) {
                build_hxe_kernel<<<(unsigned)std::min<long long>(h->n, 148 * 16), 256, 0, st>>>(h->Hx.as<double>(), h->ldH, h->Eext.as<double>(), h->epitch, j0, nj, h->n,
                                                                                          h->HxE.as<double>());
                CRM_CUDA(cudaGetLastError()); count_launch();
            }
            op.A = h->HxE.as<double>(); op.lda = (long long)nj * h->ldH; op.a_cols = op.lda;
            CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, (int)op.lda, 0, (int)B, C + (long long)j0 * h->ldH, ldE, 1, st));
        }
        h->hxe_built = (nb == h->kexp); h->hxe_built_kr = false;
        return CRM_OK;
    }
    op.A = h->Hx.as<double>(); op.lda = h->ldH; op.a_cols = h->Mx;
    op.B2 = h->Eext.as<double>(); op.ldb2 = h->epitch; op.b2_cols = h->epitch;
    return launch_gemm(GEMM_EXPAND, op, (int)h->n, 0, h->Mx, 0, (int)(B * h->kexp), C, h->ldH, h->kexp, st);
}

// Last step of the set-up once S and T of every grid point are in place: rotated null design per rho, kept ranks.
static int finish_setup(Handle* h, cudaStream_t st) {
    const int R = h->R, mp = h->mp, m = h->m, c = h->c, ldH = h->ldH;
    PhaseTrace tr(st);
    rotate_null_kernel<<<blocks_for((long long)R * mp, 128), 128, 0, st>>>(h->Tt.as<double>(), (long long)R * mp, h->gram.as<double>(), ldH, m,
                                                                         mp, R, c, h->yr.as<double>(), h->Wr.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    CRM_CHECK(h->YW.reserve((size_t)R * (1 + c) * mp * 8));
    CRM_CHECK(h->ywgram.reserve((size_t)(1 + c) * (1 + c) * 8));
    build_yw_kernel<<<blocks_for(std::max<long long>((long long)R * (1 + c) * mp, (long long)(1 + c) * (1 + c)), 256), 256, 0, st>>>(
        h->yr.as<double>(), h->Wr.as<double>(), h->stats.as<double>(), R, c, mp, h->YW.as<double>(), h->ywgram.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    std::vector<int> info(2 * R + 1, 0);
    CRM_CUDA(cudaMemcpyAsync(info.data(), h->devinfo.as<int>(), (size_t)(2 * R + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    if (h->kr_unverified) {               // verdict on the structure declared before the set-up (kr_apply)
        h->kr_unverified = false;
        if (info[2 * R] != 0) { h->kr = false; if (h->oz_built && h->oz_built_kr) h->oz_built = false; }
    }
    h->max_rank = 0;
    for (int r = 0; r < R; r++) {
        if (info[r] != 0) { set_error("cusolverDnDsyevd did not converge for grid point %d (devInfo=%d)", r, info[r]); return CRM_ERR_SOLVER; }
        h->max_rank = std::max(h->max_rank, info[R + r]);
    }
    h->ready = true;
    tr.mark("rotate null");
    tr.report("set-up (finish)");
    if (trace_on()) fprintf(stderr, "[crm trace] host %.1f ms: set-up returns\n", host_ms());
    return CRM_OK;
}

// One grid point of the basis as a packed record of 2 + mp + m * mp doubles: [kept rank, solver info, S (mp), T block (m x mp)].
__global__ void pack_basis_kernel(const double* S, const double* Tt, long long ldt, const int* info, int R, int r, int m, int mp, double* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 2 + (long long)mp + (long long)m * mp;
    if (idx >= total) return;
    double v;
    if (idx == 0) v = (double)info[R + r];
    else if (idx == 1) v = (double)info[r];
    else if (idx < 2 + mp) v = S[(long long)r * mp + (idx - 2)];
    else { const long long t = idx - 2 - mp; const long long a = t / mp, i = t - a * mp; v = Tt[a * ldt + (long long)r * mp + i]; }
    out[idx] = v;
}
__global__ void unpack_basis_kernel(const double* in, int R, int r, int m, int mp, double* S, double* Tt, long long ldt, int* info) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 2 + (long long)mp + (long long)m * mp;
    if (idx >= total) return;
    const double v = in[idx];
    if (idx == 0) info[R + r] = (int)v;
    else if (idx == 1) info[r] = (int)v;
    else if (idx < 2 + mp) S[(long long)r * mp + (idx - 2)] = v;
    else { const long long t = idx - 2 - mp; const long long a = t / mp, i = t - a * mp; Tt[a * ldt + (long long)r * mp + i] = v; }
}

static int do_setup(Handle* h, const double* y, const double* W, long long ldw, const double* E0, long long lde0, const double* E1,
                    long long lde1, const double* L, long long ldl, long long n, int c, int k0, int k1, long long mL,
                    const double* rho_host, int R, int r_first, int r_step, cudaStream_t st) {
    if (!y || !W || !E0 || !E1 || n <= 0 || c <= 0 || k0 <= 0 || k1 <= 0 || mL < 0 || R <= 0 || (mL > 0 && !L)) { set_error("crm_setup: bad arguments"); return CRM_ERR_INVALID; }
    if (n > 2000000000LL) { set_error("n too large"); return CRM_ERR_UNSUPPORTED; }
    if (c > 60) { set_error("at most 60 covariate columns are supported (got %d)", c); return CRM_ERR_UNSUPPORTED; }
    if (R > 64) { set_error("rho grid too long"); return CRM_ERR_UNSUPPORTED; }
    if (r_step < 1 || r_first < 0 || r_first >= r_step) { set_error("crm_setup: bad grid-point selection %d mod %d", r_first, r_step); return CRM_ERR_INVALID; }
    if (k1 + mL > 30000) { set_error("background half-covariance has %lld columns; limit is 30000", (long long)(k1 + mL)); return CRM_ERR_UNSUPPORTED; }
    h->ready = false;
    h->donors_set = false;
    h->oz_built = false; h->oz_built_a2 = false; h->kr = false;
    h->n = n; h->c = c; h->k0 = k0; h->k1 = k1; h->mL = mL; h->R = R;
    h->m = (int)(k1 + mL);
    h->mp = (int)round_up(h->m, 2);
    h->Mx = h->m + 1 + c;
    h->ldH = (int)round_up(h->Mx, 2);
    h->kexp = 1 + k0;
    {   // Eext pitch: >= kexp + 1 (zero column), {0,p,2p,3p} mod 16 spaced by >= 4
        int pch = (h->kexp + 2) & ~1;
        while (!(pch % 16 == 4 || pch % 16 == 12)) pch += 2;
        h->epitch = pch;
    }
    h->M2 = 1 + k0 + k0 * (k0 + 1) / 2;
    h->ld2 = (int)round_up(h->M2, 2);
    h->rho.assign(rho_host, rho_host + R);
    const int m = h->m, mp = h->mp, Mx = h->Mx, ldH = h->ldH;

    SlowSection* sec = new SlowSection("set-up: reserve");
    CRM_CHECK(h->Hx.reserve((size_t)n * ldH * 8));
    CRM_CHECK(h->Eext.reserve((size_t)n * h->epitch * 8));
    CRM_CHECK(h->A2.reserve((size_t)n * h->ld2 * 8));
    CRM_CHECK(h->gram.reserve((size_t)ldH * ldH * 8));
    CRM_CHECK(h->S.reserve((size_t)R * mp * 8));
    CRM_CHECK(h->yr.reserve((size_t)R * mp * 8));
    CRM_CHECK(h->Wr.reserve((size_t)R * c * mp * 8));
    CRM_CHECK(h->Tt.reserve((size_t)m * R * mp * 8));
    CRM_CHECK(h->stats.reserve((size_t)(1 + c + c * c + R + 4) * 8));
    CRM_CHECK(h->devinfo.reserve((size_t)(2 * R + 8) * sizeof(int)));

    delete sec;
    PhaseTrace tr(st);
    // Hx = [E1 | L | y | W]
    double* Hx = h->Hx.as<double>();
    sec = new SlowSection("set-up: operand copies + test contexts");
    CRM_CUDA(cudaMemsetAsync(Hx, 0, (size_t)n * ldH * 8, st));
    CRM_CUDA(cudaMemcpy2DAsync(Hx, (size_t)ldH * 8, E1, (size_t)lde1 * 8, (size_t)k1 * 8, (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (mL > 0) CRM_CUDA(cudaMemcpy2DAsync(Hx + k1, (size_t)ldH * 8, L, (size_t)ldl * 8, (size_t)mL * 8, (size_t)n, cudaMemcpyDeviceToDevice, st));
    CRM_CUDA(cudaMemcpy2DAsync(Hx + m, (size_t)ldH * 8, y, 8, 8, (size_t)n, cudaMemcpyDeviceToDevice, st));
    CRM_CUDA(cudaMemcpy2DAsync(Hx + m + 1, (size_t)ldH * 8, W, (size_t)ldw * 8, (size_t)c * 8, (size_t)n, cudaMemcpyDeviceToDevice, st));
    h->hxe_decided = false;     // the fp64 route decides about its pre-expanded basis when it first runs (decide_hxe)
    {
        const char* rm = getenv("CRM_ROTATION");      // "dmma": fp64 tensor cores only; "int8": exact int8 split required; default auto
        h->rotation_mode = (rm && !strcmp(rm, "dmma")) ? 1 : (rm && !strcmp(rm, "int8")) ? 2 : 0;
    }
    CRM_CHECK(build_test_contexts(h, E0, lde0, st));
    h->kr_unverified = false;
    if (h->kr_pending) {            // structure declared ahead of the set-up: verified on the device now, the verdict is read with the solver status below
        bool applicable = false;
        CRM_CHECK(kr_apply(h, h->kr_pending_hK, h->kr_pending_ld, h->kr_pending_q, h->kr_pending_M, h->kr_pending_r, st, &applicable));
        h->kr_pending = false;
    }
    delete sec;
    sec = new SlowSection("set-up: gram launch");

    // Gram of [H | y | W] by the K1 kernel (plain mode): H'H, H'y, H'W, y'y, W'y, W'W
    GemmOperands op{};
    op.A = Hx; op.lda = ldH; op.a_cols = Mx;
    op.B = Hx; op.ldb = ldH; op.b_cols = Mx;
    op.B2 = Hx; op.ldb2 = ldH; op.b2_cols = Mx;
    gemm_set_free_split(true); gemm_set_symmetric(true);
    const int gram_status = launch_gemm(GEMM_PLAIN, op, (int)n, 0, Mx, 0, Mx, h->gram.as<double>(), ldH, 1, st);
    gemm_set_free_split(false); gemm_set_symmetric(false);
    CRM_CHECK(gram_status);
    extract_stats_kernel<<<1, 1024, 0, st>>>(h->gram.as<double>(), ldH, m, c, h->stats.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    delete sec;
    tr.mark("operands+gram");

    // per-rho eigendecomposition (cuSOLVER, one-off per gene) with one pooled cuSOLVER context per device.  Measured on
    // B200: Dsyevd takes 12 ms per 1020 x 1020 problem and R of them on R streams (with or without one host thread each)
    // do not overlap.
    if (h->device < 0 || h->device >= 16) { set_error("device index %d outside the supported range", h->device); return CRM_ERR_UNSUPPORTED; }
    EigPool& pool = g_eig_pool[h->device];
    std::lock_guard<std::mutex> pool_lock(pool.mu);
    if (pool.ctx.empty()) pool.ctx.resize(1);
    EigCtx& e = pool.ctx[0];
    if (!e.solver) CRM_SOLVER(cusolverDnCreate(&e.solver));
    CRM_SOLVER(cusolverDnSetStream(e.solver, st));
    CRM_CHECK(e.mat.reserve((size_t)m * m * 8));
    CRM_CHECK(e.val.reserve((size_t)m * 8));
    int lwork = 0;
    CRM_SOLVER(cusolverDnDsyevd_bufferSize(e.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, e.mat.as<double>(), m, e.val.as<double>(), &lwork));
    CRM_CHECK(e.work.reserve((size_t)lwork * 8));
    int* info_dev = h->devinfo.as<int>();
    int* rank_dev = h->devinfo.as<int>() + R;
    const int tall = n > m ? 1 : 0;
    // grid points decomposed by this call: all of them, or every r_step-th one starting at r_first when the set-up is shared between
    // the ranks of a multi-GPU scan (the others arrive through crm_import_basis before crm_setup_finish)
    std::vector<int> sel;
    for (int r = r_first; r < R; r += r_step) sel.push_back(r);
    const int nsel = (int)sel.size();
    CRM_CUDA(cudaMemsetAsync(info_dev, 0, (size_t)2 * R * 4, st));
    // All grid points go through the batched solver of eig.cuh (tridiagonalisation of every matrix at once on its own group of
    // SMs, multisection + inverse iteration + re-orthonormalisation, back-transformation): 33 ms instead of 122 ms for 11 problems
    // of size 1020 (profiles/r01_eig_bench.txt).  The sequential cusolverDnDsyevd calls remain as the fall-back when a size is out
    // of range or the residual / orthogonality check of a matrix fails, and as the reference point (CRM_EIG=cusolver).
    static const bool native = [] { const char* v = getenv("CRM_EIG"); return !(v && !strcmp(v, "cusolver")); }();
    bool native_done = false;
    if (native && m >= 2 && m <= 4096 && R <= 64 && nsel > 0) {
        std::vector<int> n_of(nsel), a0_of(nsel);
        bool ok = true;
        for (int j = 0; j < nsel; j++) {
            const int r = sel[j];
            int a0 = 0, ms = m;
            if (mL > 0 && h->rho[r] == 1.0) { a0 = 0; ms = k1; }
            else if (mL > 0 && h->rho[r] == 0.0) { a0 = k1; ms = (int)mL; }
            n_of[j] = ms; a0_of[j] = a0;
            if (ms < 2) ok = false;
        }
        if (ok) {
            CRM_CHECK(ensure_blas(e));
            size_t ws_bytes = 0;
            CRM_CHECK(eig_workspace_bytes(m, nsel, &ws_bytes));
            CRM_CHECK(e.mat.reserve((size_t)nsel * m * m * 8)); CRM_CHECK(e.vec.reserve((size_t)nsel * m * m * 8)); CRM_CHECK(e.val.reserve((size_t)nsel * m * 8));
            CRM_CHECK(e.ws.reserve(ws_bytes)); CRM_CHECK(e.quality.reserve((size_t)nsel * 8 + (size_t)2 * nsel * 4));
            int lib_lwork = 0;
            CRM_CHECK(eig_lib_lwork(e.solver, m, &lib_lwork));
            CRM_CHECK(e.work.reserve((size_t)std::max(lib_lwork, lwork) * 8));
            for (int j = 0; j < nsel; j++) {
                scale_gram_kernel<<<blocks_for((long long)n_of[j] * n_of[j], 256), 256, 0, st>>>(h->gram.as<double>(), ldH, a0_of[j], n_of[j], k1, h->rho[sel[j]], e.mat.as<double>() + (size_t)j * m * m);
                CRM_CUDA(cudaGetLastError()); count_launch();
            }
            int* lib_info = reinterpret_cast<int*>(e.quality.as<double>() + nsel);
            CRM_CUDA(cudaMemsetAsync(lib_info, 0, (size_t)2 * nsel * 4, st));
            tr.mark("scaled grams");
            // The digit planes of the basis do not depend on the decompositions: with crm_hint_integer_genotypes they are built on a side
            // stream once the tridiagonalisation (a cooperative kernel that wants every SM) has finished, next to the latency-bound phases
            // that follow it (multisection, inverse iteration, back-transformation of 1 020-column problems leave most SMs idle).
            std::function<int()> planes_hook = [h, st]() -> int { return start_early_planes(h, st); };
            CRM_CHECK(eig_batched(e.solver, e.blas, e.mat.as<double>(), n_of.data(), m, nsel, e.val.as<double>(), e.vec.as<double>(), e.quality.as<double>(), e.ws.ptr,
                                  e.work.as<double>(), std::max(lib_lwork, lwork), lib_info, st, R, sel.data(), h->early_planes ? &planes_hook : nullptr));
            tr.mark("batched eigensolver");
            std::vector<double> q(nsel); std::vector<int> li(2 * nsel);
            CRM_CUDA(cudaMemcpyAsync(q.data(), e.quality.ptr, (size_t)nsel * 8, cudaMemcpyDeviceToHost, st));
            CRM_CUDA(cudaMemcpyAsync(li.data(), lib_info, (size_t)2 * nsel * 4, cudaMemcpyDeviceToHost, st));
            CRM_CUDA(cudaStreamSynchronize(st));
            native_done = true;
            for (int j = 0; j < nsel; j++) if (!(q[j] < 1e-11) || li[j] != 0 || li[nsel + j] != 0) native_done = false;
            static const bool verbose = [] { const char* v = getenv("CRM_TRACE"); return v && atoi(v) != 0; }();
            if (verbose) { fprintf(stderr, "[crm trace] native eigensolver: %s; residuals", native_done ? "accepted" : "rejected"); for (int j = 0; j < nsel; j++) fprintf(stderr, " %.1e", q[j]); fprintf(stderr, "\n"); }
            if (native_done) {
                for (int j = 0; j < nsel; j++) {
                    build_basis_kernel<<<std::min(1024u, blocks_for((long long)m * mp, 256)), 256, 0, st>>>(
                        e.vec.as<double>() + (size_t)j * m * m, e.val.as<double>() + (size_t)j * m, m, mp, a0_of[j], n_of[j], k1, h->rho[sel[j]], tall,
                        h->S.as<double>() + (long long)sel[j] * mp, h->Tt.as<double>(), (long long)R * mp, sel[j], rank_dev);
                    CRM_CUDA(cudaGetLastError()); count_launch();
                }
            }
        }
    }
    if (!native_done) {
        for (int r : sel) {
            // rho = 1 / rho = 0 zero one diagonal block of D: only the surviving block is decomposed
            int a0 = 0, ms = m;
            if (mL > 0 && h->rho[r] == 1.0) { a0 = 0; ms = k1; }
            else if (mL > 0 && h->rho[r] == 0.0) { a0 = k1; ms = (int)mL; }
            scale_gram_kernel<<<blocks_for((long long)ms * ms, 256), 256, 0, st>>>(h->gram.as<double>(), ldH, a0, ms, k1, h->rho[r], e.mat.as<double>());
            CRM_CUDA(cudaGetLastError()); count_launch();
            int lw = 0;
            CRM_SOLVER(cusolverDnDsyevd_bufferSize(e.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, ms, e.mat.as<double>(), ms, e.val.as<double>(), &lw));
            if (lw > lwork) { set_error("cuSOLVER workspace for the %d x %d block exceeds the one of the full problem", ms, ms); return CRM_ERR_SOLVER; }
            CRM_SOLVER(cusolverDnDsyevd(e.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, ms, e.mat.as<double>(), ms, e.val.as<double>(),
                                        e.work.as<double>(), lwork, info_dev + r));
            build_basis_kernel<<<std::min(1024u, blocks_for((long long)m * mp, 256)), 256, 0, st>>>(
                e.mat.as<double>(), e.val.as<double>(), m, mp, a0, ms, k1, h->rho[r], tall, h->S.as<double>() + (long long)r * mp,
                h->Tt.as<double>(), (long long)R * mp, r, rank_dev);
            CRM_CUDA(cudaGetLastError()); count_launch();
        }
    }
    tr.mark("eigendecompositions");
    tr.report("set-up (operands, Gram, eigendecompositions)");
    if (r_step > 1) {                       // the caller exchanges the grid points between ranks, then calls crm_setup_finish
        CRM_CUDA(cudaStreamSynchronize(st));       // the pooled solver buffers go back with the lock
        return CRM_OK;
    }
    return finish_setup(h, st);
}

// New phenotype for the same cells, contexts, covariates and background (scans of many genes over one data set): only the
// y-dependent quantities are refreshed -- H'y, y'y, W'y, the rotated phenotype per rho, the y column of the expanded bases --
// the Gram of H and the R eigendecompositions are kept.
static int do_update_phenotype(Handle* h, const double* y, cudaStream_t st) {
    if (!h->ready) { set_error("crm_update_phenotype: handle is not set up"); return CRM_ERR_STATE; }
    if (!y) { set_error("crm_update_phenotype: null phenotype"); return CRM_ERR_INVALID; }
    const int m = h->m, mp = h->mp, c = h->c, R = h->R, ldH = h->ldH, Mx = h->Mx;
    double* Hx = h->Hx.as<double>();
    CRM_CUDA(cudaMemcpy2DAsync(Hx + m, (size_t)ldH * 8, y, 8, 8, (size_t)h->n, cudaMemcpyDeviceToDevice, st));
    GemmOperands op{};
    op.A = Hx; op.lda = ldH; op.a_cols = Mx; op.B = Hx; op.ldb = ldH; op.b_cols = Mx; op.B2 = Hx; op.ldb2 = ldH; op.b2_cols = Mx;
    gemm_set_free_split(true);
    const int status = launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, Mx, m, 1, h->gram.as<double>() + (long long)m * ldH, ldH, 1, st);
    gemm_set_free_split(false);
    CRM_CHECK(status);
    symmetrise_row_kernel<<<blocks_for(Mx, 128), 128, 0, st>>>(h->gram.as<double>(), ldH, m, Mx);
    CRM_CUDA(cudaGetLastError()); count_launch();
    extract_stats_kernel<<<1, 1024, 0, st>>>(h->gram.as<double>(), ldH, m, c, h->stats.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    rotate_null_kernel<<<blocks_for((long long)R * mp, 128), 128, 0, st>>>(h->Tt.as<double>(), (long long)R * mp, h->gram.as<double>(), ldH, m, mp, R, c,
                                                                         h->yr.as<double>(), h->Wr.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    build_yw_kernel<<<blocks_for(std::max<long long>((long long)R * (1 + c) * mp, (long long)(1 + c) * (1 + c)), 256), 256, 0, st>>>(
        h->yr.as<double>(), h->Wr.as<double>(), h->stats.as<double>(), R, c, mp, h->YW.as<double>(), h->ywgram.as<double>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    CRM_CHECK(refresh_y_planes(h, st));     // the digit planes of the y column (and their exponents) change with the phenotype
    h->colsum_valid = false;
    if (h->hxe_built && h->hxe_built_kr) h->hxe_built = false;        // compact fp64 basis: rebuilt by the next float64 rotation
    else if (h->use_hxe && h->hxe_built && h->hxe_blocks == h->kexp) {
        refresh_y_hxe_kernel<<<blocks_for(h->n * h->kexp, 256), 256, 0, st>>>(Hx, ldH, m, h->Eext.as<double>(), h->epitch, h->kexp, h->n, h->HxE.as<double>());
        CRM_CUDA(cudaGetLastError()); count_launch();
    }
    if (h->donors_set) {
        refresh_y_donor_kernel<<<blocks_for(h->donors.K * h->kexp, 128), 128, 0, st>>>(Hx, ldH, m, h->Eext.as<double>(), h->epitch, h->kexp, h->dperm.as<int>(),
                                                                                   h->doff.as<int>(), h->donors.K, h->HxE_D.as<double>());
        CRM_CUDA(cudaGetLastError()); count_launch();
    }
    return CRM_OK;
}

// ------------------------------------------------------------------------------------------------
// interaction scan
// ------------------------------------------------------------------------------------------------
// What the bracket searches of all SNPs of a batch share at their common points is tabulated once (fit.cuh: FIT_TAB_*; CRM_FIT_TABLE=0: off);
// fa must describe the per-rho vectors of the fits that follow.
static int attach_fit_table(Handle* h, FitArgs& fa, cudaStream_t st) {
    const char* ft = getenv("CRM_FIT_TABLE");
    fa.tab = nullptr;
    if ((ft && atoi(ft) == 0) || fa.fixed_x || fa.c < 1 || fa.c > 7) return CRM_OK;
    fa.tab_k = fit_table_points();
    CRM_CHECK(h->fittab.reserve(fit_table_bytes(fa.R, fa.mp)));
    CRM_CHECK(launch_fit_table(fa, h->fittab.as<double>(), st));
    fa.tab = h->fittab.as<double>();
    return CRM_OK;
}

struct BatchPlan { long long batch; };

static long long pick_batch(const Handle* h, long long p, bool interaction) {
    // bytes of workspace per SNP
    const double per_snp = interaction
        ? 8.0 * ((double)h->kexp * h->ldH + h->ld2 + (double)h->R * h->mp + 2.0 * (double)h->k0 * h->mp + h->m)
        : 8.0 * ((double)h->ldH + 2.0 * h->mp + h->m);
    long long b = (long long)(6.0e9 / per_snp);
    b = std::max(64LL, std::min(b, p));
    b = std::min(b, (long long)(65535LL * GEMM_TILE_N / h->kexp));
    return b;
}

// Column-block size for host-resident genotypes: two staging buffers of at most ~4 GB each, blocks of equal size (a
// multiple of the 128-column GEMM tile) so that no launch ends in a nearly empty wave.
static long long host_chunk(const Handle* h, long long B, long long p) {
    const long long cap = std::max<long long>(128, (long long)(4.0e9 / (8.0 * (double)h->gs->K)) / 128 * 128);
    B = std::min(B, cap);
    if (p <= B) return B;
    const long long nchunks = (p + B - 1) / B;
    return std::min(B, round_up((p + nchunks - 1) / nchunks, 128));
}

static int reserve_scan(Handle* h, long long B, bool interaction) {
    const int R = h->R, mp = h->mp, m = h->m, k = h->k0;
    if (interaction) {
        CRM_CHECK(h->C.reserve((size_t)B * h->kexp * h->ldH * 8));
        CRM_CHECK(h->sq.reserve((size_t)B * h->ld2 * 8));
        CRM_CHECK(h->gr.reserve((size_t)B * R * mp * 8));
        CRM_CHECK(h->Vg.reserve((size_t)m * round_up(B * k, 2) * 8));
        CRM_CHECK(h->GEr.reserve((size_t)B * k * mp * 8));
        CRM_CHECK(h->lam.reserve((size_t)B * k * 8));
    } else {
        CRM_CHECK(h->C.reserve((size_t)B * h->ldH * 8));
        CRM_CHECK(h->sq.reserve((size_t)B * 2 * 8));
        CRM_CHECK(h->gr.reserve((size_t)B * mp * 8));
    }
    CRM_CHECK(h->Hg.reserve((size_t)m * round_up(B, 2) * 8));
    const size_t pr = (size_t)B * R;
    CRM_CHECK(h->fit_lml.reserve(pr * 8)); CRM_CHECK(h->fit_delta.reserve(pr * 8)); CRM_CHECK(h->fit_scale.reserve(pr * 8));
    CRM_CHECK(h->fit_beta.reserve(pr * (size_t)(h->c + 2) * 8)); CRM_CHECK(h->fit_x.reserve(pr * 8));
    CRM_CHECK(h->fit_nfev.reserve(pr * 4)); CRM_CHECK(h->fit_flags.reserve(pr * 4));
    CRM_CHECK(h->rho_idx.reserve((size_t)B * 4)); CRM_CHECK(h->best_lml.reserve((size_t)B * 8));
    CRM_CHECK(h->v0.reserve((size_t)B * 8)); CRM_CHECK(h->v1.reserve((size_t)B * 8));
    CRM_CHECK(h->perm.reserve((size_t)B * 4)); CRM_CHECK(h->offsets.reserve((size_t)(R + 1) * 4));
    CRM_CHECK(h->Q.reserve((size_t)B * 8)); CRM_CHECK(h->nlam.reserve((size_t)B * 4)); CRM_CHECK(h->sflags.reserve((size_t)B * 4));
    CRM_CHECK(h->liu.reserve((size_t)B * 8)); CRM_CHECK(h->ifault.reserve((size_t)B * 4)); CRM_CHECK(h->conv.reserve((size_t)B * 4));
    CRM_CHECK(h->scratch.reserve((size_t)(h->rho.size() + 16) * 8));
    return CRM_OK;
}

static int ensure_streams(Handle* h) {
    if (!h->copy_stream) {
        CRM_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CRM_CUDA(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
            CRM_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
    }
    return CRM_OK;
}

// Stage the column block [s0, s0+B) of a host matrix (and of the optional second one) into the device chunk buffers of
// `slot` (ld = Bp) on the copy stream.
static int stage_host_block(Handle* h, const double* G, long long ldg, const double* G2, long long ldg2, long long s0, long long B,
                            long long Bp, int slot) {
    const size_t rows = (size_t)h->gs->K;
    CRM_CHECK(h->gchunk[slot].reserve(rows * Bp * 8));
    if (G2) CRM_CHECK(h->gtchunk[slot].reserve(rows * Bp * 8));
    CRM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[slot], 0));   // previous consumer of this slot
    CRM_CUDA(cudaMemcpy2DAsync(h->gchunk[slot].ptr, (size_t)Bp * 8, G + s0, (size_t)ldg * 8, (size_t)B * 8, rows, cudaMemcpyHostToDevice, h->copy_stream));
    if (G2) CRM_CUDA(cudaMemcpy2DAsync(h->gtchunk[slot].ptr, (size_t)Bp * 8, G2 + s0, (size_t)ldg2 * 8, (size_t)B * 8, rows, cudaMemcpyHostToDevice, h->copy_stream));
    CRM_CUDA(cudaEventRecord(h->ev_copy[slot], h->copy_stream));
    return CRM_OK;
}

// Width cap of the feeder's column blocks (whole 256-SNP tiles of the int8 contraction, at most 3072 columns; 2816 with 16 threads).  Narrow
// enough that the first block is converted while the set-up runs (25-30 ms for 100k x 2816 float64 on 16 threads), wide enough that the
// per-block costs of the scan (launch chain, host synchronisation for the rho groups, the small g^2 contraction) stay a few per cent;
// measured at bench size with equal blocks: 1792 -> 243 ms, 2560 -> 225 ms, 5120 -> 257 ms per call (profiles/e2e_blocks_ab.py).
static long long feeder_block_cap() {
    if (const char* v = getenv("CRM_FEEDER_CAP")) { if (atoll(v) >= 256) return std::min<long long>(3072, atoll(v) / 256 * 256); }      // tests: cap of another thread count
    // few host threads (several ranks sharing the cores of a box): narrower blocks, so that a block is converted in ~25 ms whatever the
    // thread count (5 GB/s of float64 per thread) and the scan of block i hides the conversion of block i + 1
    return std::min<long long>(3072, std::max<long long>(512, round_up(176LL * host_threads(), 256)));
}
// Column blocks [starts[b], starts[b + 1]) of the feeder.  Equal blocks under the cap by default.  With the width of the basis operand
// known (hint of the Python layer) and a wide cap, the widths are chosen in units of the 256-SNP tiles of the int8 contraction so that
// its persistent grid ends on full waves: a block of t tiles is ceil(m_tiles t / SMs) waves, and four equal blocks of 10 tiles cost 28
// waves at cfg3 where {11, 11, 11, 7} cost 26 (25.4 of work).  Wider blocks come first: the first one is converted under the set-up.
static std::vector<long long> feeder_block_starts(long long p, long long basis_cols, int device) {
    std::vector<long long> starts;
    if (const char* env = getenv("CRM_FEEDER_BLOCK")) {      // tests: many small blocks
        if (atoll(env) > 0) { for (long long s0 = 0; s0 < p; s0 += atoll(env)) starts.push_back(s0); starts.push_back(p); return starts; }
    }
    const long long cap = feeder_block_cap(), T = (p + 255) / 256, capT = cap / 256;
    static int sms[32] = {0};
    if (device >= 0 && device < 32 && !sms[device]) { if (cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); sms[device] = 148; } }
    const long long nsm = (device >= 0 && device < 32) ? sms[device] : 148;
    const long long m_tiles = basis_cols > 0 ? (basis_cols + 127) / 128 : 0;
    if (m_tiles > 0 && capT >= 8 && T > capT && T <= 100000) {
        // best[t]: least cost (waves + 0.7 per block for its launch chain and host synchronisation) of covering t tiles
        std::vector<double> best(T + 1, 1e300); std::vector<int> pick(T + 1, 0);
        best[0] = 0.0;
        for (long long t = 1; t <= T; t++)
            for (long long w = 1; w <= std::min(capT, t); w++) {
                const double c = best[t - w] + (double)((m_tiles * w + nsm - 1) / nsm) + 0.7;
                if (c < best[t] - 1e-9) { best[t] = c; pick[t] = (int)w; }
            }
        std::vector<long long> widths;
        for (long long t = T; t > 0; t -= pick[t]) widths.push_back(pick[t]);
        std::sort(widths.begin(), widths.end(), [](long long a, long long b) { return a > b; });
        long long s0 = 0;
        for (long long w : widths) { starts.push_back(s0); s0 += w * 256; }
        starts.push_back(p);
        return starts;
    }
    const long long nblocks = (p + cap - 1) / cap;
    const long long block = std::min(p, round_up((p + nblocks - 1) / nblocks, 256));
    for (long long s0 = 0; s0 < p; s0 += block) starts.push_back(s0);
    starts.push_back(p);
    return starts;
}

static void drop_feed(Handle* h) {
    if (h->feed) { feeder_cancel(*h->feed); h->feed.reset(); }
    h->feed_src = nullptr;
}

// Starts the conversion of a host genotype matrix into int8 blocks (feeder.hpp); returns at once.
static int start_feed(Handle* h, const void* G, int dtype, long long ldg, long long rows, long long p, long long basis_cols) {
    drop_feed(h);
    static const bool off = [] { const char* v = getenv("CRM_NO_FEEDER"); return v && atoi(v) != 0; }();
    if (off || host_dtype_size(dtype) == 0) return CRM_OK;
    auto job = std::make_shared<FeedJob>();
    job->src = G; job->dtype = dtype; job->ld = ldg; job->rows = rows; job->p = p;
    job->starts = feeder_block_starts(p, basis_cols, h->device);
    long long block = 0;
    for (size_t b = 0; b + 1 < job->starts.size(); b++) block = std::max(block, job->starts[b + 1] - job->starts[b]);
    job->slot_ld = round_up(block, 16);
    job->nslots = (int)std::min<long long>(3, job->nblocks());
    for (int i = 0; i < job->nslots; i++) {
        job->slots[i] = pinned_slot(i, (size_t)rows * job->slot_ld);
        if (!job->slots[i]) return CRM_OK;                  // cannot page-lock that much: the scan streams the matrix as float64 instead
    }
    feeder_submit(job);
    h->feed = job; h->feed_src = G; h->feed_ld = ldg; h->feed_p = p; h->feed_rows = rows; h->feed_dtype = dtype;
    if (trace_on()) fprintf(stderr, "[crm trace] host %.1f ms: feeder started, %lld x %lld genotypes (dtype %d) in %lld blocks of up to %lld columns, %d threads\n", host_ms(), rows, p, dtype,
                            job->nblocks(), block, host_threads());
    return CRM_OK;
}

// Start moving a host genotype matrix (rows x p, leading dimension ldg) to the device ahead of the scan; returns at once.  Pinned float64:
// plain DMA in column chunks on the copy stream.  Anything else (pageable float64 -- what a reference user passes -- or integer / float32
// storage): conversion to int8 dosage blocks by the host feeder.  The next scan of the same matrix (same pointer, ldg, p) consumes it
// block by block as the data arrive -- called before crm_setup, the transfer overlaps the set-up and the first blocks of the scan.
static int do_stage_genotypes(Handle* h, const void* Gv, int dtype, long long ldg, long long rows, long long p, long long basis_cols, cudaStream_t st) {
    h->stage_valid = false;
    drop_feed(h);
    if (!Gv || rows <= 0 || p <= 0 || ldg < p) { set_error("crm_stage_genotypes: bad arguments"); return CRM_ERR_INVALID; }
    static size_t total_mem[16] = {0};
    if (!total_mem[h->device]) { size_t f = 0; CRM_CUDA(cudaMemGetInfo(&f, &total_mem[h->device])); }
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, Gv) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    static const bool force_feeder = [] { const char* v = getenv("CRM_FEEDER"); return v && atoi(v) != 0; }();
    if (dtype != HD_F64 || !pinned || force_feeder) return start_feed(h, Gv, dtype, ldg, rows, p, basis_cols);
    const double* G = static_cast<const double*>(Gv);
    const long long ld = round_up(p, 2);
    const double bytes = (double)rows * (double)ld * 8.0;
    if (bytes > 0.25 * (double)total_mem[h->device]) return CRM_OK;          // too large to hold: the scan streams it in blocks instead
    CRM_CHECK(ensure_streams(h));
    CRM_CHECK(h->gstage.reserve((size_t)bytes));
    // One copy stream, chunk after chunk: the copy engine drains one stream's queue before it turns to the next, so chunks spread
    // over several streams would complete stream by stream instead of in column order (measured).  Strided (2-D) copies of this
    // shape run at the full PCIe rate (profiles/h2d_probe.py: 55 GB/s for any chunk width from 512 columns).
    const int chunk = 512;
    const size_t nchunks = (size_t)((p + chunk - 1) / chunk);
    while (h->stage_events.size() < nchunks) { cudaEvent_t e; CRM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->stage_events.push_back(e); }
    // the buffer comes from the stream-ordered pool on `st`; the copy stream is non-blocking: order it explicitly
    CRM_CUDA(cudaEventRecord(h->ev_done[1], st));
    CRM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[1], 0));
    for (size_t c = 0; c < nchunks; c++) {
        const long long c0 = (long long)c * chunk, w = std::min<long long>(chunk, p - c0);
        CRM_CUDA(cudaMemcpy2DAsync(h->gstage.as<double>() + c0, (size_t)ld * 8, G + c0, (size_t)ldg * 8, (size_t)w * 8, (size_t)rows, cudaMemcpyHostToDevice, h->copy_stream));
        CRM_CUDA(cudaEventRecord(h->stage_events[c], h->copy_stream));
    }
    if (trace_on()) fprintf(stderr, "[crm trace] host %.1f ms: staging %lld x %lld genotypes in %zu chunks of %d columns\n", host_ms(), rows, p, nchunks, chunk);
    h->stage_src = G; h->stage_ldg = ldg; h->stage_p = p; h->stage_rows = rows; h->stage_ld = ld; h->stage_chunk = chunk;
    h->stage_valid = true;
    return CRM_OK;
}

// Columns [s0, s0 + b) of a host matrix of any element type as a float64 device block (gchunk[0], leading dimension round_up(b, 2)),
// through two pinned bounce buffers filled by the worker pool: the route of blocks that are not integer dosages.
static int bounce_block_f64(Handle* h, const GSource& src, long long s0, long long b, cudaStream_t st, GBlock* blk) {
    const long long rows = h->gs->K, ldd = round_up(b, 2);
    CRM_CHECK(ensure_streams(h));
    CRM_CHECK(h->gchunk[0].reserve((size_t)rows * ldd * 8));
    const size_t slot_bytes = std::max<size_t>((size_t)128 << 20, (size_t)rows * 16 * 8);
    const long long sub = std::max<long long>(16, std::min<long long>(b, (long long)(slot_bytes / ((size_t)rows * 8))));
    int it = 0;
    for (long long c0 = 0; c0 < b; c0 += sub, it++) {
        const long long w = std::min(sub, b - c0);
        const int slot = it & 1;
        double* bounce = reinterpret_cast<double*>(pinned_slot(4 + slot, slot_bytes));
        if (!bounce) { set_error("cannot page-lock a %zu-byte bounce buffer", slot_bytes); return CRM_ERR_CUDA; }
        if (it >= 2) CRM_CUDA(cudaEventSynchronize(h->ev_copy[slot]));          // the copy that last read this buffer
        host_parallel_widen(src.ptr, src.dtype, src.ld, rows, s0 + c0, w, bounce, w);
        CRM_CUDA(cudaMemcpy2DAsync(h->gchunk[0].as<double>() + c0, (size_t)ldd * 8, bounce, (size_t)w * 8, (size_t)w * 8, (size_t)rows, cudaMemcpyHostToDevice, st));
        CRM_CUDA(cudaEventRecord(h->ev_copy[slot], st));
    }
    CRM_CUDA(cudaStreamSynchronize(st));          // the bounce buffers are process-wide
    *blk = GBlock{h->gchunk[0].as<double>(), ldd, b, nullptr, 0, b, s0, nullptr, 0, -1};
    return CRM_OK;
}

// Walks the columns of G (and of the optional second matrix G2, same type and placement) in blocks of at most B columns.
//  * device float64: used in place when TMA can address it (16-byte aligned base, even leading dimension, even first column), otherwise
//    the block is repacked; device int8: handed over as an int8 block.
//  * host float64 staged ahead by DMA (pinned, crm_stage_genotypes): blocks wait for the chunks that cover them.
//  * other host matrices: int8 blocks from the feeder (started by crm_stage_genotypes, or here), copied to the device on the copy stream
//    while earlier blocks are scanned; a block that is not integer-valued stops the feeder and the rest goes over as float64.
//  * CRM_NO_FEEDER=1, or two host matrices (permuted tested genotypes): float64 blocks streamed through double-buffered chunks.
template <class F>
static int for_each_block(Handle* h, const GSource& src, const GSource* src2, long long p, long long B, cudaStream_t st, F&& fn) {
    const long long rows = h->gs->K;
    if (src2 && (src2->dtype != src.dtype || src2->on_host != src.on_host)) { set_error("the two genotype matrices must share type and placement"); return CRM_ERR_INVALID; }
    if (!src.on_host) {
        if (src.dtype == HD_I8) {
            const int8_t* G8 = static_cast<const int8_t*>(src.ptr);
            for (long long s0 = 0; s0 < p; s0 += B) {
                const long long b = std::min(B, p - s0);
                GBlock blk{nullptr, 0, 0, nullptr, 0, b, s0, G8 + s0, src.ld, -1};
                if (src2) {
                    const long long ld = round_up(b, 2);
                    CRM_CHECK(h->gwide2.reserve((size_t)rows * ld * 8));
                    CRM_CHECK(oz_launch_widen_i8(static_cast<const int8_t*>(src2->ptr) + s0, src2->ld, rows, b, h->gwide2.as<double>(), ld, st));
                    blk.G2 = h->gwide2.as<double>(); blk.ld2 = ld;
                }
                CRM_CHECK(fn(blk));
            }
            return CRM_OK;
        }
        if (src.dtype != HD_F64) { set_error("device genotypes must be float64 or int8"); return CRM_ERR_UNSUPPORTED; }
        const double* G = static_cast<const double*>(src.ptr); const long long ldg = src.ld;
        const double* G2 = src2 ? static_cast<const double*>(src2->ptr) : nullptr; const long long ldg2 = src2 ? src2->ld : 0;
        const bool aligned = ((reinterpret_cast<uintptr_t>(G) & 15) == 0) && (ldg % 2 == 0);
        const bool aligned2 = !G2 || (((reinterpret_cast<uintptr_t>(G2) & 15) == 0) && (ldg2 % 2 == 0));
        for (long long s0 = 0; s0 < p; s0 += B) {
            const long long b = std::min(B, p - s0), bp = round_up(b, 2);
            GBlock blk{G + s0, ldg, p - s0, G2 ? G2 + s0 : nullptr, ldg2, b, s0, nullptr, 0, -1};
            if (!aligned || !aligned2 || (s0 & 1)) {
                CRM_CHECK(h->gchunk[0].reserve((size_t)rows * bp * 8));
                CRM_CUDA(cudaMemcpy2DAsync(h->gchunk[0].ptr, (size_t)bp * 8, G + s0, (size_t)ldg * 8, (size_t)b * 8, (size_t)rows, cudaMemcpyDeviceToDevice, st));
                blk.G = h->gchunk[0].as<double>(); blk.ld = bp; blk.cols = b;
                if (G2) {
                    CRM_CHECK(h->gtchunk[0].reserve((size_t)rows * bp * 8));
                    CRM_CUDA(cudaMemcpy2DAsync(h->gtchunk[0].ptr, (size_t)bp * 8, G2 + s0, (size_t)ldg2 * 8, (size_t)b * 8, (size_t)rows, cudaMemcpyDeviceToDevice, st));
                    blk.G2 = h->gtchunk[0].as<double>(); blk.ld2 = bp;
                }
            }
            CRM_CHECK(fn(blk));
        }
        return CRM_OK;
    }
    CRM_CHECK(ensure_streams(h));
    // ---- host float64 staged ahead by DMA ----
    if (src.dtype == HD_F64 && h->stage_valid && !src2 && src.ptr == h->stage_src && src.ld == h->stage_ldg && p == h->stage_p && rows == h->stage_rows) {
        for (long long s0 = 0; s0 < p; s0 += B) {
            const long long b = std::min(B, p - s0);
            const long long c_last = (s0 + b - 1) / h->stage_chunk;                       // chunks complete in order on the copy stream
            CRM_CUDA(cudaStreamWaitEvent(st, h->stage_events[(size_t)c_last], 0));
            GBlock blk{h->gstage.as<double>() + s0, h->stage_ld, h->stage_p - s0, nullptr, 0, b, s0, nullptr, 0, -1};
            CRM_CHECK(fn(blk));
        }
        h->stage_valid = false;         // one scan per staging: the host array may change afterwards
        return CRM_OK;
    }
    // ---- int8 blocks from the feeder ----
    long long done = 0;                 // columns handled so far
    if (!src2) {
        if (!(h->feed && h->feed_src == src.ptr && h->feed_ld == src.ld && h->feed_p == p && h->feed_rows == rows && h->feed_dtype == src.dtype))
            CRM_CHECK(start_feed(h, src.ptr, src.dtype, src.ld, rows, p, h->gs == &h->cells ? plane_rows(h) : 0));
        if (h->feed) {
            std::shared_ptr<FeedJob> job = h->feed;
            const long long nb = job->nblocks();
            const size_t slot_bytes = (size_t)rows * job->slot_ld;
            int status = CRM_OK;
            // landing buffers first: they are allocated in the order of `st`, and the events below carry that order to the copy stream
            for (int d = 0; d < std::min<long long>(2, nb); d++) CRM_CHECK(h->g8dev[d].reserve(slot_bytes));
            CRM_CUDA(cudaEventRecord(h->ev_done[0], st));
            CRM_CUDA(cudaEventRecord(h->ev_done[1], st));
            for (long long ib = 0; ib < nb && status == CRM_OK; ib++) {
                int bad = 0, gmax = 0;
                const double t_wait = host_ms();
                feeder_wait_block(*job, ib, &bad, &gmax);
                if (trace_on()) fprintf(stderr, "[crm trace] host %.1f ms: feeder block %lld of %lld ready after waiting %.1f ms (bad=%d, max |g|=%d)\n", host_ms(), ib, nb, host_ms() - t_wait, bad, gmax);
                if (bad) break;                                                      // not integer dosages: the rest goes over as float64
                const int dslot = (int)(ib & 1);
                const long long s0 = job->starts[ib], bw = job->starts[ib + 1] - s0;
                cudaError_t ce = cudaStreamWaitEvent(h->copy_stream, h->ev_done[dslot], 0);          // scan of the block that used this buffer before
                const long long ld_b = job->block_ld(ib);                           // blocks are packed with their own leading dimension: one contiguous copy
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->g8dev[dslot].ptr, job->slots[ib % job->nslots], (size_t)rows * ld_b, cudaMemcpyHostToDevice, h->copy_stream);
                if (ce == cudaSuccess) ce = cudaEventRecord(h->ev_copy[dslot], h->copy_stream);
                if (ce == cudaSuccess) ce = feeder_release_after(job, ib, h->copy_stream);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, h->ev_copy[dslot], 0);
                if (ce != cudaSuccess) { set_error("feeder copy failed: %s", cudaGetErrorString(ce)); status = CRM_ERR_CUDA; break; }
                for (long long c0 = 0; c0 < bw && status == CRM_OK; c0 += B) {       // scan batches inside the block
                    const long long b = std::min(B, bw - c0);
                    GBlock blk{nullptr, 0, 0, nullptr, 0, b, s0 + c0, h->g8dev[dslot].as<int8_t>() + c0, ld_b, gmax};
                    status = fn(blk);
                }
                if (status == CRM_OK && cudaEventRecord(h->ev_done[dslot], st) != cudaSuccess) { set_error("cudaEventRecord failed"); status = CRM_ERR_CUDA; }
                done = s0 + bw;
            }
            drop_feed(h);               // one scan per conversion (and the workers must not outlive the caller's array)
            CRM_CHECK(status);
            if (done >= p) return CRM_OK;
            static const bool nofeed = [] { const char* v = getenv("CRM_NO_FEEDER"); return v && atoi(v) != 0; }();
            if (!nofeed) {
                for (long long s0 = done; s0 < p; s0 += B) {
                    const long long b = std::min(B, p - s0);
                    GBlock blk{};
                    CRM_CHECK(bounce_block_f64(h, src, s0, b, st, &blk));
                    CRM_CHECK(fn(blk));
                }
                return CRM_OK;
            }
        }
    }
    // ---- float64 blocks streamed through double-buffered chunks ----
    if (src.dtype != HD_F64) { set_error("host genotypes of element type %d need the feeder (CRM_NO_FEEDER is set) or float64 storage", src.dtype); return CRM_ERR_UNSUPPORTED; }
    const double* G = static_cast<const double*>(src.ptr); const long long ldg = src.ld;
    const double* G2 = src2 ? static_cast<const double*>(src2->ptr) : nullptr; const long long ldg2 = src2 ? src2->ld : 0;
    const long long Bp = round_up(B, 2);
    std::vector<long long> starts;
    {   // a short first block: its copy is the only one that is not hidden behind compute
        const long long first = (p > B) ? std::min<long long>(B, 512) : std::min(B, p);
        starts.push_back(0);
        for (long long s0 = first; s0 < p; s0 += B) starts.push_back(s0);
        starts.push_back(p);
    }
    const long long nb = (long long)starts.size() - 1;
    for (int d = 0; d < std::min<long long>(2, nb); d++) {       // allocated in the order of `st`; the events below carry it to the copy stream
        CRM_CHECK(h->gchunk[d].reserve((size_t)rows * Bp * 8));
        if (G2) CRM_CHECK(h->gtchunk[d].reserve((size_t)rows * Bp * 8));
    }
    CRM_CUDA(cudaEventRecord(h->ev_done[0], st));
    CRM_CUDA(cudaEventRecord(h->ev_done[1], st));
    CRM_CHECK(stage_host_block(h, G, ldg, G2, ldg2, starts[0], starts[1] - starts[0], Bp, 0));
    for (long long ib = 0; ib < nb; ib++) {
        const int slot = (int)(ib & 1);
        const long long s0 = starts[ib], b = starts[ib + 1] - s0;
        if (ib + 1 < nb) CRM_CHECK(stage_host_block(h, G, ldg, G2, ldg2, starts[ib + 1], starts[ib + 2] - starts[ib + 1], Bp, slot ^ 1));
        CRM_CUDA(cudaStreamWaitEvent(st, h->ev_copy[slot], 0));
        GBlock blk{h->gchunk[slot].as<double>(), Bp, b, G2 ? h->gtchunk[slot].as<double>() : nullptr, Bp, b, s0, nullptr, 0, -1};
        CRM_CHECK(fn(blk));
        CRM_CUDA(cudaEventRecord(h->ev_done[slot], st));
    }
    return CRM_OK;
}

static int interaction_batch(Handle* h, GBlock blk, double* out_pv, double* out_rho1, double* out_e2, double* out_g2, double* out_eps2,
                             const crm_scan_diag_t* dg, cudaStream_t st) {
    const int R = h->R, mp = h->mp, m = h->m, k = h->k0, kexp = h->kexp, c = h->c, ldH = h->ldH, Mx = h->Mx;
    const long long B = blk.b, s0 = blk.s0;
    const double* Gt = blk.G2; const long long ldgt = blk.ld2;
    if (Gt) CRM_CHECK(block_f64(h, blk, st));        // permuted tested genotypes: both designs are contracted in float64
    double* C = h->C.as<double>();
    double* sq = h->sq.as<double>();
    PhaseTrace tr(st);
    // 1. rotation of [g, g.E0] onto [H | y | W]
    if (h->prof_on) {
        cudaEvent_t e0, e1;
        CRM_CUDA(cudaEventCreate(&e0)); CRM_CUDA(cudaEventCreate(&e1));
        CRM_CUDA(cudaEventRecord(e0, st));
        h->prof_events.push_back(e0); h->prof_events.push_back(e1);
        h->prof_flops += 2.0 * (double)h->n * (double)h->m * (double)kexp * (double)B;   // algorithmic: 2 n m (1+k) per SNP
    }
    if (Gt) {
        GBlock tested{Gt, ldgt, blk.cols, nullptr, 0, B, s0, nullptr, 0, -1};
        CRM_CHECK(launch_rotation(h, tested, C, st));
    } else {
        CRM_CHECK(launch_rotation(h, blk, C, st));
    }
    if (h->prof_on) CRM_CUDA(cudaEventRecord(h->prof_events.back(), st));
    tr.mark("rotation");
    // 2. squared-genotype Grams against [1 | E0 | pairs]:  g'g, (g.E0)'g, (g.E0)'(g.E0)
    if (!Gt && h->gs == &h->cells && h->oz_block_valid && h->oz_block_gmax <= 11 &&
        (double)round_up(h->n, 16) * 64.0 * (double)(h->oz_block_gmax * h->oz_block_gmax) < 2147483648.0) {   // int32 accumulation of q * g^2
        // integer dosages: g^2 is an exact int8 operand as well -> same int8 split with the digit planes of A2
        const long long n = h->n, M2p = round_up(h->M2, 16), Kp = round_up(n, 16), Bp = round_up(B, 16);
        if (h->A28.cap == 0 || !h->oz_built_a2) {
            CRM_CHECK(h->A28.reserve((size_t)OZAKI_SLICES * M2p * Kp));
            CRM_CHECK(h->a28expo.reserve((size_t)h->M2 * sizeof(int)));
            CRM_CHECK(oz_launch_matrix_planes(h->A2.as<double>(), h->ld2, h->M2, n, h->a28expo.as<int>(), h->A28.as<int8_t>(), M2p, Kp, st));
            h->oz_built_a2 = true;
        }
        CRM_CHECK(int8_split_contract(h, h->A28.as<int8_t>(), M2p, h->M2, h->a28expo.as<int>(), h->G2t8.as<int8_t>(), Bp, B, Kp, sq, h->ld2, st));
        if (h->oz_block_affine) {       // g = a d + b: Grams of g^2 from those of d^2 (above) and of d
            CRM_CHECK(h->sq1.reserve((size_t)B * h->ld2 * 8));
            CRM_CHECK(int8_split_contract(h, h->A28.as<int8_t>(), M2p, h->M2, h->a28expo.as<int>(), h->Gt8.as<int8_t>(), Bp, B, Kp, h->sq1.as<double>(), h->ld2, st));
            CRM_CHECK(oz_launch_affine_fix_square(sq, h->sq1.as<double>(), h->ld2, B, h->M2, h->aff.as<double>(), round_up(B, 2), h->colsum2.as<double>(), st));
        }
    } else {
        CRM_CHECK(block_f64(h, blk, st));
        GemmOperands op{};
        op.A = h->gs->A2; op.lda = h->gs->ld2; op.a_cols = h->M2;
        op.B = Gt ? Gt : blk.G; op.ldb = Gt ? ldgt : blk.ld; op.b_cols = blk.cols;
        op.B2 = op.B; op.ldb2 = op.ldb; op.b2_cols = blk.cols;
        CRM_CHECK(launch_gemm(GEMM_PRODUCT, op, (int)h->gs->K, 0, h->M2, 0, (int)B, sq, h->ld2, 1, st));
    }
    const double* Gd = blk.G; const long long ldg = blk.ld, gcols = blk.cols;
    if (Gt) {
        // permuted tested genotypes (idx_G, reference :410-413): the null design still uses g itself, so the j = 0 rows of C
        // are overwritten with the rotation of g, column 0 of sq with g'g and columns 1..k with (gt * g)' E0
        GemmOperands op{};
        op.A = h->gs->Hx; op.lda = h->gs->ldHx; op.a_cols = Mx;
        op.B = Gd; op.ldb = ldg; op.b_cols = gcols; op.B2 = Gd; op.ldb2 = ldg; op.b2_cols = gcols;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->gs->K, 0, Mx, 0, (int)B, C, (long long)kexp * ldH, 1, st));
        GemmOperands o2{};
        o2.A = h->gs->A2; o2.lda = h->gs->ld2; o2.a_cols = h->M2;
        o2.B = Gd; o2.ldb = ldg; o2.b_cols = gcols; o2.B2 = Gd; o2.ldb2 = ldg; o2.b2_cols = gcols;
        CRM_CHECK(launch_gemm(GEMM_PRODUCT, o2, (int)h->gs->K, 0, 1, 0, (int)B, sq, h->ld2, 1, st));
        o2.B2 = Gt; o2.ldb2 = ldgt;
        CRM_CHECK(launch_gemm(GEMM_PRODUCT, o2, (int)h->gs->K, 1, k, 0, (int)B, sq + 1, h->ld2, 1, st));
    }
    tr.mark("g2 grams");
    // 3. H'g as a K-outer operand, 4. rotated genotype for every rho
    const long long ldhg = round_up(B, 2);
    CRM_CHECK(launch_gather_transpose(C, ldH, nullptr, kexp, 0, 1, B, m, h->Hg.as<double>(), ldhg, st));
    {
        GemmOperands op{};
        op.A = h->Tt.as<double>(); op.lda = (long long)R * mp; op.a_cols = (long long)R * mp;
        op.B = h->Hg.as<double>(); op.ldb = ldhg; op.b_cols = B;
        op.B2 = op.B; op.ldb2 = ldhg; op.b2_cols = B;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, m, 0, R * mp, 0, (int)B, h->gr.as<double>(), (long long)R * mp, 1, st));
    }
    tr.mark("transform g");
    // 5. REML fits for every (SNP, rho)
    FitArgs fa{};
    fa.S = h->S.as<double>(); fa.yr = h->yr.as<double>(); fa.Wr = h->Wr.as<double>();
    fa.gr = h->gr.as<double>(); fa.gr_ld = (long long)R * mp;
    fa.gy = C + m; fa.gy_ld = (long long)kexp * ldH;
    fa.gW = C + m + 1; fa.gW_ld = (long long)kexp * ldH;
    fa.gg = sq; fa.gg_ld = h->ld2;
    fa.stats = h->stats.as<double>();
    fa.m = m; fa.mp = mp; fa.R = R; fa.c = c; fa.p = (int)B; fa.n = (double)h->n; fa.restricted = 1; fa.fixed_x = nullptr;
    fa.lml = h->fit_lml.as<double>(); fa.delta = h->fit_delta.as<double>(); fa.scale = h->fit_scale.as<double>();
    fa.beta = h->fit_beta.as<double>(); fa.xopt = h->fit_x.as<double>(); fa.nfev = h->fit_nfev.as<int>(); fa.flags = h->fit_flags.as<int>();
    if (c + 1 > 8) {    // designs wider than the register-resident K2 kernel: shared-memory kernel K5, per-rho spectra
        BetaArgs wa{};
        wa.S = fa.S; wa.S_stride = mp; wa.Zs = h->YW.as<double>(); wa.Zs_stride = (long long)(1 + c) * mp;
        wa.Zp = h->gr.as<double>(); wa.Zp_snp_stride = (long long)R * mp; wa.Zp_rho_stride = mp;
        wa.shared_gram = h->ywgram.as<double>(); wa.rot = C; wa.rot_ld = ldH; wa.col_y = m; wa.col_W = m + 1; wa.kexp = kexp;
        wa.lin = nullptr; wa.lin_ld = 0; wa.sq = sq; wa.sq_ld = h->ld2; wa.rho = nullptr;
        wa.m = m; wa.mp = mp; wa.c = c; wa.k0 = 0; wa.R = R; wa.p = (int)B; wa.has_g = 1; wa.mix_rho = 0; wa.restricted = 1; wa.fixed_x = nullptr; wa.n = (double)h->n;
        wa.lml = fa.lml; wa.delta = fa.delta; wa.scale = fa.scale; wa.beta = fa.beta; wa.ucoef = nullptr; wa.xopt = fa.xopt; wa.nfev = fa.nfev; wa.flags = fa.flags;
        CRM_CHECK(launch_beta_fit(wa, st));
    } else {
        CRM_CHECK(attach_fit_table(h, fa, st));
        CRM_CHECK(launch_fit(fa, true, st));
    }
    tr.mark("fits");
    // 6. best rho per SNP, grouping by rho
    CRM_CHECK(launch_select(fa.lml, fa.delta, fa.scale, (int)B, R, h->rho_idx.as<int>(), h->best_lml.as<double>(), h->v0.as<double>(),
                            h->v1.as<double>(), st));
    if (dg && (dg->ov_rho_idx || dg->ov_v0 || dg->ov_v1)) {
        override_select_kernel<<<blocks_for(B, 256), 256, 0, st>>>(dg->ov_rho_idx ? dg->ov_rho_idx + s0 : nullptr, dg->ov_v0 ? dg->ov_v0 + s0 : nullptr,
                                                                  dg->ov_v1 ? dg->ov_v1 + s0 : nullptr, B, h->rho_idx.as<int>(), h->v0.as<double>(), h->v1.as<double>());
        CRM_CUDA(cudaGetLastError()); count_launch();
    }
    CRM_CHECK(launch_group(h->rho_idx.as<int>(), (int)B, R, h->perm.as<int>(), h->offsets.as<int>(), st));
    std::vector<int> off(R + 1);
    CRM_CUDA(cudaMemcpyAsync(off.data(), h->offsets.as<int>(), (size_t)(R + 1) * 4, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    tr.mark("select+group");
    // 7. H'(g.E0) in rho-sorted order as a K-outer operand, 8. rotation into the selected eigenbasis, group by group
    const long long ldvg = round_up(B * k, 2);
    CRM_CHECK(launch_gather_transpose(C, ldH, h->perm.as<int>(), kexp, 1, k, B * k, m, h->Vg.as<double>(), ldvg, st));
    for (int r = 0; r < R; r++) {
        const long long cnt = off[r + 1] - off[r];
        if (cnt <= 0) continue;
        GemmOperands op{};
        op.A = h->Tt.as<double>(); op.lda = (long long)R * mp; op.a_cols = (long long)R * mp;
        op.B = h->Vg.as<double>(); op.ldb = ldvg; op.b_cols = B * k;
        op.B2 = op.B; op.ldb2 = ldvg; op.b2_cols = B * k;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, m, r * mp, mp, (int)(off[r] * k), (int)(cnt * k), h->GEr.as<double>() + (long long)off[r] * k * mp, mp, 1, st));
    }
    tr.mark("transform gE");
    // 9. score statistic + eigenvalues
    ScoreArgs sa{};
    sa.S = fa.S; sa.yr = fa.yr; sa.Wr = fa.Wr; sa.m = m; sa.mp = mp; sa.R = R; sa.c = c; sa.k = k; sa.kexp = kexp; sa.p = (int)B;
    sa.perm = h->perm.as<int>(); sa.rho_idx = h->rho_idx.as<int>(); sa.v0 = h->v0.as<double>(); sa.v1 = h->v1.as<double>();
    sa.gr = h->gr.as<double>(); sa.gr_ld = (long long)R * mp; sa.GEr = h->GEr.as<double>();
    sa.rot = C; sa.rot_ld = ldH; sa.col_y = m; sa.col_W = m + 1;
    sa.sq = sq; sa.sq_ld = h->ld2; sa.stats = h->stats.as<double>();
    sa.Q = h->Q.as<double>(); sa.lam = h->lam.as<double>(); sa.lam_ld = k; sa.nlam = h->nlam.as<int>(); sa.flags = h->sflags.as<int>();
    sa.Mout = (dg && dg->M) ? dg->M + s0 * k * k : nullptr;
    CRM_CHECK(launch_score(sa, B, st));
    tr.mark("score");
    // 10. p-values
    PvalArgs pa{};
    pa.Q = sa.Q; pa.lam = sa.lam; pa.nlam = sa.nlam; pa.lam_ld = k; pa.count = (int)B; pa.lim = 10000; pa.acc = 1e-6;
    pa.pv = out_pv + s0; pa.liu = (dg && dg->liu) ? dg->liu + s0 : h->liu.as<double>();
    pa.ifault = (dg && dg->ifault) ? dg->ifault + s0 : h->ifault.as<int>(); pa.converged = h->conv.as<int>(); pa.trace = nullptr;
    CRM_CHECK(launch_pvalues(pa, st));
    tr.mark("p-values");
    // 11. outputs
    double* grid_dev = h->scratch.as<double>();
    CRM_CHECK(upload_small(grid_dev, h->rho.data(), R, st));
    finalize_interaction_kernel<<<blocks_for(B, 256), 256, 0, st>>>(h->rho_idx.as<int>(), h->v0.as<double>(), h->v1.as<double>(), grid_dev, B,
                                                                   out_rho1 + s0, out_e2 + s0, out_g2 + s0, out_eps2 + s0);
    CRM_CUDA(cudaGetLastError()); count_launch();
    if (dg) {
        const size_t pr = (size_t)B * R;
        if (dg->lml) CRM_CUDA(cudaMemcpyAsync(dg->lml + s0 * R, fa.lml, pr * 8, cudaMemcpyDeviceToDevice, st));
        if (dg->delta) CRM_CUDA(cudaMemcpyAsync(dg->delta + s0 * R, fa.delta, pr * 8, cudaMemcpyDeviceToDevice, st));
        if (dg->scale) CRM_CUDA(cudaMemcpyAsync(dg->scale + s0 * R, fa.scale, pr * 8, cudaMemcpyDeviceToDevice, st));
        if (dg->nfev) CRM_CUDA(cudaMemcpyAsync(dg->nfev + s0 * R, fa.nfev, pr * 4, cudaMemcpyDeviceToDevice, st));
        if (dg->Q) CRM_CUDA(cudaMemcpyAsync(dg->Q + s0, sa.Q, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
        if (dg->lam) CRM_CUDA(cudaMemcpyAsync(dg->lam + s0 * k, sa.lam, (size_t)B * k * 8, cudaMemcpyDeviceToDevice, st));
        if (dg->nlam) CRM_CUDA(cudaMemcpyAsync(dg->nlam + s0, sa.nlam, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
        if (dg->flags) {
            or_flags_kernel<<<blocks_for(B, 256), 256, 0, st>>>(fa.flags, h->rho_idx.as<int>(), sa.flags, R, B, dg->flags + s0);
            CRM_CUDA(cudaGetLastError()); count_launch();
        }
    }
    tr.mark("outputs");
    tr.report("interaction batch");
    return CRM_OK;
}

static int select_space(Handle* h, int donor_level, const char* who) {
    if (donor_level && !h->donors_set) { set_error("%s: donor-level genotypes need crm_set_donors first", who); return CRM_ERR_STATE; }
    h->gs = donor_level ? &h->donors : &h->cells;
    return CRM_OK;
}

static int do_scan_interaction(Handle* h, int donor_level, const GSource& G, long long p, const GSource* Gtest,
                               double* out_pv, double* out_rho1, double* out_e2, double* out_g2, double* out_eps2,
                               const crm_scan_diag_t* dg, cudaStream_t st) {
    if (!h->ready) { set_error("crm_scan_interaction: handle is not set up"); return CRM_ERR_STATE; }
    CRM_CHECK(select_space(h, donor_level, "crm_scan_interaction"));
    if (donor_level && Gtest) { set_error("crm_scan_interaction: permuted tested genotypes are not supported with donor-level input"); return CRM_ERR_UNSUPPORTED; }
    if (p == 0) return CRM_OK;
    if (!G.ptr || p < 0 || G.ld < p || !out_pv || !out_rho1 || !out_e2 || !out_g2 || !out_eps2) { set_error("crm_scan_interaction: bad arguments"); return CRM_ERR_INVALID; }
    long long B = pick_batch(h, p, true);
    if (G.on_host) B = host_chunk(h, B, p);
    CRM_CHECK(reserve_scan(h, B, true));
    return for_each_block(h, G, Gtest, p, B, st, [&](const GBlock& k) -> int {
        return interaction_batch(h, k, out_pv, out_rho1, out_e2, out_g2, out_eps2, dg, st);
    });
}

// ------------------------------------------------------------------------------------------------
// association scans
// ------------------------------------------------------------------------------------------------
static int do_scan_association(Handle* h, int donor_level, const GSource& G, long long p, int fast, double* out_pv,
                               double* out_alt, double* info4, double* out_null, cudaStream_t st) {
    if (!h->ready) { set_error("crm_scan_association: handle is not set up"); return CRM_ERR_STATE; }
    CRM_CHECK(select_space(h, donor_level, "crm_scan_association"));
    if ((p > 0 && (!G.ptr || G.ld < p || !out_pv)) || p < 0 || !info4) { set_error("crm_scan_association: bad arguments"); return CRM_ERR_INVALID; }
    const int R = h->R, mp = h->mp, m = h->m, c = h->c, ldH = h->ldH, Mx = h->Mx;
    long long B = pick_batch(h, std::max<long long>(p, 1), false);
    if (G.on_host) B = host_chunk(h, B, p);
    CRM_CHECK(reserve_scan(h, std::max<long long>(B, 1), false));
    // ---- null model: ML fit of y ~ W for every rho, best by strict '>' ----
    FitArgs fa{};
    fa.S = h->S.as<double>(); fa.yr = h->yr.as<double>(); fa.Wr = h->Wr.as<double>();
    fa.gr = nullptr; fa.stats = h->stats.as<double>();
    fa.m = m; fa.mp = mp; fa.R = R; fa.c = c; fa.p = 1; fa.n = (double)h->n; fa.restricted = 0; fa.fixed_x = nullptr;
    fa.lml = h->fit_lml.as<double>(); fa.delta = h->fit_delta.as<double>(); fa.scale = h->fit_scale.as<double>();
    fa.beta = h->fit_beta.as<double>(); fa.xopt = h->fit_x.as<double>(); fa.nfev = h->fit_nfev.as<int>(); fa.flags = h->fit_flags.as<int>();
    // designs wider than the register-resident K2 kernel go through the shared-memory kernel K5
    BetaArgs wa{};
    wa.S = h->S.as<double>(); wa.S_stride = mp; wa.Zs = h->YW.as<double>(); wa.Zs_stride = (long long)(1 + c) * mp; wa.Zp = nullptr;
    wa.Zp_snp_stride = 0; wa.Zp_rho_stride = 0; wa.shared_gram = h->ywgram.as<double>(); wa.rot = nullptr; wa.rot_ld = 0; wa.col_y = m; wa.col_W = m + 1;
    wa.kexp = 1; wa.lin = nullptr; wa.lin_ld = 0; wa.sq = nullptr; wa.sq_ld = 0; wa.rho = nullptr; wa.m = m; wa.mp = mp; wa.c = c; wa.k0 = 0; wa.R = R; wa.p = 1;
    wa.has_g = 0; wa.mix_rho = 0; wa.restricted = 0; wa.fixed_x = nullptr; wa.n = (double)h->n;
    wa.lml = fa.lml; wa.delta = fa.delta; wa.scale = fa.scale; wa.beta = fa.beta; wa.ucoef = nullptr; wa.xopt = fa.xopt; wa.nfev = fa.nfev; wa.flags = fa.flags;
    if (c > 7) CRM_CHECK(launch_beta_fit(wa, st)); else CRM_CHECK(launch_fit(fa, false, st));
    std::vector<double> lml(R), delta(R), scale(R), xopt(R);
    CRM_CUDA(cudaMemcpyAsync(lml.data(), fa.lml, R * 8, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaMemcpyAsync(delta.data(), fa.delta, R * 8, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaMemcpyAsync(scale.data(), fa.scale, R * 8, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaMemcpyAsync(xopt.data(), fa.xopt, R * 8, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    int rb = 0; double best = -INFINITY;
    for (int r = 0; r < R; r++) if (lml[r] > best) { best = lml[r]; rb = r; }
    const double v0 = scale[rb] * (1.0 - delta[rb]), v1 = scale[rb] * delta[rb], rho = h->rho[rb];
    const double info_host[4] = {rho, v0 * rho, v0 * (1.0 - rho), v1};
    CRM_CHECK(upload_small(info4, info_host, 4, st));
    if (out_null) CRM_CHECK(upload_small(out_null, &best, 1, st));
    double* xfix = h->scratch.as<double>() + R + 1;
    CRM_CHECK(upload_small(xfix, &xopt[rb], 1, st));
    CRM_CUDA(cudaStreamSynchronize(st));   // host temporaries above go out of scope
    if (p == 0) return CRM_OK;
    return for_each_block(h, G, nullptr, p, B, st, [&](const GBlock& kb) -> int {
        GBlock k = kb;
        CRM_CHECK(block_f64(h, k, st));
        CRM_CHECK(check_block_finite(h, k, st));
        const double* Gd = k.G; const long long ld = k.ld, cols = k.cols, b = k.b, s0 = k.s0;
        double* C = h->C.as<double>();
        double* sq = h->sq.as<double>();
        GemmOperands op{};
        op.A = h->gs->Hx; op.lda = h->gs->ldHx; op.a_cols = Mx;
        op.B = Gd; op.ldb = ld; op.b_cols = cols; op.B2 = Gd; op.ldb2 = ld; op.b2_cols = cols;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->gs->K, 0, Mx, 0, (int)b, C, ldH, 1, st));
        GemmOperands op2{};
        op2.A = h->gs->A2; op2.lda = h->gs->ld2; op2.a_cols = h->M2;
        op2.B = Gd; op2.ldb = ld; op2.b_cols = cols; op2.B2 = Gd; op2.ldb2 = ld; op2.b2_cols = cols;
        CRM_CHECK(launch_gemm(GEMM_PRODUCT, op2, (int)h->gs->K, 0, 1, 0, (int)b, sq, 2, 1, st));
        const long long ldhg = round_up(b, 2);
        CRM_CHECK(launch_gather_transpose(C, ldH, nullptr, 1, 0, 1, b, m, h->Hg.as<double>(), ldhg, st));
        GemmOperands op3{};
        op3.A = h->Tt.as<double>(); op3.lda = (long long)R * mp; op3.a_cols = (long long)R * mp;
        op3.B = h->Hg.as<double>(); op3.ldb = ldhg; op3.b_cols = b; op3.B2 = op3.B; op3.ldb2 = ldhg; op3.b2_cols = b;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op3, m, rb * mp, mp, 0, (int)b, h->gr.as<double>(), mp, 1, st));
        FitArgs fb = fa;
        fb.S = h->S.as<double>() + (long long)rb * mp; fb.yr = h->yr.as<double>() + (long long)rb * mp; fb.Wr = h->Wr.as<double>() + (long long)rb * c * mp;
        fb.gr = h->gr.as<double>(); fb.gr_ld = mp;
        fb.gy = C + m; fb.gy_ld = ldH; fb.gW = C + m + 1; fb.gW_ld = ldH; fb.gg = sq; fb.gg_ld = 2;
        fb.R = 1; fb.p = (int)b; fb.restricted = 0; fb.fixed_x = fast ? xfix : nullptr;
        fb.lml = out_alt ? out_alt + s0 : h->best_lml.as<double>();
        if (c + 1 > 8) {
            BetaArgs wb = wa;
            wb.S = fb.S; wb.S_stride = 0; wb.Zs = h->YW.as<double>() + (long long)rb * (1 + c) * mp; wb.Zs_stride = 0;
            wb.Zp = h->gr.as<double>(); wb.Zp_snp_stride = mp; wb.Zp_rho_stride = 0;
            wb.rot = C; wb.rot_ld = ldH; wb.sq = sq; wb.sq_ld = 2; wb.R = 1; wb.p = (int)b; wb.has_g = 1; wb.fixed_x = fb.fixed_x;
            wb.lml = fb.lml; wb.xopt = nullptr;
            CRM_CHECK(launch_beta_fit(wb, st));
        } else {
            CRM_CHECK(attach_fit_table(h, fb, st));
            CRM_CHECK(launch_fit(fb, true, st));
        }
        CRM_CHECK(launch_lrt(fb.lml, best, b, out_pv + s0, st));
        return (int)CRM_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// effect sizes (predict_interaction)
// ------------------------------------------------------------------------------------------------
static int do_predict(Handle* h, int donor_level, const GSource& G, long long p, const double* maf, int use_background,
                      double* out_beta_g, double* out_beta_gxe, long long ldo, double* out_rho1, cudaStream_t st) {
    if (!h->ready) { set_error("crm_predict_interaction: handle is not set up"); return CRM_ERR_STATE; }
    CRM_CHECK(select_space(h, donor_level, "crm_predict_interaction"));
    if (p == 0) return CRM_OK;
    if (!G.ptr || p < 0 || G.ld < p || !maf || !out_beta_g || !out_beta_gxe || ldo < p) { set_error("crm_predict_interaction: bad arguments"); return CRM_ERR_INVALID; }
    const int R = h->R, mp = h->mp, c = h->c, k0 = h->k0, kexp = h->kexp, ldH = h->ldH, Mx = h->Mx;
    const int ns = 1 + c + k0, P = c + 1 + k0;
    int r0 = -1;
    for (int r = 0; r < R; r++) if (h->rho[r] == 0.0) r0 = r;
    const int mB = (use_background && r0 >= 0 && h->mL > 0) ? h->m : 0;   // rows of the rotated background basis (0: no background)
    // ---- per-call shared quantities ----
    const int ldys = (int)round_up(ns, 2);
    CRM_CHECK(h->Ys.reserve((size_t)h->n * ldys * 8));
    CRM_CHECK(h->sgram.reserve((size_t)ns * ldys * 8 + 64));
    CRM_CHECK(h->HY.reserve((size_t)ns * mp * 8 + 64));
    CRM_CHECK(h->Zs.reserve((size_t)ns * mp * 8 + 64));
    CRM_CHECK(h->scratch.reserve((size_t)(R + 16) * 8));
    build_ys_kernel<<<blocks_for(h->n * ldys, 256), 256, 0, st>>>(h->Hx.as<double>(), ldH, h->m, c, h->Eext.as<double>(), h->epitch, k0, h->n, h->Ys.as<double>(), ldys);
    CRM_CUDA(cudaGetLastError()); count_launch();
    {
        GemmOperands op{};
        op.A = h->Ys.as<double>(); op.lda = ldys; op.a_cols = ns; op.B = op.A; op.ldb = ldys; op.b_cols = ns; op.B2 = op.A; op.ldb2 = ldys; op.b2_cols = ns;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, ns, 0, ns, h->sgram.as<double>(), ns, 1, st));
    }
    if (mB > 0) {
        GemmOperands op{};
        op.A = h->Hx.as<double>(); op.lda = ldH; op.a_cols = Mx; op.B = h->Ys.as<double>(); op.ldb = ldys; op.b_cols = ns; op.B2 = op.B; op.ldb2 = ldys; op.b2_cols = ns;
        CRM_CHECK(launch_gemm(GEMM_PLAIN, op, (int)h->n, 0, h->m, 0, ns, h->HY.as<double>(), mp, 1, st));
        apply_basis_kernel<<<blocks_for((long long)ns * mp, 128), 128, 0, st>>>(h->Tt.as<double>(), (long long)R * mp, (long long)r0 * mp, h->HY.as<double>(), mp, h->m, mp, ns, h->Zs.as<double>());
        CRM_CUDA(cudaGetLastError()); count_launch();
    }
    double* grid_dev = h->scratch.as<double>();
    CRM_CHECK(upload_small(grid_dev, h->rho.data(), R, st));
    // ---- batches ----
    const double per_snp = 8.0 * ((double)kexp * ldH + 2.0 * h->ld2 + 2.0 * (double)kexp * mp + (double)R * (P + k0 + 8));
    long long B = std::max<long long>(16, std::min<long long>(p, (long long)(4.0e9 / per_snp)));
    B = std::min<long long>(B, 65535LL * GEMM_TILE_N / kexp);
    if (G.on_host) B = host_chunk(h, B, p);
    CRM_CHECK(h->C.reserve((size_t)B * kexp * ldH * 8));
    CRM_CHECK(h->sq.reserve((size_t)B * h->ld2 * 8));
    CRM_CHECK(h->lin.reserve((size_t)B * h->ld2 * 8));
    const size_t pr = (size_t)B * R;
    CRM_CHECK(h->fit_lml.reserve(pr * 8)); CRM_CHECK(h->fit_delta.reserve(pr * 8)); CRM_CHECK(h->fit_scale.reserve(pr * 8));
    CRM_CHECK(h->fit_beta.reserve(pr * P * 8)); CRM_CHECK(h->ucoef.reserve(pr * k0 * 8));
    CRM_CHECK(h->fit_nfev.reserve(pr * 4)); CRM_CHECK(h->fit_flags.reserve(pr * 4));
    CRM_CHECK(h->rho_idx.reserve((size_t)B * 4)); CRM_CHECK(h->best_lml.reserve((size_t)B * 8));
    CRM_CHECK(h->v0.reserve((size_t)B * 8)); CRM_CHECK(h->v1.reserve((size_t)B * 8));
    CRM_CHECK(h->coef.reserve((size_t)B * k0 * 8));
    if (mB > 0) {
        CRM_CHECK(h->Vg.reserve((size_t)mB * round_up(B * kexp, 2) * 8));
        CRM_CHECK(h->Zp.reserve((size_t)B * kexp * mp * 8));
    }
    return for_each_block(h, G, nullptr, p, B, st, [&](const GBlock& kb) -> int {
        GBlock k = kb;
        const long long b = k.b, s0 = k.s0;
        double* C = h->C.as<double>();
        CRM_CHECK(launch_rotation(h, k, C, st));
        CRM_CHECK(block_f64(h, k, st));
        GemmOperands o2{};
        o2.A = h->gs->A2; o2.lda = h->gs->ld2; o2.a_cols = h->M2; o2.B = k.G; o2.ldb = k.ld; o2.b_cols = k.cols; o2.B2 = k.G; o2.ldb2 = k.ld; o2.b2_cols = k.cols;
        CRM_CHECK(launch_gemm(GEMM_PRODUCT, o2, (int)h->gs->K, 0, h->M2, 0, (int)b, h->sq.as<double>(), h->ld2, 1, st));
        CRM_CHECK(launch_gemm(GEMM_PLAIN, o2, (int)h->gs->K, 0, h->M2, 0, (int)b, h->lin.as<double>(), h->ld2, 1, st));
        if (mB > 0) {
            const long long ldv = round_up(b * kexp, 2);
            CRM_CHECK(launch_gather_transpose(C, ldH, nullptr, kexp, 0, kexp, b * kexp, mB, h->Vg.as<double>(), ldv, st));
            GemmOperands o3{};
            o3.A = h->Tt.as<double>(); o3.lda = (long long)R * mp; o3.a_cols = (long long)R * mp;
            o3.B = h->Vg.as<double>(); o3.ldb = ldv; o3.b_cols = b * kexp; o3.B2 = o3.B; o3.ldb2 = ldv; o3.b2_cols = b * kexp;
            CRM_CHECK(launch_gemm(GEMM_PLAIN, o3, mB, r0 * mp, mp, 0, (int)(b * kexp), h->Zp.as<double>(), mp, 1, st));
        }
        BetaArgs ba{};
        ba.S = mB > 0 ? h->S.as<double>() + (long long)r0 * mp : nullptr; ba.Zs = h->Zs.as<double>(); ba.Zp = h->Zp.as<double>();
        ba.S_stride = 0; ba.Zs_stride = 0; ba.Zp_snp_stride = (long long)kexp * mp; ba.Zp_rho_stride = 0;
        ba.has_g = 1; ba.mix_rho = 1; ba.restricted = 1; ba.fixed_x = nullptr; ba.xopt = nullptr;
        ba.shared_gram = h->sgram.as<double>();
        ba.rot = C; ba.rot_ld = ldH; ba.col_y = h->m; ba.col_W = h->m + 1; ba.kexp = kexp;
        ba.lin = h->lin.as<double>(); ba.lin_ld = h->ld2; ba.sq = h->sq.as<double>(); ba.sq_ld = h->ld2;
        ba.rho = grid_dev; ba.m = mB; ba.mp = mp; ba.c = c; ba.k0 = k0; ba.R = R; ba.p = (int)b; ba.n = (double)h->n;
        ba.lml = h->fit_lml.as<double>(); ba.delta = h->fit_delta.as<double>(); ba.scale = h->fit_scale.as<double>();
        ba.beta = h->fit_beta.as<double>(); ba.ucoef = h->ucoef.as<double>(); ba.nfev = h->fit_nfev.as<int>(); ba.flags = h->fit_flags.as<int>();
        CRM_CHECK(launch_beta_fit(ba, st));
        CRM_CHECK(launch_select(ba.lml, ba.delta, ba.scale, (int)b, R, h->rho_idx.as<int>(), h->best_lml.as<double>(), h->v0.as<double>(), h->v1.as<double>(), st));
        finalize_betas_kernel<<<blocks_for(b, 128), 128, 0, st>>>(h->rho_idx.as<int>(), h->v0.as<double>(), grid_dev, ba.beta, ba.ucoef, maf + s0, R, P, c, k0, b,
                                                              out_beta_g + s0, h->coef.as<double>());
        CRM_CUDA(cudaGetLastError()); count_launch();
        if (out_rho1) {
            finalize_interaction_kernel<<<blocks_for(b, 256), 256, 0, st>>>(h->rho_idx.as<int>(), h->v0.as<double>(), h->v1.as<double>(), grid_dev, b, out_rho1 + s0,
                                                                           h->best_lml.as<double>(), h->best_lml.as<double>(), h->best_lml.as<double>());
            CRM_CUDA(cudaGetLastError()); count_launch();
        }
        CRM_CHECK(launch_beta_gxe(h->Eext.as<double>() + 1, h->epitch, h->coef.as<double>(), k0, h->n, b, out_beta_gxe, ldo, s0, st));
        return (int)CRM_OK;
    });
}

}  // namespace crm

// ================================================================================================
// extern "C"
// ================================================================================================
using namespace crm;

struct crm_handle_s { Handle impl; };

extern "C" {

int crm_version(void) { return 100; }
const char* crm_last_error(void) { return g_err; }

int crm_create(crm_handle_t* out, int device) {
    if (!out) { set_error("crm_create: null output"); return CRM_ERR_INVALID; }
    // one model object per gene: nothing here may be slow.  cudaGetDeviceProperties is (milliseconds, with stalls of 0.1-0.3 s every few
    // calls, profiles/r02_step_trace.txt): the two attributes that matter are read once per device
    static int count = -1;
    static int cc_major[64], cc_minor[64];
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (count < 0) {
            int c = 0;
            CRM_CUDA(cudaGetDeviceCount(&c));
            for (int d = 0; d < c && d < 64; d++) {
                CRM_CUDA(cudaDeviceGetAttribute(&cc_major[d], cudaDevAttrComputeCapabilityMajor, d));
                CRM_CUDA(cudaDeviceGetAttribute(&cc_minor[d], cudaDevAttrComputeCapabilityMinor, d));
            }
            count = std::min(c, 64);
        }
    }
    if (device < 0 || device >= count) { set_error("crm_create: device %d not available (%d visible)", device, count); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(device));
    if (cc_major[device] != 10) { set_error("libcrm_b200 is built for sm_100a only; device %d is sm_%d%d", device, cc_major[device], cc_minor[device]); return CRM_ERR_UNSUPPORTED; }
    crm_handle_s* h = new crm_handle_s();
    h->impl.device = device;
    if (device < 32) {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (!g_cache[device].empty()) {
            BufCache c = std::move(g_cache[device].back());
            g_cache[device].pop_back();
            size_t i = 0;
            h->impl.for_each_buf([&](DevBuf& b) { if (i < c.bufs.size()) b = c.bufs[i]; i++; });
            h->impl.adopt_ev = c.ev;
        }
    }
    *out = h;
    return CRM_OK;
}

int crm_destroy(crm_handle_t h) {
    if (!h) return CRM_OK;
    cudaSetDevice(h->impl.device);
    // Buffers go back to the pool in the order of the stream that used them last, after the copy stream has drained into it; no
    // device-wide synchronisation (other models and streams keep running).  A stream that no longer exists falls back to one.
    drop_feed(&h->impl);
    cudaStream_t st = h->impl.last_stream;
    if (st && cudaStreamQuery(st) == cudaErrorInvalidResourceHandle) { cudaGetLastError(); cudaDeviceSynchronize(); st = nullptr; }
    cudaGetLastError();
    if (h->impl.copy_stream) {
        cudaEvent_t drained = h->impl.ev_copy[0];
        if (cudaEventRecord(drained, h->impl.copy_stream) == cudaSuccess) cudaStreamWaitEvent(st, drained, 0);
    }
    if (h->impl.planes_ev_pending) { cudaStreamWaitEvent(st, h->impl.planes_ev, 0); h->impl.planes_ev_pending = false; }
    {
        AllocScope alloc_scope(st);
        if (h->impl.adopt_ev) { cudaStreamWaitEvent(st, h->impl.adopt_ev, 0); cudaEventDestroy(h->impl.adopt_ev); h->impl.adopt_ev = nullptr; }
        static const bool recycle = [] { const char* v = getenv("CRM_NO_RECYCLE"); return !(v && atoi(v) != 0); }();
        bool kept = false;
        if (recycle && h->impl.device < 32) {
            BufCache c;
            if (cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming) == cudaSuccess && cudaEventRecord(c.ev, st) == cudaSuccess) {
                std::lock_guard<std::mutex> lock(g_cache_mu);
                if (g_cache[h->impl.device].size() < BUF_CACHE_DEPTH) {
                    h->impl.for_each_buf([&](DevBuf& b) { c.bufs.push_back(b); b.ptr = nullptr; b.cap = 0; });
                    g_cache[h->impl.device].push_back(std::move(c));
                    kept = true;
                }
            }
            if (!kept && c.ev) cudaEventDestroy(c.ev);
            cudaGetLastError();
        }
        if (!kept) h->impl.free_all();
    }
    if (h->impl.copy_stream) {
        cudaStreamDestroy(h->impl.copy_stream);     // asynchronous: resources are released once the stream has drained
        for (int i = 0; i < 2; i++) { cudaEventDestroy(h->impl.ev_copy[i]); cudaEventDestroy(h->impl.ev_done[i]); }
    }
    if (h->impl.side_stream) { cudaStreamDestroy(h->impl.side_stream); cudaEventDestroy(h->impl.side_ev); cudaEventDestroy(h->impl.planes_ev); }
    for (cudaEvent_t e : h->impl.stage_events) cudaEventDestroy(e);
    for (cudaEvent_t e : h->impl.prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : h->impl.prof_oz_events) cudaEventDestroy(e);
    delete h;
    return CRM_OK;
}

int crm_trim_pool(int device) {
    cudaMemPool_t pool = device_pool(device);
    if (!pool) { set_error("crm_trim_pool: no allocation pool for device %d", device); return CRM_ERR_INVALID; }
    std::vector<BufCache> cached;
    if (device >= 0 && device < 32) { std::lock_guard<std::mutex> lock(g_cache_mu); cached.swap(g_cache[device]); }
    if (!cached.empty()) {
        int prev = 0;
        cudaGetDevice(&prev);
        CRM_CUDA(cudaSetDevice(device));
        CRM_CUDA(cudaDeviceSynchronize());
        for (BufCache& c : cached) { for (DevBuf& b : c.bufs) b.release(); if (c.ev) cudaEventDestroy(c.ev); }
        CRM_CUDA(cudaDeviceSynchronize());
        cudaSetDevice(prev);
    }
    CRM_CUDA(cudaMemPoolTrimTo(pool, 0));
    release_pinned_slots();
    return CRM_OK;
}

int crm_setup(crm_handle_t h, const double* y, const double* W, int64_t ldw, const double* E0, int64_t lde0, const double* E1,
              int64_t lde1, const double* L, int64_t ldl, int64_t n, int c, int k0, int k1, int64_t mL, const double* rho_host, int R,
              void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return do_setup(&h->impl, y, W, ldw, E0, lde0, E1, lde1, L, ldl, n, c, k0, k1, mL, rho_host, R, 0, 1, (cudaStream_t)stream);
}

int crm_setup_partial(crm_handle_t h, const double* y, const double* W, int64_t ldw, const double* E0, int64_t lde0, const double* E1,
                      int64_t lde1, const double* L, int64_t ldl, int64_t n, int c, int k0, int k1, int64_t mL, const double* rho_host, int R,
                      int r_first, int r_step, void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return do_setup(&h->impl, y, W, ldw, E0, lde0, E1, lde1, L, ldl, n, c, k0, k1, mL, rho_host, R, r_first, r_step, (cudaStream_t)stream);
}

int64_t crm_basis_record_size(crm_handle_t h) {
    if (!h || h->impl.m <= 0) return 0;
    return 2 + (int64_t)h->impl.mp + (int64_t)h->impl.m * h->impl.mp;
}

int crm_export_basis(crm_handle_t h, int r, double* out, void* stream) {
    if (!h || !out || r < 0 || r >= h->impl.R) { set_error("crm_export_basis: bad arguments"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    CRM_CUDA(cudaSetDevice(H.device));
    const long long total = 2 + (long long)H.mp + (long long)H.m * H.mp;
    pack_basis_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(H.S.as<double>(), H.Tt.as<double>(), (long long)H.R * H.mp, H.devinfo.as<int>(), H.R, r, H.m, H.mp, out);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int crm_import_basis(crm_handle_t h, int r, const double* in, void* stream) {
    if (!h || !in || r < 0 || r >= h->impl.R) { set_error("crm_import_basis: bad arguments"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    CRM_CUDA(cudaSetDevice(H.device));
    const long long total = 2 + (long long)H.mp + (long long)H.m * H.mp;
    unpack_basis_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, H.R, r, H.m, H.mp, H.S.as<double>(), H.Tt.as<double>(), (long long)H.R * H.mp, H.devinfo.as<int>());
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int crm_setup_finish(crm_handle_t h, void* stream) {
    if (!h || h->impl.m <= 0) { set_error("crm_setup_finish: crm_setup_partial has not run"); return CRM_ERR_STATE; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return finish_setup(&h->impl, (cudaStream_t)stream);
}

int crm_stage_genotypes(crm_handle_t h, const double* G_host, int64_t ldg, int64_t rows, int64_t p, void* stream) {
    return crm_stage_genotypes_typed(h, G_host, CRM_G_F64, ldg, rows, p, 0, stream);
}

int crm_stage_genotypes_typed(crm_handle_t h, const void* G_host, int dtype, int64_t ldg, int64_t rows, int64_t p, int64_t basis_cols_hint, void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return do_stage_genotypes(&h->impl, G_host, dtype, ldg, rows, p, basis_cols_hint, (cudaStream_t)stream);
}

int crm_host_threads(void) { return host_threads(); }

int crm_feeder_blocks(int64_t p, int64_t basis_cols, int64_t* starts, int32_t capacity, int32_t* nblocks) {
    if (p <= 0 || !nblocks) { set_error("crm_feeder_blocks: bad arguments"); return CRM_ERR_INVALID; }
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = -1; }        // no device: the SM count falls back to 148
    const std::vector<long long> st = feeder_block_starts(p, basis_cols, dev);
    *nblocks = (int32_t)st.size() - 1;
    if (starts) for (size_t i = 0; i < st.size() && (int32_t)i < capacity; i++) starts[i] = st[i];
    return CRM_OK;
}

int crm_host_narrow(const void* src_host, int dtype, int64_t ld, int64_t rows, int64_t cols, int8_t* dst_host, int64_t ldd, int32_t* bad, int32_t* gmax) {
    if (!src_host || !dst_host || rows < 0 || cols < 0 || ld < cols || ldd < cols || host_dtype_size(dtype) == 0) { set_error("crm_host_narrow: bad arguments"); return CRM_ERR_INVALID; }
    int b = 0, g = 0;
    host_parallel_narrow(src_host, dtype, ld, rows, 0, cols, dst_host, ldd, &b, &g);
    if (bad) *bad = b;
    if (gmax) *gmax = g;
    return CRM_OK;
}

int crm_fp64_tensor_peak(double* tflops, void* stream) {
    if (!tflops) { set_error("crm_fp64_tensor_peak: null output"); return CRM_ERR_INVALID; }
    return measure_fp64_tensor_peak(tflops, (cudaStream_t)stream);
}

int crm_set_test_contexts(crm_handle_t h, const double* E0, int64_t lde0, void* stream) {
    if (!h || !h->impl.ready || !E0) { set_error("crm_set_test_contexts: handle not set up"); return CRM_ERR_STATE; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return build_test_contexts(&h->impl, E0, lde0, (cudaStream_t)stream);
}

int crm_set_background_factors(crm_handle_t h, const double* hK, int64_t ldhk, int q, const double* M, int r, int* accepted, void* stream) {
    if (accepted) *accepted = 0;
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    if (!hK || !M || q <= 0 || r <= 0 || ldhk < q) { set_error("crm_set_background_factors: bad arguments"); return CRM_ERR_INVALID; }
    if (!H.ready) {
        // before crm_setup: applied by the set-up once its operands exist (hK and M must stay valid until crm_setup returns)
        H.kr_pending = true; H.kr_pending_hK = hK; H.kr_pending_ld = ldhk; H.kr_pending_q = q; H.kr_pending_r = r; H.kr_pending_M = M;
        return CRM_OK;
    }
    H.kr = false;
    CRM_CUDA(cudaSetDevice(H.device));
    cudaStream_t st = (cudaStream_t)stream;
    AllocScope alloc_scope(st); H.last_stream = st;
    CRM_CHECK(adopt_buffers(&H, st));
    bool applicable = false;
    CRM_CHECK(kr_apply(&H, hK, ldhk, q, M, r, st, &applicable));
    if (!applicable) return CRM_OK;
    int flag = 1;
    CRM_CUDA(cudaMemcpyAsync(&flag, H.devinfo.as<int>() + 2 * H.R, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    H.kr_unverified = false;
    if (flag != 0) { H.kr = false; return CRM_OK; }                 // the basis is not the declared product: ignored
    if (accepted) *accepted = H.kr ? 1 : 0;
    return CRM_OK;
}

int crm_hint_integer_genotypes(crm_handle_t h, int likely) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    h->impl.early_planes = likely != 0;
    return CRM_OK;
}

int crm_rotation_rows(crm_handle_t h, int64_t* full_rows, int64_t* used_rows) {
    if (!h || !h->impl.ready) { set_error("crm_rotation_rows: handle not set up"); return CRM_ERR_STATE; }
    if (full_rows) *full_rows = (int64_t)h->impl.kexp * h->impl.ldH;
    if (used_rows) *used_rows = (int64_t)plane_rows(&h->impl);
    return CRM_OK;
}

long long crm_launch_count(void) { return g_launches.load(); }

int crm_profile(crm_handle_t h, int enable, double* rot_ms, double* rot_flops, int64_t* rot_launches) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    double ms = 0.0;
    for (size_t i = 0; i + 1 < H.prof_events.size(); i += 2) {
        float t = 0.f;
        CRM_CUDA(cudaEventSynchronize(H.prof_events[i + 1]));
        CRM_CUDA(cudaEventElapsedTime(&t, H.prof_events[i], H.prof_events[i + 1]));
        ms += t;
    }
    if (rot_ms) *rot_ms = ms;
    if (rot_flops) *rot_flops = H.prof_flops;
    if (rot_launches) *rot_launches = (int64_t)(H.prof_events.size() / 2);
    for (cudaEvent_t e : H.prof_events) cudaEventDestroy(e);
    H.prof_events.clear();
    H.prof_flops = 0.0;
    H.prof_on = enable != 0;
    return CRM_OK;
}

int crm_update_phenotype(crm_handle_t h, const double* y, void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    return do_update_phenotype(&h->impl, y, (cudaStream_t)stream);
}

int crm_set_donors(crm_handle_t h, const int32_t* perm, const int32_t* offsets, int64_t d, void* stream) {
    if (!h || !h->impl.ready) { set_error("crm_set_donors: handle not set up"); return CRM_ERR_STATE; }
    if (!perm || !offsets || d <= 0 || d > 2000000000LL) { set_error("crm_set_donors: bad arguments"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    CRM_CUDA(cudaSetDevice(H.device));
    cudaStream_t st = (cudaStream_t)stream;
    AllocScope alloc_scope(st); H.last_stream = st;
    CRM_CHECK(adopt_buffers(&H, st));
    CRM_CHECK(H.dperm.reserve((size_t)H.n * 4));
    CRM_CHECK(H.doff.reserve((size_t)(d + 1) * 4));
    CRM_CUDA(cudaMemcpyAsync(H.dperm.ptr, perm, (size_t)H.n * 4, cudaMemcpyDeviceToDevice, st));
    CRM_CUDA(cudaMemcpyAsync(H.doff.ptr, offsets, (size_t)(d + 1) * 4, cudaMemcpyDeviceToDevice, st));
    H.donors.K = d;
    H.donors_set = true;
    return aggregate_donors(&H, st);
}

int crm_profile_int8(crm_handle_t h, double* gemm_ms, double* gemm_ops, int64_t* launches) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    Handle& H = h->impl;
    double ms = 0.0;
    for (size_t i = 0; i + 1 < H.prof_oz_events.size(); i += 2) {
        float t = 0.f;
        CRM_CUDA(cudaEventSynchronize(H.prof_oz_events[i + 1]));
        CRM_CUDA(cudaEventElapsedTime(&t, H.prof_oz_events[i], H.prof_oz_events[i + 1]));
        ms += t;
    }
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_ops) *gemm_ops = H.prof_oz_gemm_ops;
    if (launches) *launches = (int64_t)(H.prof_oz_events.size() / 2);
    for (cudaEvent_t e : H.prof_oz_events) cudaEventDestroy(e);
    H.prof_oz_events.clear();
    H.prof_oz_gemm_ops = 0.0;
    return CRM_OK;
}

int crm_get_dims(crm_handle_t h, int64_t* d) {
    if (!h || !h->impl.ready || !d) { set_error("crm_get_dims: handle not set up"); return CRM_ERR_STATE; }
    d[0] = h->impl.n; d[1] = h->impl.c; d[2] = h->impl.k0; d[3] = h->impl.m; d[4] = h->impl.R; d[5] = h->impl.mp; d[6] = h->impl.max_rank; d[7] = h->impl.hxe_decided ? (h->impl.use_hxe ? 1 : 0) : -1;
    return CRM_OK;
}

int crm_get_spectrum(crm_handle_t h, int r, double* out, void* stream) {
    if (!h || !h->impl.ready || !out || r < 0 || r >= h->impl.R) { set_error("crm_get_spectrum: bad arguments"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaMemcpyAsync(out, h->impl.S.as<double>() + (long long)r * h->impl.mp, (size_t)h->impl.mp * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CRM_OK;
}

int crm_scan_interaction(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, const double* Gtest, int64_t ldgt,
                         double* out_pv, double* out_rho1, double* out_e2, double* out_g2, double* out_eps2, const crm_scan_diag_t* diag,
                         void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    const GSource src{G, ldg, (g_on_host >> 4) & 15, g_on_host & 1}, tested{Gtest, ldgt, (g_on_host >> 4) & 15, g_on_host & 1};
    return do_scan_interaction(&h->impl, (g_on_host >> 1) & 1, src, p, Gtest ? &tested : nullptr, out_pv, out_rho1, out_e2, out_g2, out_eps2, diag, (cudaStream_t)stream);
}

int crm_scan_association(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, int fast, double* out_pv,
                         double* out_alt_lml, double* info4, double* out_null_lml, void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    const GSource src{G, ldg, (g_on_host >> 4) & 15, g_on_host & 1};
    return do_scan_association(&h->impl, (g_on_host >> 1) & 1, src, p, fast, out_pv, out_alt_lml, info4, out_null_lml, (cudaStream_t)stream);
}

int crm_predict_interaction(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, const double* maf, int use_background,
                            double* out_beta_g, double* out_beta_gxe, int64_t ldo, double* out_rho1, void* stream) {
    if (!h) { set_error("null handle"); return CRM_ERR_INVALID; }
    CRM_CUDA(cudaSetDevice(h->impl.device));
    AllocScope alloc_scope((cudaStream_t)stream); h->impl.last_stream = (cudaStream_t)stream;
    CRM_CHECK(adopt_buffers(&h->impl, (cudaStream_t)stream));
    const GSource src{G, ldg, (g_on_host >> 4) & 15, g_on_host & 1};
    return do_predict(&h->impl, (g_on_host >> 1) & 1, src, p, maf, use_background, out_beta_g, out_beta_gxe, ldo, out_rho1, (cudaStream_t)stream);
}

int crm_gemm(int mode, const double* A, int64_t lda, int64_t a_cols, const double* B, int64_t ldb, int64_t b_cols, const double* B2,
             int64_t ldb2, int64_t b2_cols, int64_t K, int m_begin, int m_count, int64_t n_begin, int64_t n_count, double* out, int64_t ldc,
             int kexp, void* stream) {
    if (!A || !B || !out || K <= 0 || mode < 0 || mode > 2 || (mode != 0 && !B2)) { set_error("crm_gemm: bad arguments"); return CRM_ERR_INVALID; }
    GemmOperands op{};
    op.A = A; op.lda = lda; op.a_cols = a_cols; op.B = B; op.ldb = ldb; op.b_cols = b_cols;
    op.B2 = B2 ? B2 : B; op.ldb2 = B2 ? ldb2 : ldb; op.b2_cols = B2 ? b2_cols : b_cols;
    return launch_gemm(mode, op, (int)K, m_begin, m_count, (int)n_begin, (int)n_count, out, ldc, kexp, (cudaStream_t)stream);
}

int crm_eigh_batched(const double* A, int n, int batch, double* W, double* V, double* quality_host, float* ms, void* stream) {
    if (!A || !W || !V || n < 2 || batch < 1) { set_error("crm_eigh_batched: bad arguments"); return CRM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    AllocScope alloc_scope(st);
    int dev = 0;
    CRM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { set_error("device index %d outside the supported range", dev); return CRM_ERR_UNSUPPORTED; }
    EigPool& pool = g_eig_pool[dev];
    std::lock_guard<std::mutex> pool_lock(pool.mu);
    if (pool.ctx.empty()) pool.ctx.resize(1);
    EigCtx& e = pool.ctx[0];
    if (!e.solver) CRM_SOLVER(cusolverDnCreate(&e.solver));
    CRM_CHECK(ensure_blas(e));
    size_t ws_bytes = 0;
    CRM_CHECK(eig_workspace_bytes(n, batch, &ws_bytes));
    DevBuf mat, ws, quality, work;
    const size_t nn = (size_t)n * n;
    CRM_CHECK(mat.reserve((size_t)batch * nn * 8)); CRM_CHECK(ws.reserve(ws_bytes)); CRM_CHECK(quality.reserve((size_t)batch * 8 + (size_t)2 * batch * 4));
    int lib_lwork = 0;
    CRM_SOLVER(cusolverDnSetStream(e.solver, st));
    CRM_CHECK(eig_lib_lwork(e.solver, n, &lib_lwork));
    CRM_CHECK(work.reserve((size_t)lib_lwork * 8));
    CRM_CUDA(cudaMemcpyAsync(mat.ptr, A, (size_t)batch * nn * 8, cudaMemcpyDeviceToDevice, st));
    int* lib_info = reinterpret_cast<int*>(quality.as<double>() + batch);
    CRM_CUDA(cudaMemsetAsync(lib_info, 0, (size_t)2 * batch * 4, st));
    cudaEvent_t e0, e1;
    CRM_CUDA(cudaEventCreate(&e0)); CRM_CUDA(cudaEventCreate(&e1));
    CRM_CUDA(cudaEventRecord(e0, st));
    int status = eig_batched(e.solver, e.blas, mat.as<double>(), nullptr, n, batch, W, V, quality.as<double>(), ws.ptr, work.as<double>(), lib_lwork, lib_info, st);
    cudaEventRecord(e1, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    float t = 0.f;
    if (ce == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = t;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (status == CRM_OK && ce != cudaSuccess) { set_error("crm_eigh_batched: %s", cudaGetErrorString(ce)); status = CRM_ERR_CUDA; }
    if (status == CRM_OK && quality_host) {
        std::vector<int> li(2 * batch);
        cudaMemcpy(quality_host, quality.ptr, (size_t)batch * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(li.data(), lib_info, (size_t)2 * batch * 4, cudaMemcpyDeviceToHost);
        for (int b = 0; b < batch; b++) if (li[b] != 0 || li[batch + b] != 0) quality_host[b] = INFINITY;
    }
    mat.release(); ws.release(); quality.release(); work.release();
    return status;
}

int crm_int8_split_gemm(const double* X, int64_t ldx, int64_t cols, const double* G, int64_t ldg, int64_t B, int64_t n, int route, double* C, int64_t ldc,
                        int32_t* flags2, float* contraction_ms, void* stream) {
    if (!X || !G || !C || !flags2 || cols <= 0 || B <= 0 || n <= 0 || ldx < cols || ldg < B || ldc < cols || route < 0 || route > 2) { set_error("crm_int8_split_gemm: bad arguments"); return CRM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    AllocScope alloc_scope(st);
    const long long Mp = round_up(cols, 16), Kp = round_up(n, 16), Bp = round_up(B, 16);
    DevBuf P8, expo, Gt8, flags, D32;
    CRM_CHECK(P8.reserve((size_t)OZAKI_SLICES * Mp * Kp)); CRM_CHECK(expo.reserve((size_t)cols * 4)); CRM_CHECK(Gt8.reserve((size_t)Bp * Kp)); CRM_CHECK(flags.reserve(64));
    CRM_CUDA(cudaMemsetAsync(P8.ptr, 0, (size_t)OZAKI_SLICES * Mp * Kp, st));
    CRM_CHECK(oz_launch_matrix_planes(X, ldx, (int)cols, n, expo.as<int>(), P8.as<int8_t>(), Mp, Kp, st));
    CRM_CHECK(oz_launch_genotypes(G, ldg, n, B, Gt8.as<int8_t>(), nullptr, Bp, Kp, flags.as<int>(), st));
    CRM_CUDA(cudaMemcpyAsync(flags2, flags.ptr, 8, cudaMemcpyDeviceToHost, st));
    CRM_CUDA(cudaStreamSynchronize(st));
    int status = CRM_OK;
    if (flags2[0] == 0 && (double)Kp * 64.0 * (double)std::max(flags2[1], 1) < 2147483648.0) {
        cudaEvent_t e0, e1;
        CRM_CUDA(cudaEventCreate(&e0)); CRM_CUDA(cudaEventCreate(&e1));
        CRM_CUDA(cudaEventRecord(e0, st));
        if (route != 1) status = oz_launch_mma(P8.as<int8_t>(), Mp, cols, expo.as<int>(), Gt8.as<int8_t>(), Bp, B, Kp, C, ldc, st, route == 0 ? 1 : 2);
        else {
            status = D32.reserve((size_t)OZAKI_SLICES * Mp * Bp * sizeof(int));
            if (status == CRM_OK) status = oz_int8_gemm(P8.as<int8_t>(), (long long)OZAKI_SLICES * Mp, Gt8.as<int8_t>(), Bp, Kp, D32.as<int>(), Bp, st);
            if (status == CRM_OK) status = oz_launch_combine(D32.as<int>(), Mp, Bp, expo.as<int>(), cols, B, C, ldc, st);
        }
        cudaEventRecord(e1, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        float ms = 0.f;
        if (ce == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
        if (contraction_ms) *contraction_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (status == CRM_OK && ce != cudaSuccess) { set_error("crm_int8_split_gemm: %s", cudaGetErrorString(ce)); status = CRM_ERR_CUDA; }
    } else {
        set_error("crm_int8_split_gemm: G is not integer-valued in [-127, 127] or the int32 accumulation could overflow");
        status = CRM_ERR_UNSUPPORTED;
    }
    cudaStreamSynchronize(st);
    P8.release(); expo.release(); Gt8.release(); flags.release(); D32.release();
    return status;
}

int crm_lmm_fit_rotated(const double* S, const double* yr, const double* Wr, const double* gr, const double* gy, const double* gW,
                        const double* gg, const double* stats, int m, int mp, int R, int c, int64_t p, double n, int restricted,
                        double* lml, double* delta, double* scale, double* beta, int32_t* nfev, int32_t* flags, void* stream) {
    if (!S || !yr || !Wr || !stats || !lml || !delta || !scale || !beta || !nfev || !flags || p <= 0) { set_error("crm_lmm_fit_rotated: bad arguments"); return CRM_ERR_INVALID; }
    FitArgs fa{};
    fa.S = S; fa.yr = yr; fa.Wr = Wr; fa.gr = gr; fa.gr_ld = (long long)R * mp;
    fa.gy = gy; fa.gy_ld = 1; fa.gW = gW; fa.gW_ld = c; fa.gg = gg; fa.gg_ld = 1; fa.stats = stats;
    fa.m = m; fa.mp = mp; fa.R = R; fa.c = c; fa.p = (int)p; fa.n = n; fa.restricted = restricted; fa.fixed_x = nullptr;
    fa.lml = lml; fa.delta = delta; fa.scale = scale; fa.beta = beta; fa.xopt = nullptr; fa.nfev = nfev; fa.flags = flags;
    return launch_fit(fa, gr != nullptr, (cudaStream_t)stream);
}

int crm_davies_pvalues(const double* Q, const double* lam, const int32_t* nlam, int lam_ld, int64_t count, int lim, double acc, double* pv,
                       double* liu, int32_t* ifault, int32_t* converged, double* trace8, void* stream) {
    if (!Q || !lam || !nlam || !pv || count < 0) { set_error("crm_davies_pvalues: bad arguments"); return CRM_ERR_INVALID; }
    if (count == 0) return CRM_OK;
    PvalArgs pa{};
    pa.Q = Q; pa.lam = lam; pa.nlam = nlam; pa.lam_ld = lam_ld; pa.count = (int)count; pa.lim = lim; pa.acc = acc;
    pa.pv = pv; pa.liu = liu; pa.ifault = ifault; pa.converged = converged; pa.trace = trace8;
    return launch_pvalues(pa, (cudaStream_t)stream);
}

int crm_liu_params(const double* Q, const double* lam, const int32_t* nlam, int lam_ld, int64_t count, double* out4, void* stream) {
    if (!Q || !lam || !nlam || !out4 || count < 0) { set_error("crm_liu_params: bad arguments"); return CRM_ERR_INVALID; }
    return launch_liu_params(Q, lam, nlam, lam_ld, count, out4, (cudaStream_t)stream);
}

int crm_qmin(const double* params4, int nrho, int64_t count, double* out, void* stream) {
    if (!params4 || !out || count < 0 || nrho <= 0) { set_error("crm_qmin: bad arguments"); return CRM_ERR_INVALID; }
    return launch_qmin(params4, nrho, count, out, (cudaStream_t)stream);
}

int crm_lrt_pvalues(const double* alt_lml, double null_lml, int64_t count, double* pv, void* stream) {
    if (!alt_lml || !pv || count < 0) { set_error("crm_lrt_pvalues: bad arguments"); return CRM_ERR_INVALID; }
    if (count == 0) return CRM_OK;
    return launch_lrt(alt_lml, null_lml, count, pv, (cudaStream_t)stream);
}

int crm_lrt_pvalues_dof(const double* alt_lml, double null_lml, int64_t count, double dof, double* pv, void* stream) {
    if (!alt_lml || !pv || count < 0 || !(dof > 0.0)) { set_error("crm_lrt_pvalues_dof: bad arguments"); return CRM_ERR_INVALID; }
    if (count == 0) return CRM_OK;
    if (dof == 1.0) return launch_lrt(alt_lml, null_lml, count, pv, (cudaStream_t)stream);
    return launch_lrt_dof(alt_lml, null_lml, count, dof, pv, (cudaStream_t)stream);
}

}  // extern "C"
