// Translation unit: batched symmetric eigensolver of the set-up (eig.cuh) -- launches and the library steps around them
// (Cholesky-QR re-orthonormalisation with cuBLAS/cuSOLVER, back-transformation with cusolverDnDormtr).
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <vector>

#include "eig.cuh"
#include "launch.cuh"
#include "trace.cuh"

namespace crm {

#define CRM_BLAS(expr)                                                                              \
    do {                                                                                            \
        cublasStatus_t _s = (expr);                                                                 \
        if (_s != CUBLAS_STATUS_SUCCESS) { crm::set_error("%s failed with cuBLAS status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); return crm::CRM_ERR_SOLVER; } \
    } while (0)
#define CRM_SOLVER_(expr)                                                                           \
    do {                                                                                            \
        cusolverStatus_t _s = (expr);                                                               \
        if (_s != CUSOLVER_STATUS_SUCCESS) { crm::set_error("%s failed with cuSOLVER status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); return crm::CRM_ERR_SOLVER; } \
    } while (0)

// residual check: max_t |T z_t - (z_t' T z_t) z_t|_inf / |T| per matrix, and the Rayleigh quotients
__global__ void eig_residual_kernel(const double* d_all, const double* e_all, const double* Z_all, const double* tnorm_all, EigSizes sz, double* rq_all,
                                    unsigned long long* worst_bits) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = sz.n_of[b], nmax = sz.nmax;
    if (t >= n) return;
    const double* d = d_all + (size_t)b * nmax;
    const double* e = e_all + (size_t)b * nmax;
    const double* z = Z_all + (size_t)b * nmax * nmax + (size_t)t * n;
    double rq = 0.0, zz = 0.0;
    for (int i = 0; i < n; i++) {
        const double tz = d[i] * z[i] + (i > 0 ? e[i - 1] * z[i - 1] : 0.0) + (i + 1 < n ? e[i] * z[i + 1] : 0.0);
        rq += z[i] * tz; zz += z[i] * z[i];
    }
    rq /= zz;
    double worst = 0.0;
    for (int i = 0; i < n; i++) {
        const double tz = d[i] * z[i] + (i > 0 ? e[i - 1] * z[i - 1] : 0.0) + (i + 1 < n ? e[i] * z[i + 1] : 0.0);
        worst = fmax(worst, fabs(tz - rq * z[i]));
    }
    rq_all[(size_t)b * nmax + t] = rq;
    worst /= fmax(tnorm_all[b], 1e-300);
    if (!(worst == worst)) worst = INFINITY;
    atomicMax(&worst_bits[b], (unsigned long long)__double_as_longlong(worst));     // non-negative doubles order like their bit patterns
}

// max |Z'Z - I| per matrix from the upper triangle of G = Z'Z (column-major)
__global__ void eig_gram_error_kernel(const double* G_all, EigSizes sz, unsigned long long* emax_bits) {
    const int b = blockIdx.y;
    const int n = sz.n_of[b];
    const double* G = G_all + (size_t)b * sz.nmax * sz.nmax;
    double worst = 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * n; idx += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(idx / n), row = (int)(idx - (long long)col * n);
        if (row > col) continue;
        const double g = fabs(G[idx] - (row == col ? 1.0 : 0.0));
        worst = (g == g) ? fmax(worst, g) : INFINITY;
    }
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&emax_bits[b], (unsigned long long)__double_as_longlong(worst));
}
// Orthonormalisation step to first order: with Z'Z = I + E, |E| << 1, the Cholesky factor is I + striu(E) + diag(E)/2 up to
// O(E^2), so R^-1 = I - striu(E) - diag(E)/2; G (upper triangle, column-major) is turned into that R^-1 in place
__global__ void eig_first_order_rinv_kernel(double* G, int n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * n; idx += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(idx / n), row = (int)(idx - (long long)col * n);
        double v = 0.0;
        if (row < col) v = -G[idx];
        else if (row == col) v = 1.0 - 0.5 * (G[idx] - 1.0);
        G[idx] = v;
    }
}

int eig_workspace_bytes(int n, int batch, size_t* bytes) {
    const size_t nn = (size_t)n * n;
    // d, e, tau, xbuf, pbuf, lam, rq: 7 x [batch][n]; part [batch][32]; tnorm [batch]; worst [batch]; bar [batch]; invit work 5 nn; Z nn; G nn
    *bytes = (size_t)batch * (7 * (size_t)n + 80) * 8 + (size_t)batch * 7 * nn * 8 + 4096;
    return CRM_OK;
}

// Matrix b is n_of[b] x n_of[b] (leading dimension n_of[b]) in slot b of A [batch][nmax][nmax]: symmetric, both triangles
// filled, destroyed.  Eigenvalues (ascending) -> W [batch][nmax]; orthonormal eigenvectors, column-major like cusolverDnDsyevd
// (V[b][t * n_b + i] = component i of vector t) -> slot b of V [batch][nmax][nmax]; quality [batch] (device) =
// largest residual |T z - lambda z|_inf / |T| over the vectors (NaN/inf when the Cholesky-QR step broke down).
// ws: eig_workspace_bytes(n, batch) bytes; lib_work: at least lib_lwork doubles (max of potrf / ormtr needs, queried by the caller).
// side streams of the back-transformation, one set per device (eig_batched runs under the per-device lock of the set-up)
constexpr int EIG_WAYS = 4;
struct EigSide { bool ready = false; cudaStream_t stream[EIG_WAYS - 1]; cusolverDnHandle_t solver[EIG_WAYS - 1]; cudaEvent_t done[EIG_WAYS - 1]; cudaEvent_t fork; };
static EigSide g_eig_side[32];

int eig_batched(void* solver_v, void* blas_v, double* A, const int* n_of, int n, int batch, double* W, double* V, double* quality, void* ws, double* lib_work,
                int lib_lwork, int* info_dev, cudaStream_t st, int group_batch, const int* ids, const std::function<int()>* after_sytrd) {
    cusolverDnHandle_t solver = (cusolverDnHandle_t)solver_v;
    cublasHandle_t blas = (cublasHandle_t)blas_v;
    if (n < 2 || n > SY_MAX_N || batch < 1 || batch > SY_MAX_BATCH) { set_error("eig_batched: n = %d outside [2, %d] or batch = %d outside [1, %d]", n, SY_MAX_N, batch, SY_MAX_BATCH); return CRM_ERR_UNSUPPORTED; }
    EigSizes sz{};
    sz.nmax = n; sz.batch = batch;
    for (int b = 0; b < batch; b++) {
        sz.n_of[b] = n_of ? n_of[b] : n;
        sz.id_of[b] = ids ? ids[b] : b;
        if (sz.n_of[b] < 2 || sz.n_of[b] > n) { set_error("eig_batched: matrix %d has size %d outside [2, %d]", b, sz.n_of[b], n); return CRM_ERR_INVALID; }
    }
    const size_t nn = (size_t)n * n;
    double* p = (double*)ws;
    double* d = p; p += (size_t)batch * n;
    double* e = p; p += (size_t)batch * n;
    double* tau = p; p += (size_t)batch * n;
    double* xbuf = p; p += (size_t)batch * n;
    double* pbuf = p; p += (size_t)batch * n;
    double* rq = p; p += (size_t)batch * n;
    double* part = p; p += (size_t)batch * 32;
    double* tnorm = p; p += batch;
    unsigned int* bar = (unsigned int*)p; p += batch;     // 8 bytes each reserved, 4 used
    double* emax = p; p += batch;
    double* work = p; p += (size_t)batch * 5 * nn;
    double* G = p; p += (size_t)batch * nn;
    CRM_CUDA(cudaMemsetAsync(bar, 0, (size_t)batch * 8, st));
    CRM_CUDA(cudaMemsetAsync(quality, 0, (size_t)batch * 8, st));
    CRM_CUDA(cudaMemsetAsync(emax, 0, (size_t)batch * 8, st));
    PhaseTrace tr(st);
    // 1. tridiagonalisation: every group of SY_GROUP CTAs must be resident -> cooperative launch
    {
        static bool attr = false;
        if (!attr) { CRM_CUDA(cudaFuncSetAttribute(crm_sytrd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_MAX)); attr = true; }
        static int sms = 0;
        if (!sms) { int dev = 0; CRM_CUDA(cudaGetDevice(&dev)); CRM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
        SytrdArgs sa{};
        // CTAs per matrix.  The partial sums of a column step are combined in group order, so a matrix's result depends on the group size:
        // when this batch is a share of a larger one (set-up shared between ranks) the group size of the whole batch is used, and every
        // rank count yields the bits of the single-GPU decomposition.
        sa.group = std::max(1, std::min(SY_MAX_GROUP, sms / std::max(batch, group_batch)));
        for (int b = 0; b < batch; b++) sa.n_of[b] = sz.n_of[b];
        sa.A = A; sa.nmax = n; sa.batch = batch; sa.d = d; sa.e = e; sa.tau = tau; sa.xbuf = xbuf; sa.pbuf = pbuf; sa.part = part; sa.bar = bar;
        // shared memory left after the five work vectors holds the last rows each CTA owns (CRM_SYTRD_RESIDENT=0: none)
        const size_t npad = (size_t)((n + 1) & ~1), base = 5 * npad * 8;
        static const int res_env = [] { const char* v = getenv("CRM_SYTRD_RESIDENT"); return v ? atoi(v) : -1; }();
        const int owned_max = (n + sa.group - 1) / sa.group;
        sa.resident = base < (size_t)SY_SMEM_MAX ? (int)std::min<size_t>((size_t)owned_max, ((size_t)SY_SMEM_MAX - base) / (npad * 8)) : 0;
        if (res_env >= 0) sa.resident = std::min(sa.resident, res_env);
        const size_t smem = std::max(base + (size_t)sa.resident * npad * 8, (size_t)5 * (n + 1) * 8);
        void* params[] = {&sa};
        SlowSection sec("eig: cooperative launch of sytrd");
        CRM_CUDA(cudaLaunchCooperativeKernel((const void*)crm_sytrd_kernel, dim3((unsigned)(batch * sa.group)), dim3(SY_THREADS), params, smem, st));
        count_launch();
    }
    tr.mark("sytrd");
    if (after_sytrd && *after_sytrd) CRM_CHECK((*after_sytrd)());
    // 2. eigenvalues
    crm_tridiag_bisect_kernel<<<dim3((unsigned)(((long long)n * BS_LANES + 255) / 256), (unsigned)batch), 256, (size_t)2 * n * 8, st>>>(d, e, sz, W, tnorm);
    CRM_CUDA(cudaGetLastError()); count_launch();
    tr.mark("bisect");
    // 3. eigenvectors of the tridiagonal matrices
    crm_tridiag_invit_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)batch), 128, 0, st>>>(d, e, W, tnorm, sz, work, V, 3);
    CRM_CUDA(cudaGetLastError()); count_launch();
    tr.mark("invit");
    // 4. Cholesky-QR, twice: Z <- Z R^-1 with R'R = Z'Z (orthonormal bases inside clusters of close eigenvalues)
    CRM_BLAS(cublasSetStream(blas, st));
    CRM_SOLVER_(cusolverDnSetStream(solver, st));
    const double one = 1.0, zero = 0.0;
    // Z'Z = I + E.  Inverse-iteration vectors of well separated eigenvalues are orthogonal to ~eps |T| / gap, so E is tiny unless
    // the spectrum has clusters.  A first-order step Z <- Z (I - striu(E) - diag(E)/2) squares the error; it is applied once when
    // max |E| < 1e-8, twice when < 1e-3, and after a true Cholesky-QR step (potrf + trsm) otherwise.
    std::vector<double> emax_h(batch);
    std::vector<int> todo(batch, 1);
    for (int round = 0; round < 4; round++) {
        bool any = false;
        for (int b = 0; b < batch; b++) any = any || todo[b];
        if (!any) break;
        CRM_CUDA(cudaMemsetAsync(emax, 0, (size_t)batch * 8, st));
        for (int b = 0; b < batch; b++) {
            if (!todo[b]) continue;
            const int nb = sz.n_of[b];
            SlowSection sec1("eig: one cublasDsyrk");
            CRM_BLAS(cublasDsyrk(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, nb, nb, &one, V + (size_t)b * nn, nb, &zero, G + (size_t)b * nn, nb));
        }
        eig_gram_error_kernel<<<dim3(64, (unsigned)batch), 256, 0, st>>>(G, sz, (unsigned long long*)emax);
        CRM_CUDA(cudaGetLastError()); count_launch();
        CRM_CUDA(cudaMemcpyAsync(emax_h.data(), emax, (size_t)batch * 8, cudaMemcpyDeviceToHost, st));
        CRM_CUDA(cudaStreamSynchronize(st));
        for (int b = 0; b < batch; b++) {
            if (!todo[b]) continue;
            const int nb = sz.n_of[b];
            double* Z = V + (size_t)b * nn;
            double* Gb = G + (size_t)b * nn;
            if (!(emax_h[b] < 1e-3)) {                       // clusters (or breakdown: NaN): Cholesky-QR, then look again
                if (round == 3 || !(emax_h[b] == emax_h[b])) { todo[b] = 0; CRM_CUDA(cudaMemsetAsync(quality + b, 0x7f, 8, st)); continue; }   // 0x7f7f...: huge
                CRM_SOLVER_(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_UPPER, nb, Gb, nb, lib_work, lib_lwork, info_dev + b));
                CRM_BLAS(cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, nb, nb, &one, Gb, nb, Z, nb));
                continue;
            }
            if (emax_h[b] < 1e-14) { todo[b] = 0; continue; }   // already orthonormal to rounding
            eig_first_order_rinv_kernel<<<64, 256, 0, st>>>(Gb, nb);
            CRM_CUDA(cudaGetLastError()); count_launch();
            double* tmp = work + (size_t)b * 5 * nn;         // out of place into the free inverse-iteration workspace, then back
            SlowSection sec2("eig: one cublasDtrmm");
            CRM_BLAS(cublasDtrmm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, nb, nb, &one, Gb, nb, Z, nb, tmp, nb));
            CRM_CUDA(cudaMemcpyAsync(Z, tmp, (size_t)nb * nb * 8, cudaMemcpyDeviceToDevice, st));
            if (emax_h[b] < 1e-8) todo[b] = 0;               // error after the step ~ E^2 < 1e-16
        }
    }
    for (int b = 0; b < batch; b++) if (todo[b]) CRM_CUDA(cudaMemsetAsync(quality + b, 0x7f, 8, st));      // did not settle in four rounds
    tr.mark("re-orthonormalisation");
    // residuals of the tridiagonal eigenpairs (before the back-transformation, O(n) per vector)
    eig_residual_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)batch), 128, 0, st>>>(d, e, V, tnorm, sz, rq, (unsigned long long*)quality);
    CRM_CUDA(cudaGetLastError()); count_launch();
    // 5. back-transformation: eigenvectors of A = Q Z, Q from the reflectors left in A.  Dormtr is ~64 small launches per matrix that do
    // not fill the device: the matrices are dealt over EIG_WAYS streams (one cuSOLVER handle each; workspace = the matrix's own, now free,
    // inverse-iteration scratch), forked from and joined to `st` with events.
    CRM_SOLVER_(cusolverDnSetStream(solver, st));
    {
        static const int ways_env = [] { const char* v = getenv("CRM_EIG_WAYS"); return v ? atoi(v) : EIG_WAYS; }();
        const int ways = std::max(1, std::min(std::min(ways_env, EIG_WAYS), batch));
        int dev = 0;
        CRM_CUDA(cudaGetDevice(&dev));
        EigSide& side = g_eig_side[dev & 31];
        const bool own_ws = (size_t)lib_lwork <= 5 * nn;
        if (ways > 1 && own_ws) {
            if (!side.ready) {
                for (int w = 0; w < EIG_WAYS - 1; w++) {
                    CRM_CUDA(cudaStreamCreateWithFlags(&side.stream[w], cudaStreamNonBlocking));
                    CRM_SOLVER_(cusolverDnCreate(&side.solver[w]));
                    CRM_SOLVER_(cusolverDnSetStream(side.solver[w], side.stream[w]));
                    CRM_CUDA(cudaEventCreateWithFlags(&side.done[w], cudaEventDisableTiming));
                }
                CRM_CUDA(cudaEventCreateWithFlags(&side.fork, cudaEventDisableTiming));
                side.ready = true;
            }
            CRM_CUDA(cudaEventRecord(side.fork, st));
            for (int w = 0; w < ways - 1; w++) CRM_CUDA(cudaStreamWaitEvent(side.stream[w], side.fork, 0));
        }
        for (int b = 0; b < batch; b++) {
            const int nb = sz.n_of[b];
            const int w = (ways > 1 && own_ws) ? b % ways : 0;           // way 0 = the caller's stream and handle
            cusolverDnHandle_t hs = w == 0 ? solver : side.solver[w - 1];
            double* ws_b = own_ws ? work + (size_t)b * 5 * nn : lib_work;
            CRM_SOLVER_(cusolverDnDormtr(hs, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, nb, nb, A + (size_t)b * nn, nb, tau + (size_t)b * n, V + (size_t)b * nn, nb,
                                         ws_b, own_ws ? (int)std::min<size_t>(5 * nn, 2000000000u) : lib_lwork, info_dev + batch + b));
        }
        if (ways > 1 && own_ws) {
            for (int w = 0; w < ways - 1; w++) { CRM_CUDA(cudaEventRecord(side.done[w], side.stream[w])); CRM_CUDA(cudaStreamWaitEvent(st, side.done[w], 0)); }
        }
    }
    tr.mark("residuals + back-transformation");
    tr.report("batched eigensolver");
    return CRM_OK;
}

int eig_lib_lwork(void* solver_v, int n, int* lwork) {
    cusolverDnHandle_t solver = (cusolverDnHandle_t)solver_v;
    int l1 = 0, l2 = 0;
    CRM_SOLVER_(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_UPPER, n, nullptr, n, &l1));
    CRM_SOLVER_(cusolverDnDormtr_bufferSize(solver, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, nullptr, n, nullptr, nullptr, n, &l2));
    *lwork = std::max(l1, l2);
    return CRM_OK;
}

}  // namespace crm
