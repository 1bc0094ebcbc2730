// Batched symmetric eigensolver for the set-up (replaces the R sequential cusolverDnDsyevd calls of do_setup; reference:
// numpy_sugar.economic_qs_linear called once per rho in CellRegMap.__init__, cellregmap/_cellregmap.py:108-131).
//
// cuSOLVER's Dsyevd spends 80 % of its 11 ms per 1020 x 1020 problem in a latency-bound tridiagonalisation that fills the GPU
// with one problem at a time.  Here all problems of the rho grid advance together:
//   1. crm_sytrd_kernel       Householder tridiagonalisation, one group of CTAs per matrix (rows dealt cyclically; one group
//                             barrier per column; every CTA keeps the two current Householder vectors in shared memory and
//                             derives the next one redundantly from the pivot row; the rank-2 update of a step is applied
//                             while the next step's A v is accumulated, so the trailing matrix is read and written once per
//                             column).  LAPACK dsytrd('L') storage (reflectors below the sub-diagonal, tau, d, e).
//   2. crm_tridiag_bisect     all eigenvalues by multisection on Sturm counts, BS_LANES lanes per eigenvalue.
//   3. crm_tridiag_invit      eigenvectors by inverse iteration, one thread per eigenvalue (tridiagonal LU with partial
//                             pivoting, random start vectors); orthogonality inside clusters of close eigenvalues is restored
//                             afterwards on all vectors together (kernels_eig.cu: Gram matrix, first-order or Cholesky-QR step).
//   4. back-transformation    cusolverDnDormtr on the reflectors of step 1 (blocked, DMMA GEMMs; 4 ms for 11 x 1020^2 -- two
//                             hand-written one-warp-per-eigenvector kernels were measured at 10-20 ms and dropped).
#pragma once
#include "common.cuh"

namespace crm {

constexpr int SY_MAX_GROUP = 16;      // CTAs per matrix (chosen at launch: SMs / batch, at most this)
constexpr int SY_MAX_BATCH = 64;
constexpr int SY_THREADS = 512;
constexpr int SY_MAX_N = 4096;        // five vectors of n doubles in shared memory
constexpr int SY_SMEM_MAX = 226 * 1024;   // dynamic shared memory of the tridiagonalisation (static: < 1 KB)

struct SytrdArgs {
    double* A;            // [batch][nmax][nmax] slots; matrix b is n_of[b] x n_of[b] (leading dimension n_of[b]) at A + b * nmax * nmax
    int nmax, batch, group;
    int n_of[SY_MAX_BATCH];
    double* d;            // [batch][nmax]
    double* e;            // [batch][nmax]   (n - 1 used)
    double* tau;          // [batch][nmax]   (n - 1 used)
    double* xbuf;         // [batch][nmax]   exchange of A v, odd columns
    double* pbuf;         // [batch][nmax]   exchange of A v, even columns
    double* part;         // [batch][2][SY_MAX_GROUP]  exchange of the partial sums of p'v (even / odd columns)
    unsigned int* bar;    // [batch]      group barrier counters (zeroed before the launch)
    int resident;         // rows per CTA kept in shared memory for the whole factorisation (its last owned rows, the longest-lived ones)
};

// barrier among the `group` CTAs of one matrix (all resident: cooperative launch); counter grows monotonically
__device__ __forceinline__ void group_barrier(unsigned int* counter, unsigned int& epoch, unsigned int group) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += group;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < epoch);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SY_THREADS, 1) crm_sytrd_kernel(SytrdArgs a) {
    extern __shared__ __align__(16) double sy_smem[];
    const int SY_GROUP = a.group;
    const int b = blockIdx.x / SY_GROUP, r = blockIdx.x % SY_GROUP;
    const int n = a.n_of[b], nmax = a.nmax;
    const int npad = (n + 1) & ~1;
    double* v_prev = sy_smem;            // Householder vector of the previous step (zero above its head)
    double* w_prev = sy_smem + npad;
    double* v_cur = sy_smem + 2 * npad;
    double* buf = sy_smem + 3 * npad;    // staging: next column / A v
    double* rowbuf = sy_smem + 4 * npad; // pivot row of the next step, fetched together with A v
    // The last `resident` rows this CTA owns live in shared memory: row i is read and written once per column step j < i, so the rows
    // at the bottom carry most of the traffic (the bottom 28 % of the rows: 41 % of it) -- the kernel is bound by L2 bandwidth.
    double* res = sy_smem + 5 * npad;
    const int owned = (n - r + SY_GROUP - 1) / SY_GROUP;                 // rows r, r + G, ...
    const int l_res0 = max(0, owned - a.resident);
    __shared__ double s_red[SY_THREADS / 32];
    __shared__ double s_scalar[4];       // beta, tau, scale of the current step; p'v
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SY_THREADS / 32;
    double* A = a.A + (size_t)b * nmax * nmax;
    // exchange buffers alternate between columns: a fast CTA may write step j + 1 while a slow one still reads step j
    double* pbuf_even = a.pbuf + (size_t)b * nmax;
    double* pbuf_odd = a.xbuf + (size_t)b * nmax;
    double* part_even = a.part + (size_t)b * 2 * SY_MAX_GROUP;
    double* dout = a.d + (size_t)b * nmax;
    double* eout = a.e + (size_t)b * nmax;
    double* tout = a.tau + (size_t)b * nmax;
    unsigned int* bar = a.bar + b;
    unsigned int epoch = 0;
    for (int i = tid; i < npad; i += SY_THREADS) { v_prev[i] = 0.0; w_prev[i] = 0.0; v_cur[i] = 0.0; }
    for (int l = l_res0 + warp; l < owned; l += nwarps) {
        const double* src = A + (size_t)(r + SY_GROUP * l) * n;
        double* dst = res + (size_t)(l - l_res0) * npad;
        for (int c = lane; c < n; c += 32) dst[c] = __ldcg(&src[c]);
    }
    __syncthreads();

    // Column jn of the current matrix = its row jn (symmetry), which its owner finished updating in the previous pass; with the
    // pending rank-2 update of the previous step applied on the fly.  Every CTA of the group computes the same Householder vector
    // from it (same reduction tree), so no exchange is needed: d[jn], e[jn], tau[jn], v_cur.
    auto next_column = [&](int jn, bool staged) {
        const double vpj = v_prev[jn], wpj = w_prev[jn];
        const double* rowj = A + (size_t)jn * n;
        double nrm = 0.0;
        for (int c = jn + tid; c < n; c += SY_THREADS) {
            const double x = (staged ? rowbuf[c] : __ldcg(&rowj[c])) - (vpj * w_prev[c] + wpj * v_prev[c]);
            if (c == jn) { if (r == 0) dout[jn] = x; }
            else { buf[c] = x; if (c >= jn + 2) nrm += x * x; }
        }
        nrm = warp_sum(nrm);
        if (lane == 0) s_red[warp] = nrm;
        __syncthreads();
        if (warp == 0) {
            double s = lane < nwarps ? s_red[lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) {
                double beta = 0.0, tau = 0.0, scale = 0.0;
                if (jn + 1 < n) {
                    const double alpha = buf[jn + 1];
                    beta = alpha;
                    if (s > 0.0) {                               // dlarfg
                        const double nr = sqrt(alpha * alpha + s);
                        beta = alpha >= 0.0 ? -nr : nr;
                        tau = (beta - alpha) / beta;
                        scale = 1.0 / (alpha - beta);
                    }
                }
                s_scalar[0] = beta; s_scalar[1] = tau; s_scalar[2] = scale;
                if (r == 0) { eout[jn] = beta; tout[jn] = tau; }
            }
        }
        __syncthreads();
        const double tau = s_scalar[1], scale = s_scalar[2];
        for (int c = jn + tid; c < n; c += SY_THREADS) v_cur[c] = (c == jn) ? 0.0 : (c == jn + 1) ? (tau != 0.0 ? 1.0 : 0.0) : buf[c] * scale;
        __syncthreads();
    };

    next_column(0, false);
    for (int j = 0; j + 1 < n; j++) {
        const double tau = s_scalar[1], beta = s_scalar[0];
        double* pbuf = (j & 1) ? pbuf_odd : pbuf_even;
        double* part2 = part_even + (j & 1) * SY_MAX_GROUP;
        // fused pass over the owned rows i >= j + 1: apply the update of step j - 1, accumulate p_i = A[i, j+1:] . v_cur
        double pv = 0.0;
        {
            const int l0 = (j + 1 - r + SY_GROUP - 1) / SY_GROUP;
            for (int l = l0 + warp; ; l += nwarps) {
                const int i = r + SY_GROUP * l;
                if (i >= n) break;
                const bool in_smem = l >= l_res0;
                double* row = in_smem ? res + (size_t)(l - l_res0) * npad : A + (size_t)i * n;
                const double vpi = v_prev[i], wpi = w_prev[i];
                double acc = 0.0;
                if ((n & 1) == 0) {
                    // 16-byte accesses from the even column at or below j + 1 (column j of these rows is dead and v_cur[j] = 0)
                    double2* row2 = reinterpret_cast<double2*>(row);
                    const double2* vp2 = reinterpret_cast<const double2*>(v_prev);
                    const double2* wp2 = reinterpret_cast<const double2*>(w_prev);
                    const double2* vc2 = reinterpret_cast<const double2*>(v_cur);
                    const int h1 = n >> 1;
                    int c = ((j + 1) >> 1) + lane;
                    for (; c + 224 < h1; c += 256) {
                        double2 av[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) av[u] = row2[c + 32 * u];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const double2 w0 = wp2[c + 32 * u], p0 = vp2[c + 32 * u], u0 = vc2[c + 32 * u];
                            av[u].x -= vpi * w0.x + wpi * p0.x; av[u].y -= vpi * w0.y + wpi * p0.y;
                            row2[c + 32 * u] = av[u];
                            acc += av[u].x * u0.x + av[u].y * u0.y;
                        }
                    }
                    for (; c < h1; c += 32) {
                        double2 a0 = row2[c];
                        const double2 w0 = wp2[c], p0 = vp2[c], u0 = vc2[c];
                        a0.x -= vpi * w0.x + wpi * p0.x; a0.y -= vpi * w0.y + wpi * p0.y;
                        row2[c] = a0;
                        acc += a0.x * u0.x + a0.y * u0.y;
                    }
                } else {
                    for (int c = j + 1 + lane; c < n; c += 32) {
                        double a0 = row[c];
                        a0 -= vpi * w_prev[c] + wpi * v_prev[c];
                        row[c] = a0;
                        acc += a0 * v_cur[c];
                    }
                }
                if (in_smem && i == j + 1) {           // pivot row of the next step: every CTA reads it from global memory after the barrier
                    __syncwarp();
                    double* grow = A + (size_t)i * n;
                    for (int c = j + 1 + lane; c < n; c += 32) grow[c] = row[c];
                }
                acc = warp_sum(acc);
                if (lane == 0) { pbuf[i] = acc; pv += acc * v_cur[i]; }
            }
        }
        if (lane == 0) s_red[warp] = pv;
        __syncthreads();
        if (tid == 0) { double s = 0.0; for (int w = 0; w < nwarps; w++) s += s_red[w]; part2[r] = s; }
        group_barrier(bar, epoch, (unsigned)SY_GROUP);          // the one exchange per column: A v and p'v
        if (warp == 0) { double s = lane < SY_GROUP ? __ldcg(&part2[lane]) : 0.0; s = warp_sum(s); if (lane == 0) s_scalar[3] = s; }
        {   // A v and the pivot row of the next step (final since the pass above) in one round trip to L2
            const double* rown = A + (size_t)(j + 1) * n;
            for (int i = j + 1 + tid; i < n; i += SY_THREADS) { const double pi = __ldcg(&pbuf[i]), ri = __ldcg(&rown[i]); buf[i] = pi; rowbuf[i] = ri; }
        }
        __syncthreads();
        // w = tau p - (tau^2 p'v / 2) v, identical in every CTA; roll the vectors
        const double half = 0.5 * tau * tau * s_scalar[3];
        for (int i = tid; i < n; i += SY_THREADS) {
            const double vc = (i > j) ? v_cur[i] : 0.0;
            w_prev[i] = (i > j) ? tau * buf[i] - half * vc : 0.0;
            v_prev[i] = vc;
        }
        __syncthreads();
        // row j is dead now (every CTA has read it before the barrier): its owner stores the reflector there, LAPACK dsytrd('L')
        // storage in the column-major view (v below the sub-diagonal of column j, e[j] on the sub-diagonal)
        if (r == j % SY_GROUP) {
            for (int i = j + 2 + tid; i < n; i += SY_THREADS) A[(size_t)j * n + i] = v_prev[i];
            if (tid == 0) A[(size_t)j * n + j + 1] = beta;
        }
        next_column(j + 1, true);
    }
}

struct EigSizes { int nmax, batch; int n_of[SY_MAX_BATCH]; int id_of[SY_MAX_BATCH]; };   // id_of: identity of a matrix beyond its position in this batch (seeds)

// ---- eigenvalues of symmetric tridiagonal matrices by multisection on Sturm counts ----
constexpr int BS_LANES = 4;
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int n, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
    for (int i = 1; i < n; i++) {
        q = d[i] - x - e2[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
    }
    return cnt;      // number of eigenvalues < x
}

__global__ void __launch_bounds__(256) crm_tridiag_bisect_kernel(const double* d_all, const double* e_all, EigSizes sz, double* lam_all, double* tnorm_all) {
    extern __shared__ double bs_smem[];
    const int n = sz.n_of[blockIdx.y], nmax = sz.nmax;
    double* sd = bs_smem;
    double* se2 = bs_smem + nmax;
    __shared__ double s_bounds[3];
    const int b = blockIdx.y;
    if ((blockIdx.x * blockDim.x) / BS_LANES >= n) return;
    const double* d = d_all + (size_t)b * nmax;
    const double* e = e_all + (size_t)b * nmax;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sd[i] = d[i]; se2[i] = (i + 1 < n) ? e[i] * e[i] : 0.0; }
    __syncthreads();
    if (threadIdx.x < 32) {                                   // Gershgorin interval and pivmin
        double lo = INFINITY, hi = -INFINITY, emax = 0.0;
        for (int i = threadIdx.x; i < n; i += 32) {
            const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < n ? fabs(e[i]) : 0.0;
            lo = fmin(lo, sd[i] - el - er); hi = fmax(hi, sd[i] + el + er); emax = fmax(emax, se2[i]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
        }
        if (threadIdx.x == 0) {
            const double tn = fmax(fabs(lo), fabs(hi));
            s_bounds[0] = lo - 2.0 * tn * CRM_EPS_TINY * n - 2.0e-300; s_bounds[1] = hi + 2.0 * tn * CRM_EPS_TINY * n + 2.0e-300;
            s_bounds[2] = fmax(2.2250738585072014e-308 * fmax(1.0, emax), 2.2250738585072014e-308);
            if (blockIdx.x == 0) tnorm_all[b] = tn;
        }
    }
    __syncthreads();
    // BS_LANES lanes per eigenvalue: (BS_LANES + 1)-section per round -- the kernel is bound by FP64 throughput (one division per
    // Sturm step), so few lanes per eigenvalue beat a whole warp
    const int lane = threadIdx.x & 31, sub = lane % BS_LANES, grp = lane / BS_LANES;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) / BS_LANES;      // eigenvalue index (ascending)
    const bool live = k < n;
    double lo = s_bounds[0], hi = s_bounds[1];
    const double pivmin = s_bounds[2];
    bool done = !live;
    for (int iter = 0; iter < 128; iter++) {
        const double width = hi - lo;
        if (!done && width <= 2.0 * CRM_EPS_TINY * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin) done = true;
        if (__all_sync(0xffffffffu, done)) break;
        const double step = width / (BS_LANES + 1);
        const double x = lo + step * (sub + 1);
        const int cnt = done ? 0 : sturm_count(sd, se2, n, x, pivmin);
        const unsigned ge = (__ballot_sync(0xffffffffu, cnt >= k + 1) >> (grp * BS_LANES)) & ((1u << BS_LANES) - 1u);
        if (!done) {
            const int first = ge ? __ffs(ge) - 1 : BS_LANES;
            const double nlo = first == 0 ? lo : lo + step * first;
            const double nhi = first == BS_LANES ? hi : lo + step * (first + 1);
            if (!(nhi - nlo < width)) done = true;                                // no progress at working precision
            else { lo = nlo; hi = nhi; }
        }
    }
    if (live && sub == 0) lam_all[(size_t)b * nmax + k] = 0.5 * (lo + hi);
}

// ---- eigenvectors by inverse iteration: one thread per (matrix, eigenvalue) ----
// work: 5 arrays [batch][n (row i)][n (thread t)] so that the threads of a warp touch consecutive addresses
__device__ __forceinline__ double invit_rand(unsigned int t, unsigned int i, unsigned int it) {
    unsigned int h = t * 0x9E3779B1u ^ (i + 0x7F4A7C15u) * 0x85EBCA77u ^ (it + 1u) * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return ((double)h + 0.5) * (2.0 / 4294967296.0) - 1.0;
}

__global__ void __launch_bounds__(128) crm_tridiag_invit_kernel(const double* d_all, const double* e_all, const double* lam_all, const double* tnorm_all, EigSizes sz,
                                                                double* work_all, double* Z_all, int iters) {
    const int b = blockIdx.y;
    const int n = sz.n_of[b], nmax = sz.nmax;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double* d = d_all + (size_t)b * nmax;
    const double* e = e_all + (size_t)b * nmax;
    const size_t nn = (size_t)nmax * nmax;
    double* U0 = work_all + (size_t)b * 5 * nn;     // diagonal of U
    double* U1 = U0 + nn;                           // first super-diagonal of U
    double* U2 = U1 + nn;                           // second super-diagonal of U (from row interchanges)
    double* Lm = U2 + nn;                           // multipliers; sign bit trick not used: interchange flags in X below
    double* X = Lm + nn;                            // right-hand side / solution
    double* Z = Z_all + (size_t)b * nn;             // column-major: Z[t * n + i]
    const double tn = fmax(tnorm_all[b], 1e-300);
    const double tiny = CRM_EPS_TINY * tn;
    // shift: eigenvalues closer than a few ulps of the norm to their predecessor are nudged apart so that the LU factors of
    // neighbouring vectors differ (as dstein does inside clusters)
    double lam = lam_all[(size_t)b * nmax + t];
    if (t > 0) {
        int back = 0;
        while (t - back - 1 >= 0 && back < 64 && fabs(lam_all[(size_t)b * nmax + t - back - 1] - lam) <= 10.0 * tiny) back++;
        lam += 10.0 * tiny * back;
    }
#define IDX(i) ((size_t)(i) * n + t)
    // LU of T - lam I with partial pivoting (dgttrf): rows are eliminated top-down; `swapped` is kept in the sign of Lm's
    // companion array by storing the multiplier and a flag in two planes would double the traffic -- instead the flag is
    // encoded by storing the multiplier with an offset of 4 when the rows were interchanged (|multiplier| <= 1 always).
    {
        double du0 = (n > 1) ? e[0] : 0.0;           // super-diagonal entry of the current row
        double dd = d[0] - lam;                      // diagonal entry of the current row
        double du2 = 0.0;
        for (int i = 0; i + 1 < n; i++) {
            const double dl = e[i];                  // sub-diagonal entry of row i + 1
            const double dn = d[i + 1] - lam;        // diagonal of row i + 1
            const double un = (i + 2 < n) ? e[i + 1] : 0.0;   // super-diagonal of row i + 1
            if (fabs(dd) >= fabs(dl)) {              // no interchange
                if (fabs(dd) < tiny) dd = (dd < 0.0) ? -tiny : tiny;
                const double m = dl / dd;
                U0[IDX(i)] = dd; U1[IDX(i)] = du0; U2[IDX(i)] = 0.0; Lm[IDX(i)] = m;
                dd = dn - m * du0; du0 = un; du2 = 0.0;
            } else {                                 // interchange rows i and i + 1
                const double m = dd / dl;
                U0[IDX(i)] = dl; U1[IDX(i)] = dn; U2[IDX(i)] = un; Lm[IDX(i)] = m + 4.0;
                dd = du0 - m * dn; du0 = -m * un; du2 = 0.0;
            }
        }
        if (fabs(dd) < tiny) dd = (dd < 0.0) ? -tiny : tiny;
        U0[IDX(n - 1)] = dd; U1[IDX(n - 1)] = 0.0; U2[IDX(n - 1)] = 0.0;
        (void)du2;
    }
    for (int i = 0; i < n; i++) X[IDX(i)] = invit_rand((unsigned)t, (unsigned)i, (unsigned)sz.id_of[b]);
    constexpr int PF = 8;      // steps whose operands are fetched together: the recurrences are latency chains, the loads are not
    for (int it = 0; it < iters; it++) {
        // forward: apply the row operations of the factorisation to the right-hand side
        double cur = X[IDX(0)];
        for (int i0 = 0; i0 + 1 < n; i0 += PF) {
            double mm[PF], xn[PF];
#pragma unroll
            for (int u = 0; u < PF; u++) { const int i = i0 + u; const bool ok = i + 1 < n; mm[u] = ok ? Lm[IDX(i)] : 0.0; xn[u] = ok ? X[IDX(i + 1)] : 0.0; }
#pragma unroll
            for (int u = 0; u < PF; u++) {
                const int i = i0 + u;
                if (i + 1 < n) {
                    double m = mm[u];
                    if (m > 2.0) { m -= 4.0; X[IDX(i)] = xn[u]; cur = cur - m * xn[u]; }       // interchanged
                    else { X[IDX(i)] = cur; cur = xn[u] - m * cur; }
                }
            }
        }
        X[IDX(n - 1)] = cur;
        // backward: U x = rhs, U upper triangular with two super-diagonals
        double x1 = 0.0, x2 = 0.0, amax = 0.0;
        for (int i0 = n - 1; i0 >= 0; i0 -= PF) {
            double u0[PF], u1[PF], u2[PF], xb[PF];
#pragma unroll
            for (int u = 0; u < PF; u++) { const int i = i0 - u; const bool ok = i >= 0; u0[u] = ok ? U0[IDX(i)] : 1.0; u1[u] = ok ? U1[IDX(i)] : 0.0; u2[u] = ok ? U2[IDX(i)] : 0.0; xb[u] = ok ? X[IDX(i)] : 0.0; }
#pragma unroll
            for (int u = 0; u < PF; u++) {
                const int i = i0 - u;
                if (i >= 0) {
                    const double x0 = (xb[u] - u1[u] * x1 - u2[u] * x2) / u0[u];
                    X[IDX(i)] = x0; amax = fmax(amax, fabs(x0));
                    x2 = x1; x1 = x0;
                }
            }
        }
        // normalise (inf-norm between iterations keeps everything in range)
        const double sc = amax > 0.0 ? 1.0 / amax : 1.0;
        for (int i = 0; i < n; i++) X[IDX(i)] *= sc;
    }
    double ss = 0.0;
    for (int i = 0; i < n; i++) { const double x = X[IDX(i)]; ss += x * x; }
    const double sc = ss > 0.0 ? rsqrt(ss) : 0.0;
    for (int i = 0; i < n; i++) Z[(size_t)t * n + i] = X[IDX(i)] * sc;
#undef IDX
}

}  // namespace crm
