// K5: per-SNP effect-size model of predict_interaction / estimate_betas, one CTA per (SNP, rho1).
//
// Replaces, per SNP and rho1, the reference's economic_qs_linear([sqrt(rho) g.E0 | sqrt(1-rho) L]) +
// LMM(y, [W g E0], QS, restricted=True).fit() and, for the best rho1, the BLUP
// beta_gxe = v0 rho1 E0 (g.E0)' K^-1 (y - M beta) / sqrt(2 maf (1 - maf))
// (cellregmap/_cellregmap.py:137-205; restated in oracle/crm_port.py::predict_interaction).
//
// No per-SNP decomposition: the covariance  s [ (1-d) (rho U U' + (1-rho) B) + d I ],  U = g.E0 (n x k0),
// B = sum_i L_i L_i' = Q_B S_B Q_B' (decomposed once per gene), is a diagonalisable matrix plus a rank-k0 term:
//   A0 = (1-d)(1-rho) B + d I  (diagonal in the Q_B basis),   K/s = A0 + a U U',  a = (1-d) rho,
//   K^-1 = A0^-1 - A0^-1 U (I/a + U'A0^-1 U)^-1 U'A0^-1,   logdet(K/s) = logdet A0 + logdet(I/a + U'A0^-1U) + k0 log a.
// Every likelihood evaluation is one streaming pass forming the A0^-1-Gram of Z = [y | W g E0 | U] from the
// rotated columns, followed by small dense algebra (Cholesky k0 x k0 and P x P) on one warp.
#pragma once
#include "args.cuh"
#include "common.cuh"
#include "fit.cuh"
#include "smallmat.cuh"

namespace crm {

constexpr int BETA_THREADS = 256;
constexpr int BETA_CHUNK = 32;
constexpr int BETA_MAXE = 9;        // Gram entries per thread: NZ (NZ + 1) / 2 <= 9 * 256  ->  NZ <= 67

struct BetaProblem {
    // sizes
    int m, mp, c, k0, hg, P, NZ, lane, tid;
    bool restricted, mix_rho;
    double n, rho;
    // rotated columns
    const double* S; const double* Zs; const double* Zp;
    // shared memory
    double *ZZ, *ZZres, *Gw, *Zt, *wv, *V, *inner, *X, *Gk, *Ar, *br, *tb, *red;
    // per-thread Gram entries
    int ea[BETA_MAXE], eb[BETA_MAXE];
    // design reduction
    unsigned long long mask; int rank; double logdetXX, df;
    int nfev, flags;
    double last_delta, last_scale;

    __device__ __forceinline__ const double* column(int col) const {
        // Z = [y | W (c) | g (hg = 0/1) | E0 (k0) | U (k0)]
        if (col <= c) return Zs + (long long)col * mp;
        if (hg && col == c + 1) return Zp;
        if (col <= c + hg + k0) return Zs + (long long)(col - hg) * mp;
        return Zp + (long long)(col - c - 1 - k0) * mp;      // U_j sits at Zp[hg + j]: col - (1 + P) + hg
    }

    // weighted Gram  sum_i w_i z_a z_b  over the rotated rows into acc[]; mode 0: w = 1, mode 1: w = 1 / d_i (+ sum log d_i)
    __device__ double stream_gram(double (&acc)[BETA_MAXE], int mode, double t, double delta) {
        double ld = 0.0;
#pragma unroll
        for (int u = 0; u < BETA_MAXE; u++) acc[u] = 0.0;
        for (int i0 = 0; i0 < m; i0 += BETA_CHUNK) {
            for (int idx = tid; idx < NZ * BETA_CHUNK; idx += BETA_THREADS) {
                const int col = idx / BETA_CHUNK, ii = idx - col * BETA_CHUNK, i = i0 + ii;
                Zt[col * (BETA_CHUNK + 1) + ii] = (i < m) ? column(col)[i] : 0.0;
            }
            if (tid < BETA_CHUNK) {
                const int i = i0 + tid;
                double w = 0.0;
                if (i < m) { if (mode) { const double d = fma(S[i], t, delta); w = 1.0 / d; ld += log(d); } else w = 1.0; }
                wv[tid] = w;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < BETA_MAXE; u++) {
                if (ea[u] >= 0) {
                    const double* za = Zt + ea[u] * (BETA_CHUNK + 1);
                    const double* zb = Zt + eb[u] * (BETA_CHUNK + 1);
                    double s = acc[u];
#pragma unroll 8
                    for (int ii = 0; ii < BETA_CHUNK; ii++) s = fma(za[ii] * wv[ii], zb[ii], s);
                    acc[u] = s;
                }
            }
            __syncthreads();
        }
        if (mode) {   // only the first warp's first BETA_CHUNK threads hold pieces of ld
            if (tid < 32) { ld = warp_sum(ld); if (tid == 0) red[0] = ld; }
            __syncthreads();
            ld = red[0];
            __syncthreads();
        }
        return ld;
    }

    // -lml at logistic value x; CTA-collective, identical result in all threads
    __device__ __forceinline__ double eval_point(double x, int, int) { return eval(x); }     // (no table of bracket points for wide designs)
    __device__ double eval(double x) {
        nfev++;
        const double delta = logistic_delta(x), omd = 1.0 - delta;
        const double t = mix_rho ? omd * (1.0 - rho) : omd, a = mix_rho ? omd * rho : 0.0;
        double acc[BETA_MAXE];
        double ld = stream_gram(acc, 1, t, delta);
        const double inv_delta = 1.0 / delta;
#pragma unroll
        for (int u = 0; u < BETA_MAXE; u++)
            if (ea[u] >= 0) { const double v = acc[u] + ZZres[ea[u] * NZ + eb[u]] * inv_delta; Gw[ea[u] * NZ + eb[u]] = v; Gw[eb[u] * NZ + ea[u]] = v; }
        __syncthreads();
        const int nzm = 1 + P;            // [y | M] block
        const int u0 = 1 + P;             // first U column
        if (tid < 32) {
            double logdetK = ld + (n - m) * log(delta);
            // Gk = Gw[yM, yM]
            for (int e = lane; e < nzm * nzm; e += 32) { const int i = e / nzm, j = e - i * nzm; Gk[e] = Gw[i * NZ + j]; }
            if (a > 0.0 && k0 > 0) {
                for (int e = lane; e < k0 * k0; e += 32) { const int i = e / k0, j = e - i * k0; inner[e] = Gw[(u0 + i) * NZ + (u0 + j)] + (i == j ? 1.0 / a : 0.0); }
                for (int e = lane; e < k0 * nzm; e += 32) { const int i = e / nzm, j = e - i * nzm; X[e] = Gw[(u0 + i) * NZ + j]; }
                __syncwarp();
                const double piv = warp_cholesky(inner, k0, k0, lane);
                if (!(piv > 0.0)) flags |= 4;
                warp_forward_solve(inner, k0, k0, X, nzm, nzm, lane);
                double li = 0.0;
                for (int j = lane; j < k0; j += 32) li += log(inner[j * k0 + j]);
                logdetK += 2.0 * warp_sum(li) + k0 * log(a);
                for (int e = lane; e < nzm * nzm; e += 32) {
                    const int i = e / nzm, j = e - i * nzm;
                    double s = 0.0;
                    for (int r = 0; r < k0; r++) s += X[r * nzm + i] * X[r * nzm + j];
                    Gk[e] -= s;
                }
            }
            __syncwarp();
            // reduced design: Ar = V' Gk[M,M] V (kept directions), br = V' Gk[M,y]
            // first T = Gk[M,M] V  (stored in Ar scratch rows P..2P-1 is avoided: use X as scratch, k0*nzm >= P*P is not
            // guaranteed, so T lives in `tb` scratch area sized P*P by the host)
            double* T = tb + P;           // tb[0..P) holds the solution, T follows
            for (int e = lane; e < P * P; e += 32) {
                const int i = e / P, j = e - i * P;
                double s = 0.0;
                for (int r = 0; r < P; r++) s += Gk[(1 + i) * nzm + (1 + r)] * V[r * P + j];
                T[e] = s;
            }
            __syncwarp();
            for (int e = lane; e < P * P; e += 32) {
                const int i = e / P, j = e - i * P;
                const bool mi = (mask >> i) & 1ull, mj = (mask >> j) & 1ull;
                double s = 0.0;
                for (int r = 0; r < P; r++) s += V[r * P + i] * T[r * P + j];
                Ar[e] = (mi || mj) ? (i == j ? 1.0 : 0.0) : s;      // dropped directions decoupled with a unit pivot
            }
            for (int i = lane; i < P; i += 32) {
                double s = 0.0;
                for (int r = 0; r < P; r++) s += V[r * P + i] * Gk[(1 + r) * nzm + 0];
                br[i] = ((mask >> i) & 1ull) ? 0.0 : s;
                tb[i] = br[i];
            }
            __syncwarp();
            const double piv = warp_cholesky(Ar, P, P, lane);
            if (!(piv > 0.0)) flags |= 2;
            warp_forward_solve(Ar, P, P, tb, 1, 1, lane);
            double bt = 0.0, la = 0.0;
            for (int i = lane; i < P; i += 32) { bt += tb[i] * tb[i]; la += log(Ar[i * P + i]); }   // b' A^-1 b = |L^-1 b|^2
            bt = warp_sum(bt); la = 2.0 * warp_sum(la);
            warp_backward_solve(Ar, P, P, tb, 1, 1, lane);     // tb = A'^-1 b'
            const double scale = fmax((Gk[0] - bt) / df, CRM_EPS_SMALL);
            double lml = -0.5 * (df * CRM_LOG2PI + df + n * log(scale) + logdetK);
            if (restricted) lml += 0.5 * (logdetXX - (la - rank * log(scale)));
            if (lane == 0) { red[0] = -lml; red[1] = scale; red[2] = delta; }
        }
        __syncthreads();
        const double f = red[0];
        last_scale = red[1]; last_delta = red[2];
        __syncthreads();
        return f;
    }
};

// one CTA per (SNP, rho index)
__global__ void __launch_bounds__(BETA_THREADS) crm_beta_fit_kernel(const BetaArgs a) {
    extern __shared__ __align__(16) double bsm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int s = blockIdx.x, ri = blockIdx.y;
    const int c = a.c, k0 = a.k0, hg = a.has_g ? 1 : 0, P = c + hg + k0, NZ = 1 + P + k0, nzm = 1 + P;
    BetaProblem pr;
    pr.m = a.m; pr.mp = a.mp; pr.c = c; pr.k0 = k0; pr.hg = hg; pr.P = P; pr.NZ = NZ; pr.lane = lane; pr.tid = tid;
    pr.restricted = a.restricted != 0; pr.mix_rho = a.mix_rho != 0;
    pr.n = a.n; pr.rho = a.mix_rho ? a.rho[ri] : 0.0;
    pr.S = a.S + (long long)ri * a.S_stride; pr.Zs = a.Zs + (long long)ri * a.Zs_stride;
    pr.Zp = a.Zp + (long long)s * a.Zp_snp_stride + (long long)ri * a.Zp_rho_stride;
    double* q = bsm;
    pr.ZZ = q; q += NZ * NZ;
    pr.ZZres = q; q += NZ * NZ;
    pr.Gw = q; q += NZ * NZ;
    pr.Zt = q; q += NZ * (BETA_CHUNK + 1);
    pr.wv = q; q += BETA_CHUNK;
    pr.V = q; q += P * P;
    pr.inner = q; q += k0 * k0;
    pr.X = q; q += k0 * nzm;
    pr.Gk = q; q += nzm * nzm;
    pr.Ar = q; q += P * P;
    pr.br = q; q += P;
    pr.tb = q; q += P + P * P;
    pr.red = q; q += 8;
    const int npairs = NZ * (NZ + 1) / 2;
#pragma unroll
    for (int u = 0; u < BETA_MAXE; u++) {
        const int e = tid + u * BETA_THREADS;
        pr.ea[u] = -1; pr.eb[u] = 0;
        if (e < npairs) { int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5); while (r * (r + 1) / 2 > e) r--; while ((r + 1) * (r + 2) / 2 <= e) r++; pr.ea[u] = r; pr.eb[u] = e - r * (r + 1) / 2; }
    }
    // ---- plain Gram Z'Z ----
    const int ns = 1 + c + k0;                       // shared columns [y | W | E0]
    const double* row_g = hg ? a.rot + (long long)s * a.kexp * a.rot_ld : nullptr;
    const double* lin = a.lin ? a.lin + (long long)s * a.lin_ld : nullptr;
    const double* sq = hg ? a.sq + (long long)s * a.sq_ld : nullptr;
#pragma unroll
    for (int u = 0; u < BETA_MAXE; u++) {
        if (pr.ea[u] < 0) continue;
        const int ia = pr.ea[u], ib = pr.eb[u];      // ia >= ib
        // classify columns: kind 0 shared (index into [y|W|E0]), 1 g, 2 U_j
        int ka, ja, kb, jb;
        if (ia <= c) { ka = 0; ja = ia; } else if (hg && ia == c + 1) { ka = 1; ja = 0; } else if (ia <= c + hg + k0) { ka = 0; ja = ia - hg; } else { ka = 2; ja = ia - 1 - P; }
        if (ib <= c) { kb = 0; jb = ib; } else if (hg && ib == c + 1) { kb = 1; jb = 0; } else if (ib <= c + hg + k0) { kb = 0; jb = ib - hg; } else { kb = 2; jb = ib - 1 - P; }
        double zz;
        if (ka == 0 && kb == 0) zz = a.shared_gram[ja * ns + jb];
        else if (ka + kb == 1) {                     // g with a shared column
            const int js = (ka == 0) ? ja : jb;
            zz = (js == 0) ? row_g[a.col_y] : (js <= c ? row_g[a.col_W + js - 1] : lin[1 + (js - 1 - c)]);
        } else if (ka == 1 && kb == 1) zz = sq[0];
        else if (ka == 2 && kb == 2) zz = sq[1 + k0 + pair_index(ja > jb ? ja : jb, ja > jb ? jb : ja)];
        else if (ka + kb == 3) zz = sq[1 + (ka == 2 ? ja : jb)];     // U_j with g
        else {                                       // U_j with a shared column
            const int ju = (ka == 2) ? ja : jb, js = (ka == 2) ? jb : ja;
            const double* rowj = row_g + (long long)(1 + ju) * a.rot_ld;
            if (js == 0) zz = rowj[a.col_y];
            else if (js <= c) zz = rowj[a.col_W + js - 1];
            else { const int je = js - 1 - c; zz = lin[1 + k0 + pair_index(ju > je ? ju : je, ju > je ? je : ju)]; }
        }
        pr.ZZ[ia * NZ + ib] = zz; pr.ZZ[ib * NZ + ia] = zz;
    }
    __syncthreads();
    // ---- complement-space Gram: Z'Z - Zr'Zr ----
    {
        double acc[BETA_MAXE];
        pr.stream_gram(acc, 0, 0.0, 0.0);
#pragma unroll
        for (int u = 0; u < BETA_MAXE; u++)
            if (pr.ea[u] >= 0) { const double v = pr.ZZ[pr.ea[u] * NZ + pr.eb[u]] - acc[u]; pr.ZZres[pr.ea[u] * NZ + pr.eb[u]] = v; pr.ZZres[pr.eb[u] * NZ + pr.ea[u]] = v; }
    }
    __syncthreads();
    // ---- economic SVD of M = [W g E0] through its Gram (warp 0) ----
    if (tid < 32) {
        for (int e = lane; e < P * P; e += 32) { const int i = e / P, j = e - i * P; pr.Ar[e] = pr.ZZ[(1 + i) * NZ + (1 + j)]; }
        __syncwarp();
        warp_jacobi_vec(pr.Ar, pr.V, P, lane);
        if (lane == 0) {
            double lmax = 0.0;
            for (int i = 0; i < P; i++) lmax = fmax(lmax, pr.Ar[i * P + i]);
            unsigned long long mask = 0; int rank = 0; double ld = 0.0;
            for (int i = 0; i < P; i++) { const double l = pr.Ar[i * P + i]; if (l >= CRM_EPS_TINY && l > 1e-13 * lmax) { rank++; ld += log(l); } else mask |= 1ull << i; }
            pr.red[0] = ld; pr.red[1] = (double)rank; pr.red[2] = (double)(mask & 0xffffffffull); pr.red[3] = (double)(mask >> 32);
        }
    }
    __syncthreads();
    pr.logdetXX = pr.red[0]; pr.rank = (int)pr.red[1];
    pr.mask = (unsigned long long)pr.red[2] | ((unsigned long long)pr.red[3] << 32);
    pr.df = pr.restricted ? a.n - pr.rank : a.n;
    pr.nfev = 0; pr.flags = pr.mask ? 1 : 0;
    __syncthreads();
    // ---- fit ----
    double fbest, xbest;
    if (a.fixed_x) xbest = *a.fixed_x; else xbest = brent_minimize(pr, &fbest);
    const double f = pr.eval(xbest);
    pr.nfev--;
    // ---- outputs: lml, delta, scale, beta (design space), ucoef = U' K^-1 (y - M beta) ----
    const long long o = (long long)s * a.R + ri;
    if (tid < 32) {
        // beta = V tb (kept directions)
        for (int i = lane; i < P; i += 32) {
            double bsum = 0.0;
            for (int j = 0; j < P; j++) if (!((pr.mask >> j) & 1ull)) bsum += pr.V[i * P + j] * pr.tb[j];
            pr.br[i] = bsum;
            a.beta[o * P + i] = bsum;
        }
        __syncwarp();
        // u = U'A0^-1 (y - M beta);  U'K^-1 r = (u - Guu inner^-1 u) / s  (a > 0), with inner = Guu + I/a
        const int u0 = 1 + P;
        const double omd = 1.0 - pr.last_delta, av = omd * pr.rho;
        for (int j = lane; j < k0; j += 32) {
            double v = pr.Gw[(u0 + j) * NZ + 0];
            for (int i = 0; i < P; i++) v -= pr.Gw[(u0 + j) * NZ + 1 + i] * pr.br[i];
            pr.X[j] = v; pr.X[k0 + j] = v;
        }
        __syncwarp();
        if (k0 == 0) {
        } else if (av > 0.0) {
            // pr.inner still holds the Cholesky factor of the last evaluation (same x)
            warp_forward_solve(pr.inner, k0, k0, pr.X + k0, 1, 1, lane);
            warp_backward_solve(pr.inner, k0, k0, pr.X + k0, 1, 1, lane);
            for (int j = lane; j < k0; j += 32) {
                double v = pr.X[j];
                for (int r = 0; r < k0; r++) v -= pr.Gw[(u0 + j) * NZ + (u0 + r)] * pr.X[k0 + r];
                a.ucoef[o * k0 + j] = v / pr.last_scale;
            }
        } else {
            for (int j = lane; j < k0; j += 32) a.ucoef[o * k0 + j] = pr.X[j] / pr.last_scale;
        }
        if (lane == 0) {
            a.lml[o] = -f; a.delta[o] = pr.last_delta; a.scale[o] = pr.last_scale; a.nfev[o] = pr.nfev; a.flags[o] = pr.flags;
            if (a.xopt) a.xopt[o] = xbest;
        }
    }
}

// beta_gxe[i][s] = coef[s][:] . E0[i][:]   (n x p output, s contiguous: the reference's (1, n, p) array)
__global__ void crm_beta_gxe_kernel(const double* E0, long long lde0, const double* coef, int k0, long long n, long long p,
                                    double* out, long long ldo, long long s0) {
    extern __shared__ double sc[];          // coef tile: 128 SNPs x k0
    const long long sbase = (long long)blockIdx.x * 128;
    for (int e = threadIdx.x; e < 128 * k0; e += blockDim.x) { const long long s = sbase + e / k0; sc[e] = s < p ? coef[s * k0 + e % k0] : 0.0; }
    __syncthreads();
    const int sl = threadIdx.x & 127;
    const long long s = sbase + sl;
    for (long long i = (long long)blockIdx.y * 2 + (threadIdx.x >> 7); i < n; i += (long long)gridDim.y * 2) {
        double v = 0.0;
        for (int j = 0; j < k0; j++) v += E0[i * lde0 + j] * sc[sl * k0 + j];
        if (s < p) out[i * ldo + s0 + s] = v;
    }
}

}  // namespace crm
