// K2: batched two-component LMM fits, one warp per (SNP, rho1) pair.
//
// Replaces glimix_core.lmm.LMM(y, [W g], QS, restricted).fit() + .lml()/.v0/.v1/.beta inside the rho1 loop of the
// reference (cellregmap/_cellregmap.py:345-357; also :254-255,274-276 with restricted=False), restated in
// oracle/lmm_port.py + oracle/brent_port.py.  The warp works on the rotated sufficient statistics
// (S0, Q0'y, Q0'W shared per rho1; Q0'g per SNP) and runs the same bracket + Brent search on logit(delta)
// (rtol = atol = 1e-6), with beta and scale at their conditional optima at every evaluation.
#pragma once
#include "common.cuh"
#include "smallmat.cuh"
#include "args.cuh"

namespace crm {


__device__ __forceinline__ double logistic_delta(double x) {
    double v;
    if (x > 0.0) v = 1.0 / (1.0 + exp(-x));
    else { v = exp(x); v = v / (v + 1.0); }
    return fmin(fmax(v, CRM_EPS_TINY), 1.0 - CRM_EPS_TINY);
}

// ---- table of the bracket points ----
// The bracket search starts at logit(delta) = 0 with a step of 2e-6 and doubles its stride until the objective rises: whatever the SNP,
// it walks one of two fixed sequences of points (towards +inf or -inf), ~20 of the ~27 evaluations of a fit.  At those points everything
// that does not involve the genotype -- the weights w_i = 1 / ((1 - delta) S_i + delta), sum log D_i, the weighted Grams of [y | W] -- is
// the same for every SNP of a rho: crm_fit_table_kernel computes it once per rho and point with the arithmetic of FitProblem::eval, and
// eval_tab() adds the three genotype terms (one multiply and P fused multiply-adds per element instead of a division, a quarter of a log and
// the full Gram).  Same operations in the same order on the genotype terms, identical values elsewhere: the fits are bit-identical.
constexpr int FIT_TAB_K = 30;          // points per direction (2e-6 * 2^30 is far beyond where any search stops)
constexpr int FIT_TAB_HDR = 64;        // doubles of header per record: [x, delta, ld, syy, sXy (C), sXX (C x C)], then w[mp]
__host__ __device__ inline long long fit_tab_record(int mp) { return (long long)FIT_TAB_HDR + mp; }
// point k of the sequence in direction dir (0: towards +inf, 1: towards -inf; k = 0, 1 are the two starting points, shared)
__device__ __forceinline__ double fit_bracket_point(int dir, int k) {
    const double a0 = -CRM_LOGMAX, b0 = CRM_LOGMAX, rtol = 1e-6, atol = 1e-6, gfactor = 2.0;
    double x0 = fmin(fmax(0.0, a0), b0);
    const double step0 = gfactor * (rtol * fabs(x0) + atol);
    double x1 = (x0 - a0 > b0 - x0) ? fmax(x0 - step0, a0) : fmin(x0 + step0, b0);
    if (k == 0) return x0;
    if (k == 1) return x1;
    if (dir) { const double t = x0; x0 = x1; x1 = t; }
    for (int t = 2; t <= k; t++) {
        double x2 = x1 + (x1 - x0) * gfactor;
        x2 = fmin(fmax(x2, a0), b0);
        x0 = x1; x1 = x2;
    }
    return x1;
}

template <int P, bool HAS_G>
struct FitProblem {
    const double* tab = nullptr;   // records of this rho: [2][tab_k][FIT_TAB_HDR + mp]
    int tab_k = 0;
    const double *S, *yr, *Wr, *gr;
    int m, mp, lane;
    double n, df;
    bool restricted;
    double yy_res, Xy_res[P], XX_res[P][P];
    double Vx[P][P];      // eigenvectors of X'X (design reparametrisation tX = X Vx)
    unsigned mask;        // dropped design directions
    int rank;
    double logdetXX;
    int nfev, flags;
    // results of the last evaluation
    double last_delta, last_scale, last_tbeta[P];

    __device__ __forceinline__ void load_x(int i, double (&xv)[P]) const {
        constexpr int C = HAS_G ? P - 1 : P;
#pragma unroll
        for (int a = 0; a < C; a++) xv[a] = Wr[(long long)a * mp + i];
        if (HAS_G) xv[P - 1] = gr[i];
    }

    // plain Grams XX (P x P), Xy (P), yy -> residual (complement-space) terms and the design reduction
    __device__ void init(const double (&XX)[P][P], const double (&Xy)[P], double yy) {
        double syy = 0.0, sXy[P], sXX[P][P];
#pragma unroll
        for (int a = 0; a < P; a++) { sXy[a] = 0.0;
#pragma unroll
            for (int b = 0; b < P; b++) sXX[a][b] = 0.0; }
        for (int i = lane; i < m; i += 32) {
            double xv[P]; load_x(i, xv);
            const double yv = yr[i];
            syy += yv * yv;
#pragma unroll
            for (int a = 0; a < P; a++) { sXy[a] += xv[a] * yv;
#pragma unroll
                for (int b = 0; b <= a; b++) sXX[a][b] += xv[a] * xv[b]; }
        }
        yy_res = yy - warp_sum(syy);
#pragma unroll
        for (int a = 0; a < P; a++) { Xy_res[a] = Xy[a] - warp_sum(sXy[a]);
#pragma unroll
            for (int b = 0; b <= a; b++) { XX_res[a][b] = XX[a][b] - warp_sum(sXX[a][b]); XX_res[b][a] = XX_res[a][b]; } }
        // economic SVD of X through its Gram: keep sigma >= sqrt(eps)
        double A[P][P];
#pragma unroll
        for (int a = 0; a < P; a++)
#pragma unroll
            for (int b = 0; b < P; b++) A[a][b] = XX[a][b];
        jacobi_eig<P>(A, Vx);
        double lmax = 0.0;
#pragma unroll
        for (int a = 0; a < P; a++) lmax = fmax(lmax, A[a][a]);
        mask = 0; rank = 0; logdetXX = 0.0;
#pragma unroll
        for (int a = 0; a < P; a++) {
            const double l = A[a][a];
            if (l >= CRM_EPS_TINY && l > 1e-13 * lmax) { rank++; logdetXX += log(l); }
            else mask |= 1u << a;
        }
        df = restricted ? n - rank : n;
        nfev = 0; flags = 0;
        if (mask) flags |= 1;   // rank-deficient design
    }

    // -lml at logistic value x (beta, scale optimal); warp-collective, result identical in all lanes
    __device__ double eval(double x) {
        nfev++;
        const double delta = logistic_delta(x), omd = 1.0 - delta;
        double syy = 0.0, ld = 0.0, sXy[P], sXX[P][P];
#pragma unroll
        for (int a = 0; a < P; a++) { sXy[a] = 0.0;
#pragma unroll
            for (int b = 0; b < P; b++) sXX[a][b] = 0.0; }
        // sum of log D_i as the log of products of four (D_i in [eps, max S]: no over/underflow), one log per four elements
        double prod = 1.0;
        int cnt = 0;
        for (int i = lane; i < m; i += 32) {
            const double D = fma(S[i], omd, delta);
            const double w = 1.0 / D;
            prod *= D;
            if (++cnt == 4) { ld += log(prod); prod = 1.0; cnt = 0; }
            double xv[P]; load_x(i, xv);
            const double yv = yr[i], wy = w * yv;
            syy += wy * yv;
#pragma unroll
            for (int a = 0; a < P; a++) { const double wx = w * xv[a]; sXy[a] += wx * yv;
#pragma unroll
                for (int b = 0; b <= a; b++) sXX[a][b] += wx * xv[b]; }
        }
        ld += log(prod);
        syy = warp_sum(syy);
        ld = warp_sum(ld) + (n - m) * log(delta);
#pragma unroll
        for (int a = 0; a < P; a++) { sXy[a] = warp_sum(sXy[a]);
#pragma unroll
            for (int c2 = 0; c2 <= a; c2++) sXX[a][c2] = warp_sum(sXX[a][c2]); }
        return finish(delta, syy, ld, sXy, sXX);
    }

    // the same evaluation at bracket point (dir, k) of the table: only the genotype terms are summed here
    __device__ double eval_tab(int dir, int k) {
        constexpr int C = HAS_G ? P - 1 : P;
        nfev++;
        const double* rec = tab + ((long long)dir * tab_k + k) * fit_tab_record(mp);
        const double* w = rec + FIT_TAB_HDR;
        double sXy[P], sXX[P][P];
#pragma unroll
        for (int a = 0; a < P; a++) { sXy[a] = 0.0;
#pragma unroll
            for (int b = 0; b < P; b++) sXX[a][b] = 0.0; }
        if (HAS_G) {
            for (int i = lane; i < m; i += 32) {
                const double wx = w[i] * gr[i];
                sXy[P - 1] += wx * yr[i];
#pragma unroll
                for (int b = 0; b < C; b++) sXX[P - 1][b] += wx * Wr[(long long)b * mp + i];
                sXX[P - 1][P - 1] += wx * gr[i];
            }
            sXy[P - 1] = warp_sum(sXy[P - 1]);
#pragma unroll
            for (int b = 0; b < P; b++) sXX[P - 1][b] = warp_sum(sXX[P - 1][b]);
        }
#pragma unroll
        for (int a = 0; a < C; a++) { sXy[a] = rec[4 + a];
#pragma unroll
            for (int b = 0; b <= a; b++) sXX[a][b] = rec[4 + C + a * C + b]; }
        return finish(rec[1], rec[3], rec[2], sXy, sXX);
    }
    // bracket point (dir, k) at x: from the table when it holds exactly this point
    __device__ __forceinline__ double eval_point(double x, int dir, int k) {
        if (tab && k < tab_k && tab[((long long)dir * tab_k + k) * fit_tab_record(mp)] == x) return eval_tab(dir, k);
        return eval(x);
    }

    // objective from the reduced sums (lower triangle of sXX): reduced design, conditional optimum of beta and scale
    __device__ double finish(double delta, double syy_sum, double ld, const double (&sXy)[P], const double (&sXX)[P][P]) {
        const double inv_delta = 1.0 / delta;
        const double yKy = syy_sum + yy_res * inv_delta;
        double A[P][P], b[P];
#pragma unroll
        for (int a = 0; a < P; a++) { b[a] = sXy[a] + Xy_res[a] * inv_delta;
#pragma unroll
            for (int c2 = 0; c2 <= a; c2++) { A[a][c2] = sXX[a][c2] + XX_res[a][c2] * inv_delta; A[c2][a] = A[a][c2]; } }
        // reparametrise: A' = Vx' A Vx, b' = Vx' b ; dropped directions zeroed
        double T[P][P], Ar[P][P], br[P];
#pragma unroll
        for (int a = 0; a < P; a++)
#pragma unroll
            for (int j = 0; j < P; j++) { double s = 0.0;
#pragma unroll
                for (int r = 0; r < P; r++) s += A[a][r] * Vx[r][j]; T[a][j] = s; }
#pragma unroll
        for (int i = 0; i < P; i++) { double sb = 0.0;
#pragma unroll
            for (int r = 0; r < P; r++) sb += Vx[r][i] * b[r];
            br[i] = (mask >> i & 1u) ? 0.0 : sb;
#pragma unroll
            for (int j = 0; j < P; j++) { double s = 0.0;
#pragma unroll
                for (int r = 0; r < P; r++) s += Vx[r][i] * T[r][j];
                Ar[i][j] = ((mask >> i & 1u) || (mask >> j & 1u)) ? 0.0 : s; } }
#pragma unroll
        for (int i = 0; i < P; i++)
#pragma unroll
            for (int j = i + 1; j < P; j++) { const double s = 0.5 * (Ar[i][j] + Ar[j][i]); Ar[i][j] = s; Ar[j][i] = s; }
        double tb[P], logdetA; bool pd;
        sym_pinv_solve<P>(Ar, br, mask, CRM_EPS_SMALL, tb, &logdetA, &pd);
        double bt = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++) bt += br[i] * tb[i];
        const double scale = fmax((yKy - bt) / df, CRM_EPS_SMALL);
        double lml = -0.5 * (df * CRM_LOG2PI + df + n * log(scale) + ld);
        if (restricted) {
            if (!pd) flags |= 2;   // det(H) not positive (the reference raises ValueError)
            lml += 0.5 * (logdetXX - (logdetA - rank * log(scale)));
        }
        last_delta = delta; last_scale = scale;
#pragma unroll
        for (int i = 0; i < P; i++) last_tbeta[i] = tb[i];
        return -lml;
    }
};

// bracket + Brent (oracle/brent_port.py), f = -lml, on [-LOGMAX, LOGMAX], rtol = atol = 1e-6
template <class Prob>
__device__ double brent_minimize(Prob& pr, double* fbest) {
    const double a0 = -CRM_LOGMAX, b0 = CRM_LOGMAX, rtol = 1e-6, atol = 1e-6, gfactor = 2.0, GOLD = 0.381966011250105097;
    const int maxiter = 500;
    // ---- bracket ----
    double x0 = fmin(fmax(0.0, a0), b0);
    const double step0 = gfactor * (rtol * fabs(x0) + atol);
    double x1 = (x0 - a0 > b0 - x0) ? fmax(x0 - step0, a0) : fmin(x0 + step0, b0);
    double f0 = pr.eval_point(x0, 0, 0), f1 = pr.eval_point(x1, 0, 1);
    int dir = 0;                         // which of the two fixed point sequences the search walks (fit_bracket_point)
    if (f0 < f1) { double tx = x0; x0 = x1; x1 = tx; double tf = f0; f0 = f1; f1 = tf; dir = 1; }
    double x2 = x1, f2 = f1;
    for (int it = 0; it < maxiter; it++) {
        x2 = x1 + (x1 - x0) * gfactor;
        x2 = fmin(fmax(x2, a0), b0);
        if (x2 == x1) { f2 = f1; break; }
        f2 = pr.eval_point(x2, dir, it + 2);
        if (f2 > f1) break;
        x0 = x1; f0 = f1; x1 = x2; f1 = f2;
    }
    if (x0 > x2) { double tx = x0; x0 = x2; x2 = tx; double tf = f0; f0 = f2; f2 = tf; }
    // ---- Brent localmin on [x0, x2] from (x1, f1) ----
    double a = x0, b = x2;
    double xb = x1, fb = f1;            // best point
    double xv1 = xb, fv1 = fb, xv2 = xb, fv2 = fb;
    double d = 0.0, e = 0.0;
    for (int it = 0; it < maxiter; it++) {
        const double mid = 0.5 * (a + b);
        const double tol = rtol * fabs(xb) + atol, tol2 = 2.0 * tol;
        if (fabs(xb - mid) <= tol2 - 0.5 * (b - a)) break;
        double p = 0.0, q = 0.0, r = 0.0;
        if (tol < fabs(e)) {
            r = (xb - xv1) * (fb - fv2);
            q = (xb - xv2) * (fb - fv1);
            p = (xb - xv2) * q - (xb - xv1) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = d;
        }
        double u;
        if (fabs(p) < fabs(0.5 * q * r) && q * (a - xb) < p && p < q * (b - xb)) {
            d = p / q;
            u = xb + d;
            if ((u - a) < tol2 || (b - u) < tol2) d = (xb < mid) ? tol : -tol;
        } else {
            e = ((xb < mid) ? b : a) - xb;
            d = GOLD * e;
        }
        if (fabs(d) >= tol) u = xb + d;
        else if (d > 0.0) u = xb + tol;
        else u = xb - tol;
        const double fu = pr.eval(u);
        if (fu <= fb) {
            if (u < xb) b = xb; else a = xb;
            xv2 = xv1; fv2 = fv1; xv1 = xb; fv1 = fb; xb = u; fb = fu;
        } else {
            if (u < xb) a = u; else b = u;
            if (fu <= fv1 || xv1 == xb) { xv2 = xv1; fv2 = fv1; xv1 = u; fv1 = fu; }
            else if (fu <= fv2 || xv2 == xb || xv2 == xv1) { xv2 = u; fv2 = fu; }
        }
    }
    *fbest = fb;
    return xb;
}

// one warp per (rho, direction, point): the genotype-free part of FitProblem::eval at that point, same arithmetic (see FIT_TAB_* above)
template <int C>
__global__ void __launch_bounds__(32) crm_fit_table_kernel(const FitArgs args, double* tab_all) {
    const int k = blockIdx.x, dir = blockIdx.y, rho = blockIdx.z, lane = threadIdx.x;
    if (dir == 1 && k < 2) return;                      // the two starting points are shared (stored under direction 0)
    const int mp = args.mp, m = args.m;
    const double* S = args.S + (long long)rho * mp;
    const double* yr = args.yr + (long long)rho * mp;
    const double* Wr = args.Wr + (long long)rho * C * mp;
    double* rec = tab_all + (((long long)rho * 2 + dir) * args.tab_k + k) * fit_tab_record(mp);
    const double x = fit_bracket_point(dir, k);
    const double delta = logistic_delta(x), omd = 1.0 - delta;
    double syy = 0.0, ld = 0.0, sXy[C], sXX[C][C];
#pragma unroll
    for (int a = 0; a < C; a++) { sXy[a] = 0.0;
#pragma unroll
        for (int b = 0; b < C; b++) sXX[a][b] = 0.0; }
    double prod = 1.0;
    int cnt = 0;
    for (int i = lane; i < m; i += 32) {
        const double D = fma(S[i], omd, delta);
        const double w = 1.0 / D;
        prod *= D;
        if (++cnt == 4) { ld += log(prod); prod = 1.0; cnt = 0; }
        rec[FIT_TAB_HDR + i] = w;
        double xv[C];
#pragma unroll
        for (int a = 0; a < C; a++) xv[a] = Wr[(long long)a * mp + i];
        const double yv = yr[i], wy = w * yv;
        syy += wy * yv;
#pragma unroll
        for (int a = 0; a < C; a++) { const double wx = w * xv[a]; sXy[a] += wx * yv;
#pragma unroll
            for (int b = 0; b <= a; b++) sXX[a][b] += wx * xv[b]; }
    }
    ld += log(prod);
    syy = warp_sum(syy);
    ld = warp_sum(ld) + (args.n - m) * log(delta);
#pragma unroll
    for (int a = 0; a < C; a++) { sXy[a] = warp_sum(sXy[a]);
#pragma unroll
        for (int b = 0; b <= a; b++) sXX[a][b] = warp_sum(sXX[a][b]); }
    if (lane == 0) {
        rec[0] = x; rec[1] = delta; rec[2] = ld; rec[3] = syy;
#pragma unroll
        for (int a = 0; a < C; a++) { rec[4 + a] = sXy[a];
#pragma unroll
            for (int b = 0; b <= a; b++) rec[4 + C + a * C + b] = sXX[a][b]; }
    }
}

constexpr int FIT_WARPS = 8;

template <int P, bool HAS_G>
__global__ void __launch_bounds__(FIT_WARPS * 32, (P <= 3 ? 2 : 1)) crm_fit_kernel(const FitArgs args, const int use_smem) {
    extern __shared__ __align__(16) double fsm[];
    constexpr int C = HAS_G ? P - 1 : P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rho = blockIdx.y;
    const int s = blockIdx.x * FIT_WARPS + warp;
    const int mp = args.mp, m = args.m;
    const double* S = args.S + (long long)rho * mp;
    const double* yr = args.yr + (long long)rho * mp;
    const double* Wr = args.Wr + (long long)rho * C * mp;
    const double* gr = HAS_G ? args.gr + (long long)(s < args.p ? s : 0) * args.gr_ld + (long long)rho * mp : nullptr;
    if (use_smem) {   // stage the shared per-rho vectors and each warp's rotated genotype
        for (int i = threadIdx.x; i < mp; i += blockDim.x) {
            fsm[i] = S[i]; fsm[mp + i] = yr[i];
            for (int a = 0; a < C; a++) fsm[(2 + a) * mp + i] = Wr[(long long)a * mp + i];
        }
        if (HAS_G && s < args.p) {
            double* mine = fsm + (long long)(2 + C + warp) * mp;
            for (int i = lane; i < mp; i += 32) mine[i] = gr[i];
            gr = mine;
        }
        __syncthreads();
        S = fsm; yr = fsm + mp; Wr = fsm + 2 * mp;
    }
    if (s >= args.p) return;

    FitProblem<P, HAS_G> pr;
    pr.S = S; pr.yr = yr; pr.Wr = Wr; pr.gr = gr; pr.m = m; pr.mp = mp; pr.lane = lane;
    pr.n = args.n; pr.restricted = args.restricted != 0;
    if (HAS_G && args.tab) { pr.tab = args.tab + (long long)rho * 2 * args.tab_k * fit_tab_record(mp); pr.tab_k = args.tab_k; }
    double XX[P][P], Xy[P];
#pragma unroll
    for (int a = 0; a < C; a++) { Xy[a] = args.stats[1 + a];
#pragma unroll
        for (int b = 0; b < C; b++) XX[a][b] = args.stats[1 + C + a * C + b]; }
    if (HAS_G) {
        Xy[P - 1] = args.gy[(long long)s * args.gy_ld];
#pragma unroll
        for (int a = 0; a < C; a++) { XX[P - 1][a] = args.gW[(long long)s * args.gW_ld + a]; XX[a][P - 1] = XX[P - 1][a]; }
        XX[P - 1][P - 1] = args.gg[(long long)s * args.gg_ld];
    }
    pr.init(XX, Xy, args.stats[0]);
    double fbest, xbest;
    if (args.fixed_x) xbest = *args.fixed_x;
    else xbest = brent_minimize(pr, &fbest);
    const double f = pr.eval(xbest);   // refresh beta / scale at the optimum (LMM.fit epilogue)
    pr.nfev--;
    if (lane == 0) {
        const long long o = (long long)s * args.R + rho;
        args.lml[o] = -f;
        args.delta[o] = pr.last_delta;
        args.scale[o] = pr.last_scale;
        if (args.xopt) args.xopt[o] = xbest;
        args.nfev[o] = pr.nfev;
        args.flags[o] = pr.flags;
#pragma unroll
        for (int a = 0; a < P; a++) {
            double bsum = 0.0;
#pragma unroll
            for (int j = 0; j < P; j++) if (!(pr.mask >> j & 1u)) bsum += pr.Vx[a][j] * pr.last_tbeta[j];
            args.beta[o * P + a] = bsum;
        }
    }
}

}  // namespace crm
