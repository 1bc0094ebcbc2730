// K4: p-values of Q ~ sum_j lambda_j chi2_1 -- Davies' AS 155 inversion with the modified-Liu fall-backs,
// one warp per SNP.
//
// Replaces chiscore.davies_pvalue(Q, M, True) (cellregmap/_cellregmap.py:333,435), chiscore.liu_sf
// (cellregmap/_math.py:169,179) and scipy chi2 tails (cellregmap/_cellregmap.py:465-468); restated in
// oracle/chiscore_port.py and oracle/qfc_oracle.c.  The control flow of the published routine
// (findu / ctff / cfe / truncation, the auxiliary-integration loop, the count>lim abort, its constants)
// is kept step by step; only the inversion sum `integrate` is spread over the 32 lanes.
// Specialised to what the path uses: all degrees of freedom 1, all non-centralities 0, sigma = 0.
#pragma once
#include "common.cuh"
#include "args.cuh"

namespace crm {

// ---------- regularised incomplete gamma, chi-square tails ----------
__device__ inline double igam_series(double a, double x) {   // P(a, x), x small relative to a
    double ax = a * log(x) - x - lgamma(a);
    if (ax < -709.0) return 0.0;
    ax = exp(ax);
    double r = a, c = 1.0, ans = 1.0;
    do { r += 1.0; c *= x / r; ans += c; } while (c / ans > 1.1102230246251565e-16);
    return ans * ax / a;
}
__device__ inline double igamc(double a, double x) {         // Q(a, x) = 1 - P(a, x)
    if (x <= 0.0 || a <= 0.0) return 1.0;
    if (x < 1.0 || x < a) return 1.0 - igam_series(a, x);
    double ax = a * log(x) - x - lgamma(a);
    if (ax < -745.0) return 0.0;
    ax = exp(ax);
    const double big = 4.503599627370496e15, biginv = 2.22044604925031308085e-16;
    double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
    double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x, ans = pkm1 / qkm1, t;
    int it = 0;
    do {
        c += 1.0; y += 1.0; z += 2.0;
        const double yc = y * c, pk = pkm1 * z - pkm2 * yc, qk = qkm1 * z - qkm2 * yc;
        if (qk != 0.0) { const double r = pk / qk; t = fabs((ans - r) / r); ans = r; } else t = 1.0;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
        if (fabs(pk) > big) { pkm2 *= biginv; pkm1 *= biginv; qkm2 *= biginv; qkm1 *= biginv; }
    } while (t > 1.1102230246251565e-16 && ++it < 5000);
    return ans * ax;
}
__device__ inline double chi2_sf(double x, double df) { return igamc(0.5 * df, 0.5 * x); }

// non-central chi-square survival function as a Poisson mixture of central ones (nc is tiny on this path)
__device__ inline double ncx2_sf(double x, double df, double nc) {
    if (x <= 0.0) return 1.0;
    const double h = 0.5 * nc;
    double w = exp(-h), sum = 0.0, wsum = 0.0;
    for (int j = 0; j < 5000; j++) {
        sum += w * igamc(0.5 * df + j, 0.5 * x);
        wsum += w;
        w *= h / (j + 1.0);
        if (j + 1 > h && w < 1e-17 * wsum) break;
    }
    return sum;
}

// chi-square quantile: x with P(df/2, x/2) = p  (scipy.stats.chi2.ppf as used by qmin, cellregmap/_math.py:195);
// safeguarded Newton on the regularised lower incomplete gamma function
__device__ inline double chi2_ppf(double p, double df) {
    if (!(p > 0.0)) return 0.0;
    if (!(p < 1.0)) return INFINITY;
    const double a = 0.5 * df, lga = lgamma(a);
    double lo = 0.0, hi = fmax(df, 1.0);
    for (int it = 0; it < 200 && 1.0 - igamc(a, 0.5 * hi) < p; it++) { lo = hi; hi *= 2.0; }
    double x;
    {   // Wilson-Hilferty start, clipped into the bracket
        const double z = normcdfinv(p), c = 2.0 / (9.0 * df), w = 1.0 - c + z * sqrt(c);
        x = df * w * w * w;
        if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
    }
    for (int it = 0; it < 100; it++) {
        const double f = (1.0 - igamc(a, 0.5 * x)) - p;
        if (f > 0.0) hi = x; else lo = x;
        const double pdf = 0.5 * exp((a - 1.0) * log(0.5 * x) - 0.5 * x - lga);
        double xn = x - f / pdf;
        if (!(xn > lo && xn < hi) || !(pdf > 0.0)) xn = 0.5 * (lo + hi);
        if (fabs(xn - x) <= 1e-15 * fabs(x)) { x = xn; break; }
        x = xn;
    }
    return x;
}

struct LiuParams { double pv, dof_x, delta_x, mu_q, sigma_q; };

// chiscore.liu_sf(q, lambda, dofs=1, deltas=0, kurtosis=True)
__device__ inline LiuParams liu_mod(double q, const double* lam, int r) {
    double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    for (int j = 0; j < r; j++) { const double l = lam[j], l2 = l * l; c1 += l; c2 += l2; c3 += l2 * l; c4 += l2 * l2; }
    const double s1 = c3 / (sqrt(c2) * sqrt(c2) * sqrt(c2)), s2 = c4 / (c2 * c2), s12 = s1 * s1;
    double delta_x, dof_x;
    if (s12 > s2) {
        const double a = 1.0 / (s1 - sqrt(s12 - s2));
        delta_x = s1 * a * a * a - a * a;
        dof_x = a * a - 2.0 * delta_x;
    } else { delta_x = 0.0; dof_x = 1.0 / s2; }
    LiuParams o;
    o.mu_q = c1; o.sigma_q = sqrt(2.0 * c2);
    const double mu_x = dof_x + delta_x, sigma_x = sqrt(2.0 * (dof_x + 2.0 * delta_x));
    const double tfinal = (q - o.mu_q) / o.sigma_q * sigma_x + mu_x;
    o.pv = ncx2_sf(tfinal, dof_x, fmax(delta_x, 1e-9));
    o.dof_x = dof_x; o.delta_x = delta_x;
    return o;
}

// ---------- Davies AS 155 ----------
#define QF_PI 3.14159265358979
#define QF_LOG28 0.0866

struct Davies {
    const double* lb;   // eigenvalues (shared or global), r of them
    const int* th;      // indices ordered by |lb| descending
    int r, lim, count, lane;
    bool fail, aborted;
    double c, sigsq, lmax, lmin, mean, intl, ersm;

    __device__ static double exp1(double x) { return x < -50.0 ? 0.0 : exp(x); }
    __device__ static double log1(double x, bool first) {
        if (fabs(x) > 0.1) return first ? log(1.0 + x) : (log(1.0 + x) - x);
        double y = x / (2.0 + x), term = 2.0 * y * y * y, k = 3.0;
        double acc = (first ? 2.0 : -x) * y;
        y = y * y;
        for (double nxt = acc + term / k; nxt != acc; nxt = acc + term / k) { k += 2.0; term *= y; acc = nxt; }
        return acc;
    }
    __device__ void tick() { count++; if (count > lim) aborted = true; }

    __device__ double errbd(double u, double* cx) {
        tick();
        double xconst = u * sigsq, sum1 = u * xconst;
        u = 2.0 * u;
        for (int j = r - 1; j >= 0; j--) {
            const double lj = lb[j], x = u * lj, y = 1.0 - x;
            xconst += lj / y;
            sum1 += x * x / y + log1(-x, false);
        }
        *cx = xconst;
        return exp1(-0.5 * sum1);
    }
    __device__ double ctff(double accx, double* upn) {
        double u2 = *upn, u1 = 0.0, c1 = mean, c2 = 0.0, xc;
        const double rb = 2.0 * ((u2 > 0.0) ? lmax : lmin);
        double u = u2 / (1.0 + u2 * rb);
        while (errbd(u, &c2) > accx) {
            if (aborted) return c2;
            u1 = u2; c1 = c2; u2 = 2.0 * u2;
            u = u2 / (1.0 + u2 * rb);
        }
        u = (c1 - mean) / (c2 - mean);
        while (u < 0.9) {
            if (aborted) return c2;
            u = (u1 + u2) / 2.0;
            if (errbd(u / (1.0 + u * rb), &xc) > accx) { u1 = u; c1 = xc; } else { u2 = u; c2 = xc; }
            u = (c1 - mean) / (c2 - mean);
        }
        *upn = u2;
        return c2;
    }
    __device__ double truncation(double u, double tausq) {
        tick();
        double prod2 = 0.0, prod3 = 0.0;
        int ns = 0;
        const double sum2 = (sigsq + tausq) * u * u;
        double prod1 = 2.0 * sum2;
        u = 2.0 * u;
        for (int j = 0; j < r; j++) {
            const double ul = u * lb[j], x = ul * ul;
            if (x > 1.0) { prod2 += log(x); prod3 += log1(x, true); ns += 1; }
            else prod1 += log1(x, true);
        }
        prod2 += prod1; prod3 += prod1;
        double x = exp1(-0.25 * prod2) / QF_PI;
        const double y = exp1(-0.25 * prod3) / QF_PI;
        double err1 = (ns == 0) ? 1.0 : x * 2.0 / ns;
        double err2 = (prod3 > 1.0) ? 2.5 * y : 1.0;
        if (err2 < err1) err1 = err2;
        x = 0.5 * sum2;
        err2 = (x <= y) ? 1.0 : y / x;
        return (err1 < err2) ? err1 : err2;
    }
    __device__ void findu(double* utx, double accx) {
        const double divis[4] = {2.0, 1.4, 1.2, 1.1};
        double ut = *utx, u = ut / 4.0;
        if (truncation(u, 0.0) > accx) {
            for (u = ut; truncation(u, 0.0) > accx; u = ut) { if (aborted) return; ut *= 4.0; }
        } else {
            ut = u;
            for (u = u / 4.0; truncation(u, 0.0) <= accx; u = u / 4.0) { if (aborted) return; ut = u; }
        }
        for (int i = 0; i < 4; i++) { u = ut / divis[i]; if (truncation(u, 0.0) <= accx) ut = u; }
        *utx = ut;
    }
    // inversion sum: terms k = 0..nterm spread over the lanes, partial sums combined by a butterfly
    __device__ void integrate(int nterm, double interv, double tausq, bool mainx) {
        const double inpi = interv / QF_PI;
        double s_int = 0.0, s_err = 0.0;
        for (int k = nterm - lane; k >= 0; k -= 32) {
            const double u = (k + 0.5) * interv;
            double sum1 = -2.0 * u * c, sum2 = fabs(sum1), sum3 = -0.5 * sigsq * u * u;
            for (int j = r - 1; j >= 0; j--) {
                const double x = 2.0 * lb[j] * u;
                sum3 -= 0.25 * log1(x * x, true);
                const double z = atan(x);
                sum1 += z; sum2 += fabs(z);
            }
            double x = inpi * exp1(sum3) / u;
            if (!mainx) x *= (1.0 - exp1(-0.5 * tausq * u * u));
            s_int += sin(0.5 * sum1) * x;
            s_err += 0.5 * sum2 * x;
        }
        intl += warp_sum(s_int);
        ersm += warp_sum(s_err);
    }
    __device__ double cfe(double x) {
        tick();
        double axl = fabs(x), sum1 = 0.0;
        const double sxl = (x > 0.0) ? 1.0 : -1.0;
        for (int j = r - 1; j >= 0; j--) {
            const int t = th[j];
            if (lb[t] * sxl > 0.0) {
                const double lj = fabs(lb[t]);
                const double axl1 = axl - lj, axl2 = lj / QF_LOG28;
                if (axl1 > axl2) axl = axl1;
                else {
                    if (axl > axl2) axl = axl2;
                    sum1 = (axl - axl1) / lj;
                    for (int k = j - 1; k >= 0; k--) sum1 += 1.0;
                    break;
                }
            }
        }
        if (sum1 > 100.0) { fail = true; return 1.0; }
        return pow(2.0, sum1 / 4.0) / (QF_PI * axl * axl);
    }

    // returns qfval = P(Q < c); *ifault as in the published routine; trace[0..6] optional
    __device__ double run(double cq, int lim_, double acc, int* ifault, double* trace) {
        double qfval = -1.0, acc1 = acc, xlim = (double)lim_;
        double utx, tausq, sd, intv = 0.0, intv1, x, up, un, d1, d2, almx, xnt = 0.0, xntm;
        double tr[7] = {0, 0, 0, 0, 0, 0, 0};
        *ifault = 0;
        lim = lim_; c = cq; count = 0; intl = 0.0; ersm = 0.0; fail = false; aborted = false;
        sigsq = 0.0; sd = 0.0; lmax = 0.0; lmin = 0.0; mean = 0.0;
        for (int j = 0; j < r; j++) {
            const double lj = lb[j];
            sd += lj * lj * 2.0; mean += lj;
            if (lmax < lj) lmax = lj; else if (lmin > lj) lmin = lj;
        }
        if (sd == 0.0) { qfval = (c > 0.0) ? 1.0 : 0.0; goto done; }
        if (lmin == 0.0 && lmax == 0.0) { *ifault = 3; goto done; }
        sd = sqrt(sd);
        almx = (lmax < -lmin) ? -lmin : lmax;
        utx = 16.0 / sd; up = 4.5 / sd; un = -up;
        findu(&utx, 0.5 * acc1);
        if (aborted) goto abort;
        if (c != 0.0 && almx > 0.07 * sd) {
            tausq = 0.25 * acc1 / cfe(c);
            if (aborted) goto abort;
            if (fail) fail = false;
            else {
                const double trn = truncation(utx, tausq);
                if (aborted) goto abort;
                if (trn < 0.2 * acc1) {
                    sigsq += tausq;
                    findu(&utx, 0.25 * acc1);
                    if (aborted) goto abort;
                    tr[5] = sqrt(tausq);
                }
            }
        }
        tr[4] = utx; acc1 = 0.5 * acc1;
        for (;;) {
            d1 = ctff(acc1, &up) - c;
            if (aborted) goto abort;
            if (d1 < 0.0) { qfval = 1.0; goto done; }
            d2 = c - ctff(acc1, &un);
            if (aborted) goto abort;
            if (d2 < 0.0) { qfval = 0.0; goto done; }
            intv = 2.0 * QF_PI / ((d1 > d2) ? d1 : d2);
            xnt = utx / intv; xntm = 3.0 / sqrt(acc1);
            if (xnt > xntm * 1.5) {
                if (xntm > xlim) { *ifault = 1; goto done; }
                const int ntm = (int)floor(xntm + 0.5);
                intv1 = utx / ntm; x = 2.0 * QF_PI / intv1;
                if (x <= fabs(c)) break;
                const double e1 = cfe(c - x); if (aborted) goto abort;
                const double e2 = cfe(c + x); if (aborted) goto abort;
                tausq = 0.33 * acc1 / (1.1 * (e1 + e2));
                if (fail) break;
                acc1 = 0.67 * acc1;
                integrate(ntm, intv1, tausq, false);
                xlim -= xntm; sigsq += tausq;
                tr[2] += 1; tr[1] += ntm + 1;
                findu(&utx, 0.25 * acc1);
                if (aborted) goto abort;
                acc1 = 0.75 * acc1;
                continue;
            }
            break;
        }
        tr[3] = intv;
        if (xnt > xlim) { *ifault = 1; goto done; }
        {
            const int nt = (int)floor(xnt + 0.5);
            integrate(nt, intv, 0.0, true);
            tr[2] += 1; tr[1] += nt + 1;
            qfval = 0.5 - intl;
            tr[0] = ersm;
            up = ersm; x = up + acc / 10.0;
            const double rats[4] = {1.0, 2.0, 4.0, 8.0};
            for (int j = 0; j < 4; j++) if (rats[j] * x == rats[j] * up) *ifault = 2;
        }
        goto done;
    abort:
        *ifault = 4;
    done:
        tr[6] = (double)count;
        if (trace) for (int j = 0; j < 7; j++) trace[j] = tr[j];
        return qfval;
    }
};

constexpr int PV_WARPS = 4;


// chiscore._pvalue_lambda: Davies p-value with the Liu fall-backs (one eigenvalue; p > 1 or p <= 0)
__global__ void __launch_bounds__(PV_WARPS * 32) crm_pvalue_kernel(const PvalArgs a) {
    __shared__ double s_lam[PV_WARPS][PV_MAXLAM];
    __shared__ int s_th[PV_WARPS][PV_MAXLAM];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * PV_WARPS + warp;
    if (i >= a.count) return;
    const int r = a.nlam[i];
    double* lam = s_lam[warp];
    int* th = s_th[warp];
    if (r <= 0 || r > PV_MAXLAM) {   // "No eigenvalue is bigger than 0!!" / unsupported
        if (lane == 0) {
            a.pv[i] = nan("");
            if (a.liu) a.liu[i] = nan("");
            if (a.ifault) a.ifault[i] = -1;
            if (a.converged) a.converged[i] = 0;
        }
        return;
    }
    for (int j = lane; j < r; j += 32) lam[j] = a.lam[(long long)i * a.lam_ld + j];
    __syncwarp();
    if (lane == 0) {   // stable insertion ordering by |lambda| descending (order() of the published routine)
        for (int j = 0; j < r; j++) {
            const double lj = fabs(lam[j]);
            int k = j - 1;
            while (k >= 0 && lj > fabs(lam[th[k]])) { th[k + 1] = th[k]; k--; }
            th[k + 1] = j;
        }
    }
    __syncwarp();
    const double q = a.Q[i];
    const LiuParams lp = liu_mod(q, lam, r);
    Davies dv;
    dv.lb = lam; dv.th = th; dv.r = r; dv.lane = lane;
    int ifault; double tr[7];
    const double qfval = dv.run(q, a.lim, a.acc, &ifault, tr);
    double p = 1.0 - qfval;
    int conv = 1;
    if (r == 1) p = lp.pv;
    else if (ifault != 0) conv = 0;
    if (p > 1.0 || p <= 0.0) { conv = 0; p = lp.pv; }
    if (lane == 0) {
        a.pv[i] = p;
        if (a.liu) a.liu[i] = lp.pv;
        if (a.ifault) a.ifault[i] = ifault;
        if (a.converged) a.converged[i] = conv;
        if (a.trace) { double* t = a.trace + (long long)i * 8; t[0] = qfval; for (int j = 0; j < 7; j++) t[1 + j] = tr[j]; }
    }
}

// score_statistic_liu_params (cellregmap/_math.py:163-180), batched: out[i] = {pv, mu_q, sigma_q, dof_x}
__global__ void crm_liu_params_kernel(const double* Q, const double* lam, const int* nlam, int lam_ld, int count, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const LiuParams lp = liu_mod(Q[i], lam + (long long)i * lam_ld, nlam[i]);
    double* o = out + (long long)i * 4;
    o[0] = lp.pv; o[1] = lp.mu_q; o[2] = lp.sigma_q; o[3] = lp.dof_x;
}

// qmin (cellregmap/_math.py:183-201), batched over `count` tests with `nrho` parameter sets {pv, mu_q, sigma_q, dof_x} each
__global__ void crm_qmin_kernel(const double* params, int nrho, int count, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double* pr = params + (long long)i * nrho * 4;
    double T = INFINITY;
    for (int r = 0; r < nrho; r++) T = fmin(T, pr[r * 4]);
    for (int r = 0; r < nrho; r++) {
        const double mu_q = pr[r * 4 + 1], sigma_q = pr[r * 4 + 2], dof = pr[r * 4 + 3];
        const double q = chi2_ppf(1.0 - T, dof);
        out[(long long)i * nrho + r] = (q - dof) / sqrt(2.0 * dof) * sigma_q + mu_q;
    }
}

// lrt_pvalues (cellregmap/_cellregmap.py:443-469), dof = 1:  clip(chi2_1.sf(clip(2(l1 - l0), tiny_min, inf)))
__global__ void crm_lrt_kernel(const double* alt_lml, double null_lml, int count, double* pv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double lr = -2.0 * null_lml + 2.0 * alt_lml[i];
    lr = fmax(lr, 2.2250738585072014e-308);
    double p = erfc(sqrt(0.5 * lr));
    pv[i] = fmin(fmax(p, 2.2250738585072014e-308), 1.0 - CRM_EPS_TINY);
}

// the same for any number of degrees of freedom: chi2(dof).sf through the regularised upper incomplete gamma function
__global__ void crm_lrt_dof_kernel(const double* alt_lml, double null_lml, int count, double dof, double* pv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double lr = -2.0 * null_lml + 2.0 * alt_lml[i];
    lr = fmax(lr, 2.2250738585072014e-308);
    const double p = chi2_sf(lr, dof);
    pv[i] = fmin(fmax(p, 2.2250738585072014e-308), 1.0 - CRM_EPS_TINY);
}

}  // namespace crm
