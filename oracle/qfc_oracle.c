/*
 * ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of Davies' algorithm AS 155 ("The distribution of a linear combination of
 * chi-squared random variables", Appl. Statist. 29 (1980) 323-333) as it is reached by the
 * reference through   cellregmap/_cellregmap.py:333,435  ->  chiscore.davies_pvalue  ->
 * chi2comb (C library, a repackaging of the published `qfc` routine; chiscore>=0.2.3 is pinned
 * in setup.cfg:27-34, chi2comb is its dependency; neither is vendored under /root/reference nor
 * installable here).  PARITY UNPINNED against the installed dependency; checked against an
 * independent Imhof quadrature and closed-form chi-square tails in tests/test_oracle_qfc.py.
 *
 * State lives in a struct (re-entrant); the published routine's longjmp on `count > lim`
 * is expressed with an abort flag that every loop tests.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/libqfc_oracle.so oracle/qfc_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>

#define QF_PI 3.14159265358979
#define QF_LOG28 0.0866

typedef struct {
    int r, lim, count, sorted, fail, aborted;
    const double *lb, *nc;
    const int *n;
    int *th;
    double c, sigsq, lmax, lmin, mean, intl, ersm;
} qf_t;

static double sq(double x) { return x * x; }
static double exp1(double x) { return x < -50.0 ? 0.0 : exp(x); }

static void tick(qf_t *s) { s->count += 1; if (s->count > s->lim) s->aborted = 1; }

/* first ? log(1+x) : log(1+x)-x, with the series of the published routine for |x| <= 0.1 */
static double log1(double x, int first)
{
    if (fabs(x) > 0.1) return first ? log(1.0 + x) : (log(1.0 + x) - x);
    double y = x / (2.0 + x), term = 2.0 * y * y * y, k = 3.0;
    double acc = (first ? 2.0 : -x) * y;
    y = y * y;
    for (double nxt = acc + term / k; nxt != acc; nxt = acc + term / k) { k += 2.0; term *= y; acc = nxt; }
    return acc;
}

/* stable insertion ordering of |lb| descending into th[] */
static void order(qf_t *s)
{
    for (int j = 0; j < s->r; j++) {
        double lj = fabs(s->lb[j]);
        int k = j - 1;
        while (k >= 0 && lj > fabs(s->lb[s->th[k]])) { s->th[k + 1] = s->th[k]; k--; }
        s->th[k + 1] = j;
    }
    s->sorted = 1;
}

/* Chernoff-type bound on the tail probability; cut-off point returned in *cx */
static double errbd(qf_t *s, double u, double *cx)
{
    tick(s);
    double xconst = u * s->sigsq, sum1 = u * xconst;
    u = 2.0 * u;
    for (int j = s->r - 1; j >= 0; j--) {
        double nj = s->n[j], lj = s->lb[j], ncj = s->nc[j];
        double x = u * lj, y = 1.0 - x;
        xconst += lj * (ncj / y + nj) / y;
        sum1 += ncj * sq(x / y) + nj * (sq(x) / y + log1(-x, 0));
    }
    *cx = xconst;
    return exp1(-0.5 * sum1);
}

/* cut-off so that P(qf > ctff) < accx (upn > 0) or P(qf < ctff) < accx (upn < 0) */
static double ctff(qf_t *s, double accx, double *upn)
{
    double u2 = *upn, u1 = 0.0, c1 = s->mean, c2 = 0.0, xc;
    double rb = 2.0 * ((u2 > 0.0) ? s->lmax : s->lmin);
    double u = u2 / (1.0 + u2 * rb);
    while (errbd(s, u, &c2) > accx) {
        if (s->aborted) return c2;
        u1 = u2; c1 = c2; u2 = 2.0 * u2;
        u = u2 / (1.0 + u2 * rb);
    }
    u = (c1 - s->mean) / (c2 - s->mean);
    while (u < 0.9) {
        if (s->aborted) return c2;
        u = (u1 + u2) / 2.0;
        if (errbd(s, u / (1.0 + u * rb), &xc) > accx) { u1 = u; c1 = xc; }
        else { u2 = u; c2 = xc; }
        u = (c1 - s->mean) / (c2 - s->mean);
    }
    *upn = u2;
    return c2;
}

/* bound on the integration error caused by truncating at u */
static double truncation(qf_t *s, double u, double tausq)
{
    tick(s);
    double sum1 = 0.0, prod2 = 0.0, prod3 = 0.0;
    int ns = 0;
    double sum2 = (s->sigsq + tausq) * sq(u), prod1 = 2.0 * sum2;
    u = 2.0 * u;
    for (int j = 0; j < s->r; j++) {
        double lj = s->lb[j], ncj = s->nc[j]; int nj = s->n[j];
        double x = sq(u * lj);
        sum1 += ncj * x / (1.0 + x);
        if (x > 1.0) { prod2 += nj * log(x); prod3 += nj * log1(x, 1); ns += nj; }
        else prod1 += nj * log1(x, 1);
    }
    sum1 *= 0.5; prod2 += prod1; prod3 += prod1;
    double x = exp1(-sum1 - 0.25 * prod2) / QF_PI;
    double y = exp1(-sum1 - 0.25 * prod3) / QF_PI;
    double err1 = (ns == 0) ? 1.0 : x * 2.0 / ns;
    double err2 = (prod3 > 1.0) ? 2.5 * y : 1.0;
    if (err2 < err1) err1 = err2;
    x = 0.5 * sum2;
    err2 = (x <= y) ? 1.0 : y / x;
    return (err1 < err2) ? err1 : err2;
}

/* u such that truncation(u) < accx and truncation(u/1.2) > accx */
static void findu(qf_t *s, double *utx, double accx)
{
    static const double divis[4] = {2.0, 1.4, 1.2, 1.1};
    double ut = *utx, u = ut / 4.0;
    if (truncation(s, u, 0.0) > accx) {
        for (u = ut; truncation(s, u, 0.0) > accx; u = ut) { if (s->aborted) return; ut *= 4.0; }
    } else {
        ut = u;
        for (u = u / 4.0; truncation(s, u, 0.0) <= accx; u = u / 4.0) { if (s->aborted) return; ut = u; }
    }
    for (int i = 0; i < 4; i++) { u = ut / divis[i]; if (truncation(s, u, 0.0) <= accx) ut = u; }
    *utx = ut;
}

/* nterm+1 terms of the inversion integral at step interv; !mainx multiplies the integrand by
 * 1 - exp(-0.5 tausq u^2).  Terms are accumulated from k = nterm DOWN to 0. */
static void integrate(qf_t *s, int nterm, double interv, double tausq, int mainx)
{
    double inpi = interv / QF_PI;
    for (int k = nterm; k >= 0; k--) {
        double u = (k + 0.5) * interv;
        double sum1 = -2.0 * u * s->c, sum2 = fabs(sum1), sum3 = -0.5 * s->sigsq * sq(u);
        for (int j = s->r - 1; j >= 0; j--) {
            int nj = s->n[j];
            double x = 2.0 * s->lb[j] * u, y = sq(x);
            sum3 -= 0.25 * nj * log1(y, 1);
            y = s->nc[j] * x / (1.0 + y);
            double z = nj * atan(x) + y;
            sum1 += z; sum2 += fabs(z); sum3 -= 0.5 * x * y;
        }
        double x = inpi * exp1(sum3) / u;
        if (!mainx) x *= (1.0 - exp1(-0.5 * tausq * sq(u)));
        s->intl += sin(0.5 * sum1) * x;
        s->ersm += 0.5 * sum2 * x;
    }
}

/* coefficient of tausq in the error when the convergence factor exp(-0.5 tausq u^2) is used */
static double cfe(qf_t *s, double x)
{
    tick(s);
    if (!s->sorted) order(s);
    double axl = fabs(x), sxl = (x > 0.0) ? 1.0 : -1.0, sum1 = 0.0;
    for (int j = s->r - 1; j >= 0; j--) {
        int t = s->th[j];
        if (s->lb[t] * sxl > 0.0) {
            double lj = fabs(s->lb[t]);
            double axl1 = axl - lj * (s->n[t] + s->nc[t]), axl2 = lj / QF_LOG28;
            if (axl1 > axl2) axl = axl1;
            else {
                if (axl > axl2) axl = axl2;
                sum1 = (axl - axl1) / lj;
                for (int k = j - 1; k >= 0; k--) sum1 += (s->n[s->th[k]] + s->nc[s->th[k]]);
                break;
            }
        }
    }
    if (sum1 > 100.0) { s->fail = 1; return 1.0; }
    return pow(2.0, sum1 / 4.0) / (QF_PI * sq(axl));
}

/*
 * P(sum_j lb[j] chi2(n[j], nc[j]) + sigma N(0,1) < c).  trace[7] as in the published routine:
 * 0 abs sum, 1 total terms, 2 integrations, 3 main interval, 4 truncation point,
 * 5 sd of convergence factor, 6 cycles.  Returns qfval; *ifault in {0,1,2,3,4,5}.
 */
double crm_oracle_qfc(const double *lb, const double *nc, const int *n, int r, double sigma, double c,
                      int lim, double acc, double *trace, int *ifault)
{
    static const int rats[4] = {1, 2, 4, 8};
    qf_t st; qf_t *s = &st;
    double qfval = -1.0, acc1 = acc, xlim = (double)lim;
    double utx, tausq, sd, intv, intv1, x, up, un, d1, d2, almx, xnt, xntm;
    int nt, ntm;
    for (int j = 0; j < 7; j++) trace[j] = 0.0;
    *ifault = 0;
    s->r = r; s->lim = lim; s->c = c; s->lb = lb; s->nc = nc; s->n = n;
    s->count = 0; s->intl = 0.0; s->ersm = 0.0; s->sorted = 0; s->fail = 0; s->aborted = 0;
    s->th = (int *)malloc((r > 0 ? r : 1) * sizeof(int));
    if (!s->th) { *ifault = 5; return qfval; }

    s->sigsq = sq(sigma); sd = s->sigsq; s->lmax = 0.0; s->lmin = 0.0; s->mean = 0.0;
    for (int j = 0; j < r; j++) {
        int nj = n[j]; double lj = lb[j], ncj = nc[j];
        if (nj < 0 || ncj < 0.0) { *ifault = 3; goto done; }
        sd += sq(lj) * (2 * nj + 4.0 * ncj);
        s->mean += lj * (nj + ncj);
        if (s->lmax < lj) s->lmax = lj; else if (s->lmin > lj) s->lmin = lj;
    }
    if (sd == 0.0) { qfval = (c > 0.0) ? 1.0 : 0.0; goto done; }
    if (s->lmin == 0.0 && s->lmax == 0.0 && sigma == 0.0) { *ifault = 3; goto done; }
    sd = sqrt(sd);
    almx = (s->lmax < -s->lmin) ? -s->lmin : s->lmax;

    utx = 16.0 / sd; up = 4.5 / sd; un = -up;
    findu(s, &utx, 0.5 * acc1);
    if (s->aborted) goto aborted;
    if (c != 0.0 && almx > 0.07 * sd) {
        tausq = 0.25 * acc1 / cfe(s, c);
        if (s->aborted) goto aborted;
        if (s->fail) s->fail = 0;
        else {
            double tr = truncation(s, utx, tausq);
            if (s->aborted) goto aborted;
            if (tr < 0.2 * acc1) {
                s->sigsq += tausq;
                findu(s, &utx, 0.25 * acc1);
                if (s->aborted) goto aborted;
                trace[5] = sqrt(tausq);
            }
        }
    }
    trace[4] = utx; acc1 = 0.5 * acc1;

    for (;;) {
        d1 = ctff(s, acc1, &up) - c;
        if (s->aborted) goto aborted;
        if (d1 < 0.0) { qfval = 1.0; goto done; }
        d2 = c - ctff(s, acc1, &un);
        if (s->aborted) goto aborted;
        if (d2 < 0.0) { qfval = 0.0; goto done; }
        intv = 2.0 * QF_PI / ((d1 > d2) ? d1 : d2);
        xnt = utx / intv; xntm = 3.0 / sqrt(acc1);
        if (xnt > xntm * 1.5) {
            if (xntm > xlim) { *ifault = 1; goto done; }
            ntm = (int)floor(xntm + 0.5);
            intv1 = utx / ntm; x = 2.0 * QF_PI / intv1;
            if (x <= fabs(c)) break;
            {
                double e1 = cfe(s, c - x); if (s->aborted) goto aborted;
                double e2 = cfe(s, c + x); if (s->aborted) goto aborted;
                tausq = 0.33 * acc1 / (1.1 * (e1 + e2));
            }
            if (s->fail) break;
            acc1 = 0.67 * acc1;
            integrate(s, ntm, intv1, tausq, 0);
            xlim -= xntm; s->sigsq += tausq;
            trace[2] += 1; trace[1] += ntm + 1;
            findu(s, &utx, 0.25 * acc1);
            if (s->aborted) goto aborted;
            acc1 = 0.75 * acc1;
            continue;
        }
        break;
    }
    trace[3] = intv;
    if (xnt > xlim) { *ifault = 1; goto done; }
    nt = (int)floor(xnt + 0.5);
    integrate(s, nt, intv, 0.0, 1);
    trace[2] += 1; trace[1] += nt + 1;
    qfval = 0.5 - s->intl;
    trace[0] = s->ersm;
    up = s->ersm; x = up + acc / 10.0;
    for (int j = 0; j < 4; j++) if (rats[j] * x == rats[j] * up) *ifault = 2;
    goto done;

aborted:
    *ifault = 4;
done:
    free(s->th);
    trace[6] = (double)s->count;
    return qfval;
}
