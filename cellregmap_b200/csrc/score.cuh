// K3: selection of the best rho1 per SNP, grouping by rho1, and the fused score statistic / null-distribution
// weights.
//
// Replaces the tail of the per-SNP loop of the reference: best-rho1 rule (cellregmap/_cellregmap.py:354-357,
// 366-369), QSCov / PMat / ScoreStatistic (cellregmap/_math.py:40-128; call sites _cellregmap.py:379-416) and the
// eigenvalue step of chiscore.davies_pvalue (_cellregmap.py:435).  Works on Grams of Z = [y | X | g.E0]:
//   Z' K0^-1 Z = (Z'Z - Zr' diag(w) Zr) / v1,   w_i = v0 S_i / (v0 S_i + v1),   Zr = Q0' Z,
//   A = X'K0^-1 X,  t = GE'K0^-1 y - GE'K0^-1 X A^+ X'K0^-1 y,  Q = t't / 2,
//   M = (GE'K0^-1 GE - GE'K0^-1 X A^+ X'K0^-1 GE) / 2,  lambda = eig(M).
#pragma once
#include "common.cuh"
#include "smallmat.cuh"
#include "args.cuh"

namespace crm {

// ---- best rho1 per SNP: strict '>' over the ascending grid, first maximum wins ----
__global__ void crm_select_kernel(const double* lml, const double* delta, const double* scale, int p, int R,
                                  int* rho_idx, double* best_lml, double* v0, double* v1) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p) return;
    double best = -INFINITY; int bi = 0;
    for (int r = 0; r < R; r++) { const double l = lml[(long long)s * R + r]; if (l > best) { best = l; bi = r; } }
    rho_idx[s] = bi;
    best_lml[s] = best;
    const double d = delta[(long long)s * R + bi], sc = scale[(long long)s * R + bi];
    v0[s] = sc * (1.0 - d);
    v1[s] = sc * d;
}

// ---- stable counting sort of SNPs by rho index (single CTA): perm[pos] = s, offsets[R+1] ----
__global__ void __launch_bounds__(1024) crm_group_kernel(const int* rho_idx, int p, int R, int* perm, int* offsets) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { base = 0; offsets[0] = 0; }
    __syncthreads();
    for (int r = 0; r < R; r++) {
        for (int start = 0; start < p; start += 1024) {
            const int s = start + tid;
            const int flag = (s < p && rho_idx[s] == r) ? 1 : 0;
            int incl = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int woff = 0, total = 0;
            for (int w = 0; w < 32; w++) { const int v = warp_tot[w]; if (w < warp) woff += v; total += v; }
            if (flag) perm[base + woff + incl - 1] = s;
            __syncthreads();
            if (tid == 0) base += total;
            __syncthreads();
        }
        if (tid == 0) offsets[r + 1] = base;
    }
}

// ---- gather + transpose:  out[a][q] = C[row(q)][a],  q = pos * kcols + jj,  row = perm[pos] * kexp + joff + jj ----
__global__ void crm_gather_transpose_kernel(const double* C, long long ldc, const int* perm, int kexp, int joff, int kcols,
                                            long long nq, int na, double* out, long long ldo) {
    __shared__ double tile[32][33];
    const long long q0 = (long long)blockIdx.x * 32;
    const int a0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long q = q0 + r;
        const int a = a0 + threadIdx.x;
        double v = 0.0;
        if (q < nq && a < na) {
            const long long pos = q / kcols; const int jj = (int)(q - pos * kcols);
            const long long src = (perm ? (long long)perm[pos] : pos) * kexp + joff + jj;
            v = C[src * ldc + a];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int a = a0 + r;
        const long long q = q0 + threadIdx.x;
        if (a < na && q < nq) out[(long long)a * ldo + q] = tile[threadIdx.x][r];
    }
}


constexpr int SCORE_THREADS = 256;
constexpr int SCORE_CHUNK = 32;
constexpr int SCORE_MAXE = 8;


__global__ void __launch_bounds__(SCORE_THREADS) crm_score_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) double ssm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = blockIdx.x;
    const int s = a.perm[pos];
    const int rho = a.rho_idx[s];
    const int P = a.c + 1, C = a.c;
    const int k = a.k, NZ = 1 + P + k, mp = a.mp, m = a.m;
    const double v0 = a.v0[s], v1 = a.v1[s];
    // shared layout
    double* G = ssm;                       // NZ x NZ Gram (first rotated-weighted, then K0^-1 Gram)
    double* Zt = G + NZ * NZ;              // NZ x (CHUNK+1) staged rotated columns
    double* wv = Zt + NZ * (SCORE_CHUNK + 1);   // CHUNK weights
    double* Mm = wv + SCORE_CHUNK;         // k x k
    double* sol = Mm + k * k;              // P x (1 + k): A^+ [X'Ky | X'K GE]
    double* tv = sol + P * (1 + k);        // 2 k
    double* Aw = tv + 2 * k;               // P x P
    double* Vw = Aw + P * P;               // P x P
    double* yw = Vw + P * P;               // P x (1 + k)
    const double* S = a.S + (long long)rho * mp;
    const double* yr = a.yr + (long long)rho * mp;
    const double* Wr = a.Wr + (long long)rho * C * mp;
    const double* gr = a.gr + (long long)s * a.gr_ld + (long long)rho * mp;
    const double* GEr = a.GEr + (long long)pos * k * mp;

    // ---- 1. rotated weighted Gram  RG = sum_i w_i Zr_i Zr_i' ----
    const int npairs = NZ * (NZ + 1) / 2;
    double acc[SCORE_MAXE];                // up to SCORE_MAXE Gram entries per thread (NZ <= 63)
    int ea[SCORE_MAXE], eb[SCORE_MAXE];
#pragma unroll
    for (int u = 0; u < SCORE_MAXE; u++) {
        int e = tid + u * SCORE_THREADS; ea[u] = -1; eb[u] = 0; acc[u] = 0.0;
        if (e < npairs) { int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5); while (r * (r + 1) / 2 > e) r--; while ((r + 1) * (r + 2) / 2 <= e) r++; ea[u] = r; eb[u] = e - r * (r + 1) / 2; }
    }
    for (int i0 = 0; i0 < m; i0 += SCORE_CHUNK) {
        for (int idx = tid; idx < NZ * SCORE_CHUNK; idx += SCORE_THREADS) {
            const int col = idx / SCORE_CHUNK, ii = idx - col * SCORE_CHUNK, i = i0 + ii;
            double v = 0.0;
            if (i < m) {
                if (col == 0) v = yr[i];
                else if (col <= C) v = Wr[(long long)(col - 1) * mp + i];
                else if (col == P) v = gr[i];
                else v = GEr[(long long)(col - 1 - P) * mp + i];
            }
            Zt[col * (SCORE_CHUNK + 1) + ii] = v;
        }
        if (tid < SCORE_CHUNK) { const int i = i0 + tid; double w = 0.0; if (i < m) { const double vs = v0 * S[i]; w = vs / (vs + v1); } wv[tid] = w; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < SCORE_MAXE; u++) {
            if (ea[u] >= 0) {
                const double* za = Zt + ea[u] * (SCORE_CHUNK + 1);
                const double* zb = Zt + eb[u] * (SCORE_CHUNK + 1);
                double sacc = acc[u];
#pragma unroll 8
                for (int ii = 0; ii < SCORE_CHUNK; ii++) sacc = fma(za[ii] * wv[ii], zb[ii], sacc);
                acc[u] = sacc;
            }
        }
        __syncthreads();
    }
    // ---- 2. K0^-1 Gram: (Z'Z - RG) / v1 ----
    const double* row0 = a.rot + (long long)s * a.kexp * a.rot_ld;   // row of g
    const double* sq = a.sq + (long long)s * a.sq_ld;
#pragma unroll
    for (int u = 0; u < SCORE_MAXE; u++) {
        if (ea[u] < 0) continue;
        const int ia = ea[u], ib = eb[u];   // ia >= ib ; columns: 0 = y, 1..C = W, P = g, P+1.. = GE_j
        double zz;
        if (ia == 0) zz = a.stats[0];
        else if (ia <= C) zz = (ib == 0) ? a.stats[ia] : a.stats[1 + C + (ia - 1) * C + (ib - 1)];
        else if (ia == P) zz = (ib == 0) ? row0[a.col_y] : (ib <= C ? row0[a.col_W + ib - 1] : sq[0]);
        else {
            const int j = ia - P - 1;       // GE column j (0-based)
            const double* rowj = row0 + (long long)(1 + j) * a.rot_ld;
            if (ib == 0) zz = rowj[a.col_y];
            else if (ib <= C) zz = rowj[a.col_W + ib - 1];
            else if (ib == P) zz = sq[1 + j];
            else zz = sq[1 + k + pair_index(j, ib - P - 1)];
        }
        const double v = (zz - acc[u]) / v1;
        G[ia * NZ + ib] = v; G[ib * NZ + ia] = v;
    }
    __syncthreads();
    // ---- 3. A^+ [X'Ky | X'K GE]: pseudo-inverse of the P x P block through its eigendecomposition (warp 0, Jacobi in
    //         shared memory) with the relative cut-off eps * P of numpy's lstsq(rcond=None), reference _math.py:33-37 ----
    if (warp == 0) {
        for (int e = lane; e < P * P; e += 32) { const int i = e / P, j = e - i * P; Aw[e] = G[(1 + i) * NZ + (1 + j)]; }
        __syncwarp();
        warp_jacobi_vec(Aw, Vw, P, lane);
    }
    __syncthreads();
    {
        double lmax = 0.0;
        for (int r = 0; r < P; r++) lmax = fmax(lmax, fabs(Aw[r * P + r]));
        const double cut = CRM_EPS_TINY * P * lmax;
        for (int e = tid; e < P * (1 + k); e += SCORE_THREADS) {      // yw[r][j] = (V' b_j)_r / lambda_r
            const int r = e / (1 + k), j = e - r * (1 + k);
            const int col = (j == 0) ? 0 : P + j;
            const double l = Aw[r * P + r];
            double v = 0.0;
            if (fabs(l) > cut && fabs(l) > 0.0) {
                for (int t = 0; t < P; t++) v += Vw[t * P + r] * G[(1 + t) * NZ + col];
                v /= l;
            }
            yw[e] = v;
        }
        __syncthreads();
        for (int e = tid; e < P * (1 + k); e += SCORE_THREADS) {      // sol[i][j] = sum_r V[i][r] yw[r][j]
            const int i = e / (1 + k), j = e - i * (1 + k);
            double v = 0.0;
            for (int r = 0; r < P; r++) v += Vw[i * P + r] * yw[r * (1 + k) + j];
            sol[e] = v;
        }
    }
    __syncthreads();
    // t_j and M
    for (int j = tid; j < k; j += SCORE_THREADS) {
        double t = G[(P + 1 + j) * NZ + 0];
        for (int i = 0; i < P; i++) t -= G[(P + 1 + j) * NZ + (1 + i)] * sol[i * (1 + k) + 0];
        tv[j] = t;
    }
    for (int e = tid; e < k * k; e += SCORE_THREADS) {
        const int j = e / k, l = e - j * k;
        double v = G[(P + 1 + j) * NZ + (P + 1 + l)];
        for (int i = 0; i < P; i++) v -= G[(P + 1 + j) * NZ + (1 + i)] * sol[i * (1 + k) + 1 + l];
        Mm[e] = 0.5 * v;
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += SCORE_THREADS) {   // symmetrise away the round-off asymmetry
        const int j = e / k, l = e - j * k;
        if (j < l) { const double v = 0.5 * (Mm[j * k + l] + Mm[l * k + j]); Mm[j * k + l] = v; Mm[l * k + j] = v; }
    }
    __syncthreads();
    if (a.Mout) for (int e = tid; e < k * k; e += SCORE_THREADS) a.Mout[(long long)s * k * k + e] = Mm[e];
    if (warp != 0) return;
    // ---- 4. Q, eigenvalues of M by cyclic Jacobi (warp 0), filter ----
    double qs = 0.0;
    for (int j = lane; j < k; j += 32) qs += tv[j] * tv[j];
    qs = 0.5 * warp_sum(qs);
    warp_jacobi_vec(Mm, nullptr, k, lane);      // eigenvalues of M on its diagonal
    // descending rank sort of the diagonal into tv[], then the chiscore filter: lambda > mean(lambda >= 0) / 1e5
    for (int j = lane; j < k; j += 32) {
        const double lj = Mm[j * k + j];
        int rank = 0;
        for (int l = 0; l < k; l++) { const double ll = Mm[l * k + l]; if (ll > lj || (ll == lj && l < j)) rank++; }
        sol[rank] = lj;   // sol is free now (k <= P*(1+k))
    }
    __syncwarp();
    double psum = 0.0; int pcnt = 0;
    for (int j = lane; j < k; j += 32) { if (sol[j] >= 0.0) { psum += sol[j]; pcnt++; } }
    psum = warp_sum(psum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pcnt += __shfl_xor_sync(0xffffffffu, pcnt, o);
    int nl = 0;
    if (pcnt > 0) {
        const double thr = (psum / pcnt) / 100000.0;
        for (int j = 0; j < k; j++) if (sol[j] > thr) nl++;   // descending, so the kept ones are the first nl
    }
    for (int j = lane; j < k; j += 32) a.lam[(long long)s * a.lam_ld + j] = sol[j];
    if (lane == 0) { a.Q[s] = qs; a.nlam[s] = nl; a.flags[s] = (pcnt == 0 || nl == 0) ? 1 : 0; }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Warp-per-SNP version of the kernel above for NZ <= 8 NB columns (NB = 3 covers one covariate + 20 contexts): the weighted Gram
// runs on the FP64 tensor cores (mma.sync m8n8k4: lane (g, t) feeds row 4 s + t of the columns 8 b + g of Z, the NB (NB + 1) / 2
// lower block pairs accumulate in registers), the small algebra and both Jacobi eigensolvers stay inside the warp.  The CTA-per-SNP
// kernel spends most of its life with one warp in the 20 x 20 Jacobi while seven have exited (ncu: 17 % of the warp slots active,
// FP64 pipe 7 %, profiles/r02_ncu_score_fit.txt); here every warp of a CTA owns a SNP from start to end.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int SCOREW_WARPS = 4;
constexpr int SCOREW_WCHUNK = 128;       // rows whose weights are staged per pass

__host__ __device__ inline size_t scorew_warp_doubles(int NZ, int P, int k) {
    return (size_t)NZ * NZ + SCOREW_WCHUNK + (size_t)k * k + 2 * (size_t)P * (1 + k) + 2 * (size_t)k + 2 * (size_t)P * P + 2;
}

template <int NB>
__global__ void __launch_bounds__(SCOREW_WARPS * 32) crm_score_warp_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) double ssm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pos = blockIdx.x * SCOREW_WARPS + warp;
    if (pos >= a.p) return;
    const int s = a.perm[pos];
    const int rho = a.rho_idx[s];
    const int P = a.c + 1, C = a.c;
    const int k = a.k, NZ = 1 + P + k, mp = a.mp, m = a.m;
    const double v0 = a.v0[s], v1 = a.v1[s];
    double* G = ssm + (size_t)warp * scorew_warp_doubles(NZ, P, k);
    double* wch = G + NZ * NZ;
    double* Mm = wch + SCOREW_WCHUNK;
    double* sol = Mm + k * k;
    double* tv = sol + P * (1 + k);
    double* Aw = tv + 2 * k;
    double* Vw = Aw + P * P;
    double* yw = Vw + P * P;
    const double* S = a.S + (long long)rho * mp;
    const double* yr = a.yr + (long long)rho * mp;
    const double* Wr = a.Wr + (long long)rho * C * mp;
    const double* gr = a.gr + (long long)s * a.gr_ld + (long long)rho * mp;
    const double* GEr = a.GEr + (long long)pos * k * mp;

    // ---- 1. rotated weighted Gram  RG = sum_i w_i Zr_i Zr_i'  on the tensor cores ----
    const int g = lane >> 2, t = lane & 3;
    const double* colp[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int col = 8 * b + g;
        colp[b] = (col == 0) ? yr : (col <= C) ? Wr + (long long)(col - 1) * mp : (col == P) ? gr : (col < NZ) ? GEr + (long long)(col - 1 - P) * mp : nullptr;
    }
    double acc[NB * (NB + 1) / 2][2];
#pragma unroll
    for (int e = 0; e < NB * (NB + 1) / 2; e++) { acc[e][0] = 0.0; acc[e][1] = 0.0; }
    for (int i0 = 0; i0 < m; i0 += SCOREW_WCHUNK) {
        for (int ii = lane; ii < SCOREW_WCHUNK; ii += 32) {
            const int i = i0 + ii;
            double w = 0.0;
            if (i < m) { const double vs = v0 * S[i]; w = vs / (vs + v1); }
            wch[ii] = w;
        }
        __syncwarp();
        const int steps = min(SCOREW_WCHUNK, m - i0 + 3) / 4;
#pragma unroll 4
        for (int st = 0; st < steps; st++) {
            const int i = i0 + 4 * st + t;
            const bool ok = i < m;
            const double w = wch[4 * st + t];
            double z[NB], wz[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) { z[b] = (ok && colp[b]) ? colp[b][i] : 0.0; wz[b] = w * z[b]; }
            int e = 0;
#pragma unroll
            for (int bi = 0; bi < NB; bi++)
#pragma unroll
                for (int bj = 0; bj <= bi; bj++, e++) dmma884(acc[e][0], acc[e][1], wz[bi], z[bj]);
        }
        __syncwarp();
    }
    {
        int e = 0;
#pragma unroll
        for (int bi = 0; bi < NB; bi++)
#pragma unroll
            for (int bj = 0; bj <= bi; bj++, e++) {
                // lower triangle only (row >= column), mirrored: in a diagonal block the entry (c, r) is also some other lane's (r', c'),
                // computed with the roles of the two factors swapped -- one writer per location keeps the result deterministic
                const int r = 8 * bi + g, c0 = 8 * bj + 2 * t;
                if (r < NZ) {
                    if (c0 <= r) { G[r * NZ + c0] = acc[e][0]; G[c0 * NZ + r] = acc[e][0]; }
                    if (c0 + 1 <= r) { G[r * NZ + c0 + 1] = acc[e][1]; G[(c0 + 1) * NZ + r] = acc[e][1]; }
                }
            }
    }
    __syncwarp();
    // ---- 2. K0^-1 Gram: (Z'Z - RG) / v1 ----
    const double* row0 = a.rot + (long long)s * a.kexp * a.rot_ld;   // row of g
    const double* sq = a.sq + (long long)s * a.sq_ld;
    const int npairs = NZ * (NZ + 1) / 2;
    for (int e = lane; e < npairs; e += 32) {
        int ia = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (ia * (ia + 1) / 2 > e) ia--;
        while ((ia + 1) * (ia + 2) / 2 <= e) ia++;
        const int ib = e - ia * (ia + 1) / 2;      // ia >= ib ; columns: 0 = y, 1..C = W, P = g, P+1.. = GE_j
        double zz;
        if (ia == 0) zz = a.stats[0];
        else if (ia <= C) zz = (ib == 0) ? a.stats[ia] : a.stats[1 + C + (ia - 1) * C + (ib - 1)];
        else if (ia == P) zz = (ib == 0) ? row0[a.col_y] : (ib <= C ? row0[a.col_W + ib - 1] : sq[0]);
        else {
            const int j = ia - P - 1;
            const double* rowj = row0 + (long long)(1 + j) * a.rot_ld;
            if (ib == 0) zz = rowj[a.col_y];
            else if (ib <= C) zz = rowj[a.col_W + ib - 1];
            else if (ib == P) zz = sq[1 + j];
            else zz = sq[1 + k + pair_index(j, ib - P - 1)];
        }
        const double v = (zz - G[ia * NZ + ib]) / v1;
        G[ia * NZ + ib] = v; G[ib * NZ + ia] = v;
    }
    __syncwarp();
    // ---- 3. A^+ [X'Ky | X'K GE]: pseudo-inverse of the P x P block (Jacobi, lstsq(rcond=None) cut-off, reference _math.py:33-37) ----
    for (int e = lane; e < P * P; e += 32) { const int i = e / P, j = e - i * P; Aw[e] = G[(1 + i) * NZ + (1 + j)]; }
    __syncwarp();
    warp_jacobi_vec(Aw, Vw, P, lane);
    {
        double lmax = 0.0;
        for (int r = 0; r < P; r++) lmax = fmax(lmax, fabs(Aw[r * P + r]));
        const double cut = CRM_EPS_TINY * P * lmax;
        for (int e = lane; e < P * (1 + k); e += 32) {
            const int r = e / (1 + k), j = e - r * (1 + k);
            const int col = (j == 0) ? 0 : P + j;
            const double l = Aw[r * P + r];
            double v = 0.0;
            if (fabs(l) > cut && fabs(l) > 0.0) {
                for (int tt = 0; tt < P; tt++) v += Vw[tt * P + r] * G[(1 + tt) * NZ + col];
                v /= l;
            }
            yw[e] = v;
        }
        __syncwarp();
        for (int e = lane; e < P * (1 + k); e += 32) {
            const int i = e / (1 + k), j = e - i * (1 + k);
            double v = 0.0;
            for (int r = 0; r < P; r++) v += Vw[i * P + r] * yw[r * (1 + k) + j];
            sol[e] = v;
        }
    }
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
        double tj = G[(P + 1 + j) * NZ + 0];
        for (int i = 0; i < P; i++) tj -= G[(P + 1 + j) * NZ + (1 + i)] * sol[i * (1 + k) + 0];
        tv[j] = tj;
    }
    for (int e = lane; e < k * k; e += 32) {
        const int j = e / k, l = e - j * k;
        double v = G[(P + 1 + j) * NZ + (P + 1 + l)];
        for (int i = 0; i < P; i++) v -= G[(P + 1 + j) * NZ + (1 + i)] * sol[i * (1 + k) + 1 + l];
        Mm[e] = 0.5 * v;
    }
    __syncwarp();
    for (int e = lane; e < k * k; e += 32) {   // symmetrise away the round-off asymmetry
        const int j = e / k, l = e - j * k;
        if (j < l) { const double v = 0.5 * (Mm[j * k + l] + Mm[l * k + j]); Mm[j * k + l] = v; Mm[l * k + j] = v; }
    }
    __syncwarp();
    if (a.Mout) for (int e = lane; e < k * k; e += 32) a.Mout[(long long)s * k * k + e] = Mm[e];
    // ---- 4. Q, eigenvalues of M by cyclic Jacobi, filter ----
    double qs = 0.0;
    for (int j = lane; j < k; j += 32) qs += tv[j] * tv[j];
    qs = 0.5 * warp_sum(qs);
    __syncwarp();
    warp_jacobi_vec(Mm, nullptr, k, lane);
    for (int j = lane; j < k; j += 32) {
        const double lj = Mm[j * k + j];
        int rank = 0;
        for (int l = 0; l < k; l++) { const double ll = Mm[l * k + l]; if (ll > lj || (ll == lj && l < j)) rank++; }
        sol[rank] = lj;
    }
    __syncwarp();
    double psum = 0.0; int pcnt = 0;
    for (int j = lane; j < k; j += 32) { if (sol[j] >= 0.0) { psum += sol[j]; pcnt++; } }
    psum = warp_sum(psum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pcnt += __shfl_xor_sync(0xffffffffu, pcnt, o);
    int nl = 0;
    if (pcnt > 0) {
        const double thr = (psum / pcnt) / 100000.0;
        for (int j = 0; j < k; j++) if (sol[j] > thr) nl++;
    }
    for (int j = lane; j < k; j += 32) a.lam[(long long)s * a.lam_ld + j] = sol[j];
    if (lane == 0) { a.Q[s] = qs; a.nlam[s] = nl; a.flags[s] = (pcnt == 0 || nl == 0) ? 1 : 0; }
}

}  // namespace crm
