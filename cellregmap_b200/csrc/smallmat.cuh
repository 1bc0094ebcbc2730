// Small dense symmetric helpers evaluated redundantly by every lane (compile-time sizes -> registers).
#pragma once
#include "common.cuh"

namespace crm {

// Cyclic Jacobi eigendecomposition of a symmetric P x P matrix A (destroyed; eigenvalues end on its diagonal),
// eigenvectors in the columns of V.  Exactly-zero off-diagonal entries are skipped, so decoupled (masked)
// coordinates keep their index.
template <int P>
__device__ __forceinline__ void jacobi_eig(double (&A)[P][P], double (&V)[P][P]) {
#pragma unroll
    for (int i = 0; i < P; i++)
#pragma unroll
        for (int j = 0; j < P; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    if (P == 1) return;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < P; j++) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;      // |off| <= 1e-15 |diag|: below the round-off of the rotations
#pragma unroll
        for (int p = 0; p < P - 1; p++) {
#pragma unroll
            for (int q = p + 1; q < P; q++) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                if (fabs(apq) <= 1e-17 * sqrt(fabs(A[p][p] * A[q][q]))) { A[p][q] = 0.0; A[q][p] = 0.0; continue; }   // negligible coupling
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                A[p][p] -= tt * apq;
                A[q][q] += tt * apq;
                A[p][q] = 0.0;
                A[q][p] = 0.0;
#pragma unroll
                for (int r = 0; r < P; r++) {
                    if (r != p && r != q) {
                        const double arp = A[r][p], arq = A[r][q];
                        A[r][p] = c * arp - s * arq; A[p][r] = A[r][p];
                        A[r][q] = s * arp + c * arq; A[q][r] = A[r][q];
                    }
                    const double vrp = V[r][p], vrq = V[r][q];
                    V[r][p] = c * vrp - s * vrq;
                    V[r][q] = s * vrp + c * vrq;
                }
            }
        }
    }
}

// Minimum-norm least-squares solve of the symmetric PSD system A x = b with relative cut-off rcond on the
// eigenvalues (numpy.linalg.lstsq semantics for a symmetric matrix); returns sum of log(eigenvalue) over the
// unmasked coordinates in *logdet and whether all of them are positive in *posdef.  `mask` bit i set = coordinate i
// is a dropped design direction (A row/col i and b[i] are zero) and is ignored.
template <int P>
__device__ __forceinline__ void sym_pinv_solve(const double (&Ain)[P][P], const double (&b)[P], unsigned mask, double rcond,
                                               double (&x)[P], double* logdet, bool* posdef) {
    double A[P][P], V[P][P];
#pragma unroll
    for (int i = 0; i < P; i++)
#pragma unroll
        for (int j = 0; j < P; j++) A[i][j] = Ain[i][j];
    jacobi_eig<P>(A, V);
    double lmax = 0.0, ld = 0.0;
    bool pd = true;
#pragma unroll
    for (int i = 0; i < P; i++) {
        if (mask >> i & 1u) continue;
        const double l = A[i][i];
        lmax = fmax(lmax, fabs(l));
        if (l > 0.0) ld += log(l); else pd = false;
    }
#pragma unroll
    for (int i = 0; i < P; i++) x[i] = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++) {
        if (mask >> i & 1u) continue;
        const double l = A[i][i];
        if (fabs(l) > rcond * lmax && fabs(l) > 0.0) {
            double vb = 0.0;
#pragma unroll
            for (int r = 0; r < P; r++) vb += V[r][i] * b[r];
            vb /= l;
#pragma unroll
            for (int r = 0; r < P; r++) x[r] += V[r][i] * vb;
        }
    }
    *logdet = ld;
    *posdef = pd;
}

// ---- warp-level dense helpers on shared-memory matrices (row-major, leading dimension ld) ----
// in-place Cholesky (lower) of the leading n x n block; returns the smallest pivot (<= 0: not positive definite)
__device__ inline double warp_cholesky(double* A, int n, int ld, int lane) {
    double minpiv = INFINITY;
    for (int j = 0; j < n; j++) {
        double s = 0.0;
        for (int t = lane; t < j; t += 32) s += A[j * ld + t] * A[j * ld + t];
        const double d = A[j * ld + j] - warp_sum(s);
        minpiv = fmin(minpiv, d);
        const double l = sqrt(fmax(d, 1e-300));
        __syncwarp();
        for (int i = j + 1 + lane; i < n; i += 32) {
            double v = A[i * ld + j];
            for (int t = 0; t < j; t++) v -= A[i * ld + t] * A[j * ld + t];
            A[i * ld + j] = v / l;
        }
        if (lane == 0) A[j * ld + j] = l;
        __syncwarp();
    }
    return minpiv;
}
// X (n x nrhs, ld ldx) <- L^-1 X, one lane per right-hand side
__device__ inline void warp_forward_solve(const double* L, int n, int ld, double* X, int nrhs, int ldx, int lane) {
    for (int c = lane; c < nrhs; c += 32)
        for (int i = 0; i < n; i++) {
            double v = X[i * ldx + c];
            for (int t = 0; t < i; t++) v -= L[i * ld + t] * X[t * ldx + c];
            X[i * ldx + c] = v / L[i * ld + i];
        }
    __syncwarp();
}
__device__ inline void warp_backward_solve(const double* L, int n, int ld, double* X, int nrhs, int ldx, int lane) {
    for (int c = lane; c < nrhs; c += 32)
        for (int i = n - 1; i >= 0; i--) {
            double v = X[i * ldx + c];
            for (int t = i + 1; t < n; t++) v -= L[t * ld + i] * X[t * ldx + c];
            X[i * ldx + c] = v / L[i * ld + i];
        }
    __syncwarp();
}
// Jacobi eigendecomposition of a symmetric n x n matrix in shared memory by one warp (n <= 64), parallel (round-robin
// tournament) ordering: each round rotates n/2 disjoint index pairs at once -- lane i computes the rotation of pair i, then all
// lanes apply the n/2 rotations to the columns (one row per lane) and to the rows (one column per lane).  Eigenvalues end on
// the diagonal; V (may be null) receives the eigenvectors in its columns.
__device__ inline void warp_jacobi_vec(double* A, double* V, int n, int lane) {
    if (V) for (int e = lane; e < n * n; e += 32) V[e] = (e / n == e % n) ? 1.0 : 0.0;
    __syncwarp();
    if (n < 2) return;
    const int ne = n + (n & 1);            // even number of players; index n (if present) is a bye
    const int npairs = ne / 2;             // <= 32
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0.0, dg = 0.0;
        for (int e = lane; e < n * n; e += 32) { const double v = A[e]; if (e / n == e % n) dg += v * v; else off += v * v; }
        off = warp_sum(off); dg = warp_sum(dg);
        if (off <= 1e-30 * dg || off == 0.0) break;      // |off| <= 1e-15 |diag|: below the round-off of the rotations
        for (int rd = 0; rd < ne - 1; rd++) {
            // pair of this lane in round rd (circle method: player ne-1 fixed, the others rotate)
            int p = -1, q = -1;
            double c = 1.0, sn = 0.0;
            if (lane < npairs) {
                int a0 = (lane == 0) ? ne - 1 : (rd + lane) % (ne - 1);
                int b0 = (lane == 0) ? rd % (ne - 1) : (rd - lane + ne - 1) % (ne - 1);
                if (a0 > b0) { const int t = a0; a0 = b0; b0 = t; }
                if (b0 < n) {
                    const double apq = A[a0 * n + b0], app = A[a0 * n + a0], aqq = A[b0 * n + b0];
                    if (apq != 0.0 && fabs(apq) > 1e-17 * sqrt(fabs(app * aqq))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt * tt + 1.0); sn = tt * c;
                        p = a0; q = b0;
                    }
                }
            }
            __syncwarp();
            // columns: A <- A J (and V <- V J), one row per lane
            for (int i = 0; i < npairs; i++) {
                const int pp = __shfl_sync(0xffffffffu, p, i), qq = __shfl_sync(0xffffffffu, q, i);
                const double cc = __shfl_sync(0xffffffffu, c, i), ss = __shfl_sync(0xffffffffu, sn, i);
                if (pp < 0) continue;
                for (int r = lane; r < n; r += 32) {
                    const double arp = A[r * n + pp], arq = A[r * n + qq];
                    A[r * n + pp] = cc * arp - ss * arq; A[r * n + qq] = ss * arp + cc * arq;
                    if (V) { const double vrp = V[r * n + pp], vrq = V[r * n + qq]; V[r * n + pp] = cc * vrp - ss * vrq; V[r * n + qq] = ss * vrp + cc * vrq; }
                }
            }
            __syncwarp();
            // rows: A <- J' A, one column per lane
            for (int i = 0; i < npairs; i++) {
                const int pp = __shfl_sync(0xffffffffu, p, i), qq = __shfl_sync(0xffffffffu, q, i);
                const double cc = __shfl_sync(0xffffffffu, c, i), ss = __shfl_sync(0xffffffffu, sn, i);
                if (pp < 0) continue;
                for (int col = lane; col < n; col += 32) {
                    const double apc = A[pp * n + col], aqc = A[qq * n + col];
                    A[pp * n + col] = cc * apc - ss * aqc; A[qq * n + col] = ss * apc + cc * aqc;
                }
            }
            __syncwarp();
            if (lane < npairs && p >= 0) { A[p * n + q] = 0.0; A[q * n + p] = 0.0; }     // annihilated exactly
            __syncwarp();
        }
    }
    __syncwarp();
}

}  // namespace crm
