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
// cyclic Jacobi with eigenvectors on a symmetric n x n shared matrix (eigenvalues end on the diagonal, V columns)
__device__ inline void warp_jacobi_vec(double* A, double* V, int n, int lane) {
    for (int e = lane; e < n * n; e += 32) V[e] = (e / n == e % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0.0, dg = 0.0;
        for (int e = lane; e < n * n; e += 32) { const double v = A[e]; if (e / n == e % n) dg += v * v; else off += v * v; }
        off = warp_sum(off); dg = warp_sum(dg);
        if (off <= 1e-30 * dg || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double app = A[p * n + p], aqq = A[q * n + q];
                if (fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) { __syncwarp(); if (lane == 0) { A[p * n + q] = 0.0; A[q * n + p] = 0.0; } __syncwarp(); continue; }
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                __syncwarp();
                for (int r = lane; r < n; r += 32) {
                    if (r != p && r != q) {
                        const double arp = A[r * n + p], arq = A[r * n + q];
                        const double nrp = c * arp - s * arq, nrq = s * arp + c * arq;
                        A[r * n + p] = nrp; A[p * n + r] = nrp; A[r * n + q] = nrq; A[q * n + r] = nrq;
                    }
                    const double vrp = V[r * n + p], vrq = V[r * n + q];
                    V[r * n + p] = c * vrp - s * vrq; V[r * n + q] = s * vrp + c * vrq;
                }
                if (lane == 0) { A[p * n + p] = app - tt * apq; A[q * n + q] = aqq + tt * apq; A[p * n + q] = 0.0; A[q * n + p] = 0.0; }
                __syncwarp();
            }
    }
    __syncwarp();
}

}  // namespace crm
