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
        if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < P - 1; p++) {
#pragma unroll
            for (int q = p + 1; q < P; q++) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
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

}  // namespace crm
