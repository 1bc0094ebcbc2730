// Host-callable launchers of the kernels in this directory (each defined in its own translation unit).
#pragma once
#include "common.cuh"
#include "args.cuh"

namespace crm {

struct GemmOperands {
    // A: K x Mtot row-major
    const double* A; long long lda; long long a_cols;
    // B (PLAIN/PRODUCT) or G (EXPAND): K x Ntot row-major
    const double* B; long long ldb; long long b_cols;
    // PRODUCT: second factor, same shape as B.  EXPAND: Eext, K x epitch row-major (ld == epitch)
    const double* B2; long long ldb2; long long b2_cols;
};
enum GemmMode : int { GEMM_PLAIN = 0, GEMM_PRODUCT = 1, GEMM_EXPAND = 2 };
constexpr int GEMM_TILE_N = 128;

// C[n_count][m_count] (ldc) = B[:, n_begin:+n_count]^T A[:, m_begin:+m_count], B built according to `mode`.
int launch_gemm(int mode, const GemmOperands& op, int K, int m_begin, int m_count, int n_begin, int n_count,
                double* out, long long ldc, int kexp, cudaStream_t stream);

// next launch_gemm calls on this thread may choose the K split by grid size (operands without a SNP dimension)
void gemm_set_free_split(bool on);

int launch_fit_with_g(const FitArgs& fa, cudaStream_t st);   // design [W g], P = c + 1 in 1..8
int launch_fit_null(const FitArgs& fa, cudaStream_t st);     // design W,     P = c     in 1..7
inline int launch_fit(const FitArgs& fa, bool has_g, cudaStream_t st) { return has_g ? launch_fit_with_g(fa, st) : launch_fit_null(fa, st); }

int launch_score(const ScoreArgs& sa, long long count, cudaStream_t st);
int launch_select(const double* lml, const double* delta, const double* scale, int p, int R, int* rho_idx, double* best_lml,
                  double* v0, double* v1, cudaStream_t st);
int launch_group(const int* rho_idx, int p, int R, int* perm, int* offsets, cudaStream_t st);
int launch_gather_transpose(const double* C, long long ldc, const int* perm, int kexp, int joff, int kcols, long long nq,
                            int na, double* out, long long ldo, cudaStream_t st);
int launch_pvalues(const PvalArgs& pa, cudaStream_t st);
int launch_beta_fit(const BetaArgs& ba, cudaStream_t st);
int launch_beta_gxe(const double* E0, long long lde0, const double* coef, int k0, long long n, long long p, double* out,
                    long long ldo, long long s0, cudaStream_t st);
int launch_liu_params(const double* Q, const double* lam, const int* nlam, int lam_ld, long long count, double* out, cudaStream_t st);
int launch_qmin(const double* params, int nrho, long long count, double* out, cudaStream_t st);
int launch_lrt(const double* alt_lml, double null_lml, long long count, double* pv, cudaStream_t st);

}  // namespace crm
