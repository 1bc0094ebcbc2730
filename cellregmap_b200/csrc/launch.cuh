// Host-callable launchers of the kernels in this directory (each defined in its own translation unit).
#pragma once
#include <functional>
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

// stream-ordered scratch from the library's private memory pool (abi.cu); release with cudaFreeAsync
cudaError_t pool_alloc_async(void** ptr, size_t bytes, cudaStream_t st);

// FP64 tensor-core (DMMA) issue-rate peak of the current device in TFLOP/s, measured with a register-resident probe kernel (~5 ms)
int measure_fp64_tensor_peak(double* tflops, cudaStream_t stream);

// next launch_gemm calls on this thread may choose the K split by grid size (operands without a SNP dimension)
void gemm_set_free_split(bool on);
// with free split: next launch_gemm calls on this thread whose two operands are the same matrix compute only the lower tile triangle
void gemm_set_symmetric(bool on);

int launch_fit_with_g(const FitArgs& fa, cudaStream_t st);   // design [W g], P = c + 1 in 1..8
int launch_fit_null(const FitArgs& fa, cudaStream_t st);     // design W,     P = c     in 1..7
inline int launch_fit(const FitArgs& fa, bool has_g, cudaStream_t st) { return has_g ? launch_fit_with_g(fa, st) : launch_fit_null(fa, st); }
// table of the bracket points of the search (fit.cuh): bytes, points per direction, builder (fa.tab_k = fit_table_points())
size_t fit_table_bytes(int R, int mp);
int fit_table_points();
int launch_fit_table(const FitArgs& fa, double* tab, cudaStream_t st);

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
int launch_lrt_dof(const double* alt_lml, double null_lml, long long count, double dof, double* pv, cudaStream_t st);

// exact int8 split of the rotation (ozaki.cuh / kernels_ozaki.cu)
constexpr int OZAKI_SLICES = 8;
int oz_launch_exponents(const double* Hx, int ldH, const double* Eext, int epitch, int kexp, long long n, int* expo, cudaStream_t st);
int oz_launch_slices(const double* Hx, int ldH, const double* Eext, int epitch, int j0, int nj, long long n, const int* expo, int8_t* A8, long long Mp,
                     long long Kp, cudaStream_t st);
// general form: product columns F[:, j0 + jj] * X[:, a] (a < cols, jj < nj) at rows row0 + jj * rstride + a of the plane set
int oz_launch_fill_exponents(int* expo, long long rows, cudaStream_t st);
int oz_launch_fill_exponents_strided(int* expo, long long row0, long long rstride, int count, cudaStream_t st);
int oz_launch_product_exponents(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, int nj, long long n, int* expo, long long row0,
                                long long rstride, cudaStream_t st);
int oz_launch_product_slices(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, int nj, long long n, const int* expo, int8_t* A8,
                             long long Mp, long long Kp, long long row0, long long rstride, cudaStream_t st);
int oz_launch_genotypes(const double* G, long long ldg, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp, long long Kp, int* flags, cudaStream_t st);
// int8 dosages stored row-major (cells x SNPs): K-major operands of the contraction (flags may be null), float64 image, finiteness
int oz_launch_transpose_i8(const int8_t* G8, long long ld8, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp, long long Kp, int* flags, cudaStream_t st);
int oz_launch_widen_i8(const int8_t* G8, long long ld8, long long n, long long B, double* out, long long ldo, cudaStream_t st);
int oz_launch_finite_check(const double* G, long long ldg, long long n, long long B, int* flags, cudaStream_t st);
// affine-integer genotype columns g = a d + b (ozaki.cuh): detection + int8 image of d, and the maps back from contractions of d
constexpr int OZ_AFFINE_CHUNKS = 128;
size_t oz_affine_scratch_bytes(long long B);
int oz_launch_affine_genotypes(const double* G, long long ldg, long long n, long long B, void* scratch, double* aff, long long lda, int8_t* Gt8, int8_t* G2t8,
                               long long Bp, long long Kp, int* flags, cudaStream_t st);
int oz_launch_affine_fix(double* C, long long ldc, long long B, long long cols, const double* aff, long long lda, const double* colsum, cudaStream_t st);
int oz_launch_affine_fix_square(double* sq, const double* lin, long long ld, long long B, int cols, const double* aff, long long lda, const double* colsum2, cudaStream_t st);
int oz_launch_matrix_planes(const double* X, long long ldx, int cols, long long n, int* expo, int8_t* P8, long long Mp, long long Kp, cudaStream_t st);
int oz_launch_combine(const int* D, long long Mp, long long ldd, const int* expo, long long Mtot, long long B, double* C, long long ldc, cudaStream_t st);
// the same contraction + recombination in one hand-written tcgen05 kernel (oz_mma.cuh): C[s][col], bit-identical
int oz_launch_mma(const int8_t* A8, long long Mp, long long Mtot, const int* expo, const int8_t* Gt8, long long Bp, long long B, long long Kp, double* C,
                  long long ldc, cudaStream_t st, int variant = -1);   // variant: -1 process default, 1 single CTA, 2 CTA pairs
int oz_int8_gemm(const int8_t* A8, long long Mrows, const int8_t* Gt8, long long Bp, long long Kp, int* D, long long ldd, cudaStream_t st);


// batched symmetric eigensolver of the set-up (eig.cuh / kernels_eig.cu); solver / blas are cusolverDnHandle_t / cublasHandle_t
int eig_workspace_bytes(int n, int batch, size_t* bytes);
int eig_lib_lwork(void* solver, int n, int* lwork);
int eig_batched(void* solver, void* blas, double* A, const int* n_of, int n, int batch, double* W, double* V, double* quality, void* ws, double* lib_work, int lib_lwork,
                int* info_dev, cudaStream_t st, int group_batch = 0, const int* ids = nullptr, const std::function<int()>* after_sytrd = nullptr);
// after_sytrd: called on the host right after the tridiagonalisation has been enqueued (work that may share the device with the
// latency-bound phases that follow)
// group_batch: size of the batch this one is a share of (0: itself); ids: position of each matrix in that batch (seeds of the start vectors).
// With both given, a matrix is decomposed to the same bits whichever share of the batch it is solved in.

}  // namespace crm
