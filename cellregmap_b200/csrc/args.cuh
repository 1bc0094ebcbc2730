// Argument blocks of the kernels (shared between the kernel translation units and the host orchestration).
#pragma once
#include "common.cuh"

namespace crm {

// index of pair (j >= l) in the packed lower-triangular layout used for the E0_j * E0_l product columns
__host__ __device__ __forceinline__ int pair_index(int j, int l) { return j * (j + 1) / 2 + l; }

struct FitArgs {
    // shared per rho1 (padded leading dimension mp; entries beyond the kept rank are zero)
    const double* S;    // [R][mp]
    const double* yr;   // [R][mp]
    const double* Wr;   // [R][c][mp]
    // per SNP
    const double* gr;   // [p][R*mp]   rotated genotype, row stride gr_ld
    long long gr_ld;
    const double* gy;   // g'y per SNP, element stride gy_ld
    long long gy_ld;
    const double* gW;   // g'W_a per SNP at gW[s * gW_ld + a]
    long long gW_ld;
    const double* gg;   // g'g per SNP, element stride gg_ld
    long long gg_ld;
    // plain statistics of the null design: [y'y, W'y (c), W'W (c*c)]
    const double* stats;
    int m, mp, R, c, p;
    double n;           // number of samples
    int restricted;
    const double* fixed_x;   // when non-null: no search, evaluate at logit(delta) = *fixed_x (FastScanner semantics)
    // optional table of the SNP-independent part of the objective at the points every bracket search visits (fit.cuh: FIT_TAB_*), [R][2][tab_k] records
    const double* tab; int tab_k;
    // outputs, [p][R] (beta: [p][R][P]); xopt may be null
    double* lml; double* delta; double* scale; double* beta; double* xopt; int* nfev; int* flags;
};

struct ScoreArgs {
    // per-rho shared rotated quantities
    const double* S; const double* yr; const double* Wr;   // [R][mp], [R][mp], [R][c][mp]
    int m, mp, R, c, k, kexp, p;
    // per SNP
    const int* perm;        // sorted position -> SNP
    const int* rho_idx;     // [p]
    const double* v0; const double* v1;   // [p]
    const double* gr; long long gr_ld;    // rotated genotype [p][R*mp]
    const double* GEr;      // rotated g.E0 in sorted order: [(pos*k + j)][mp]
    const double* rot; long long rot_ld;  // K1 output rows (s*kexp + j), columns col_y / col_W + a
    int col_y, col_W;
    const double* sq; long long sq_ld;    // squared-genotype Grams per SNP: [gg | GE'g (k) | GE'GE pairs (k(k+1)/2)]
    const double* stats;    // [y'y, W'y (c), W'W (c*c)]
    // outputs (indexed by SNP)
    double* Q; double* lam; int lam_ld; int* nlam; int* flags;
    double* Mout;           // optional [p][k*k] weight matrices (diagnostics / stage parity), may be null
};

constexpr int SCORE_MAX_NZ = 63;

struct PvalArgs {
    const double* Q;        // [count]
    const double* lam;      // [count][lam_ld] eigenvalues, any order, nlam[i] valid entries
    const int* nlam;        // [count]
    int lam_ld;
    int count;
    int lim; double acc;
    double* pv;             // [count]
    double* liu;            // [count] modified-Liu p-value (may be null)
    int* ifault;            // [count] (may be null)
    int* converged;         // [count] (may be null)
    double* trace;          // [count][8]: qfval, trace[0..6] (may be null)
};

constexpr int PV_MAXLAM = 128;

// K5 (betas.cuh): general-design LMM fit in shared memory, one CTA per (SNP, rho index).  Two uses:
//  * effect-size model of predict_interaction: design [W g E0], covariance (1-d)(rho U U' + (1-rho) B) + d I with
//    U = g.E0 (k0 columns, Woodbury update), S / Zs shared by all rho (mix_rho = 1);
//  * plain two-component fits with wide designs (more than 8 columns; K2 covers the narrow ones): k0 = 0, covariance
//    (1-d) K_rho + d I with per-rho spectra and rotated columns (mix_rho = 0, strides select the rho block).
struct BetaArgs {
    const double* S;            // [mp] spectrum (zeros beyond the kept rank); + rho_index * S_stride
    const double* Zs;           // [1 + c + k0][mp]  rotated shared columns  Q'[y | W | E0]; + rho_index * Zs_stride
    const double* Zp;           // [p][has_g + k0][mp] rotated per-SNP columns Q'[g | g.E0]; SNP stride Zp_snp_stride, + rho_index * Zp_rho_stride
    long long S_stride, Zs_stride, Zp_snp_stride, Zp_rho_stride;
    const double* shared_gram;  // [(1 + c + k0)^2]  plain Gram of [y | W | E0]
    const double* rot; long long rot_ld; int col_y, col_W, kexp;   // K1 output rows (s*kexp + j): y'(.) and W'(.) columns
    const double* lin; long long lin_ld;   // per SNP [sum g | g'E0 (k0) | g'(E0_j E0_l) pairs]   (k0 > 0 only)
    const double* sq; long long sq_ld;     // per SNP [g'g | (g^2)'E0 (k0) | (g^2)'(E0_j E0_l) pairs]
    const double* rho;          // [R] grid (device); used when mix_rho
    int m, mp, c, k0, R, p, has_g, mix_rho, restricted;
    const double* fixed_x;      // non-null: evaluate at logit(delta) = *fixed_x instead of searching
    double n;
    // outputs [p][R]: lml, delta, scale; beta [p][R][c + has_g + k0]; ucoef [p][R][k0] = (g.E0)' K^-1 (y - M beta) (k0 > 0)
    double* lml; double* delta; double* scale; double* beta; double* ucoef; double* xopt; int* nfev; int* flags;
};
constexpr int BETA_MAX_NZ = 67;

// exact int8 split (ozaki.cuh, oz_mma.cuh)
constexpr int OZ_SLICES = 8;               // signed 7-bit digit planes per column
constexpr int OZ_EXP_EMPTY = -100000;      // exponent marker of an all-zero column

}  // namespace crm
