/*
 * crm_b200.h -- C ABI of libcrm_b200.so, the B200 (sm_100a) implementation of CellRegMap's per-variant scans.
 *
 * The reference (limix/CellRegMap) is pure Python and has no FFI of its own; this ABI is the thin layer the
 * Python mirror (cellregmap_b200/_cellregmap.py) binds with ctypes.  Each entry point names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions: all matrices are float64, row-major, with an explicit leading dimension in elements; every data
 * pointer is a DEVICE pointer unless the parameter name ends in `_host`; sizes are int64_t / int; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream); the caller owns inputs and outputs (outputs must be
 * allocated by the caller), internal workspaces are owned by the handle.  Return value: 0 ok, <0 invalid
 * argument / unsupported shape / bad state (-4: non-finite values in an input matrix), >0 CUDA or cuSOLVER failure; crm_last_error() gives the message of
 * the last failure on the calling thread.  No exceptions cross the boundary.  A handle is not thread-safe;
 * different handles may be used from different threads.
 */
#ifndef CRM_B200_H
#define CRM_B200_H
#include <stdint.h>
#if defined(__GNUC__)
#define CRM_API __attribute__((visibility("default")))
#else
#define CRM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crm_handle_s* crm_handle_t;

CRM_API int crm_version(void);
CRM_API const char* crm_last_error(void);

/* Model object: replaces CellRegMap.__init__ state (cellregmap/_cellregmap.py:63-131). */
CRM_API int crm_create(crm_handle_t* out, int device);
CRM_API int crm_destroy(crm_handle_t h);
/* Device memory is allocated stream-ordered (on the `stream` of the call in flight) from a private per-device pool that keeps freed
 * blocks for the next model object; crm_destroy releases a handle's buffers in the order of the stream of its latest call, without a
 * device-wide synchronisation.  crm_trim_pool returns the cached blocks of `device` to the driver. */
CRM_API int crm_trim_pool(int device);

/*
 * Constructor set-up (cellregmap/_cellregmap.py:93-131 + numpy_sugar.economic_qs_linear, semantics in
 * cellregmap/_math.py:204-256): eigendecomposition of rho*E1 E1' + (1-rho)*L L' for every rho of the grid,
 * done in the column space of the shared half-basis H = [E1 | L].
 *   y (n), W (n x c, ldw), E0 (n x k0, lde0) contexts of the tested GxC term, E1 (n x k1, lde1) contexts of the
 *   background, L (n x mL, ldl) = concatenated Ls / hK (NULL with mL = 0 for the E1-only background),
 *   rho_host[R] the rho1 grid (host memory).
 */
CRM_API int crm_setup(crm_handle_t h, const double* y, const double* W, int64_t ldw, const double* E0, int64_t lde0,
              const double* E1, int64_t lde1, const double* L, int64_t ldl, int64_t n, int c, int k0, int k1,
              int64_t mL, const double* rho_host, int R, void* stream);

/*
 * Set-up shared between the ranks of a multi-GPU scan (one process per GPU).  Every rank builds the operands and the Gram of
 * the half-basis, but decomposes only the grid points r = r_first, r_first + r_step, ... (r_first = rank, r_step = world size);
 * crm_export_basis packs a decomposed grid point into a record of crm_basis_record_size() doubles ([kept rank, solver info, S0 (mp),
 * T_rho (m x mp)], device memory), the caller all-gathers the records (NCCL over NVLink) and hands the other ranks' grid points to
 * crm_import_basis; crm_setup_finish completes the model.  Every rank ends up with the same bits in every grid point, so the sharded scan
 * returns exactly the per-SNP results of the single-GPU call that uses this basis.
 */
CRM_API int crm_setup_partial(crm_handle_t h, const double* y, const double* W, int64_t ldw, const double* E0, int64_t lde0,
                              const double* E1, int64_t lde1, const double* L, int64_t ldl, int64_t n, int c, int k0, int k1,
                              int64_t mL, const double* rho_host, int R, int r_first, int r_step, void* stream);
CRM_API int64_t crm_basis_record_size(crm_handle_t h);
CRM_API int crm_export_basis(crm_handle_t h, int r, double* out, void* stream);
CRM_API int crm_import_basis(crm_handle_t h, int r, const double* in, void* stream);
CRM_API int crm_setup_finish(crm_handle_t h, void* stream);

/* Start the host-to-device transfer of a host-resident genotype matrix (rows x p doubles, leading dimension ldg; pinned memory
 * for an asynchronous copy) ahead of the scan: returns at once, the copy runs in column chunks on the handle's copy stream.  May be
 * called right after crm_create, so that the transfer overlaps crm_setup; the next crm_scan_* call with g_on_host = 1 and the same
 * (pointer, ldg, p) consumes the staged matrix block by block as the chunks arrive (one scan per staging).  A matrix larger than a
 * quarter of the device memory is not staged (the scan streams it as usual).  The host array must stay alive and unchanged until that
 * scan returns.  (Extension: the reference passes G to scan_interaction only, cellregmap/_cellregmap.py:317.) */
CRM_API int crm_stage_genotypes(crm_handle_t h, const double* G_host, int64_t ldg, int64_t rows, int64_t p, void* stream);

/* Element types of genotype matrices.  Device matrices may be CRM_G_F64 or CRM_G_I8; host matrices any of them.  The scan entry points
 * take the type in bits 4..7 of their `g_on_host` argument (0 = float64, the reference's asarray(G, float)); the data pointer is then
 * reinterpreted and the leading dimension counts elements of that type. */
enum { CRM_G_F64 = 0, CRM_G_I8 = 1, CRM_G_U8 = 2, CRM_G_I16 = 3, CRM_G_I32 = 4, CRM_G_F32 = 5, CRM_G_I64 = 6 };
#define CRM_G_FLAGS(on_host, donor_level, dtype) (((on_host) ? 1 : 0) | ((donor_level) ? 2 : 0) | ((dtype) << 4))

/* crm_stage_genotypes for any element type and any host memory.  Pinned float64 is moved by DMA as above.  Everything else -- in
 * particular a pageable float64 numpy array, which is what a user of the reference passes -- is converted to int8 dosage blocks by a
 * pool of host threads (CRM_HOST_THREADS, default min(CPUs, 16)) into pinned buffers, block by block ahead of the device: 1 byte per
 * dosage crosses PCIe, and the conversion overlaps the set-up and the scan of earlier blocks.  A block that is not integer-valued in
 * [-127, 127] ends the conversion; the rest of the matrix is moved as float64.  basis_cols_hint: (1 + k0) * (m + 1 + c) rounded up to
 * even, or 0 (sizes the blocks to whole waves of the int8 contraction).  The host array must stay alive and unchanged until the scan
 * that consumes it returns (or the handle is destroyed). */
CRM_API int crm_stage_genotypes_typed(crm_handle_t h, const void* G_host, int dtype, int64_t ldg, int64_t rows, int64_t p,
                                      int64_t basis_cols_hint, void* stream);
/* FP64 tensor-core (DMMA) peak of the current device in TFLOP/s, measured with a register-resident mma.sync.m8n8k4.f64 probe
 * (a few milliseconds; synchronises the stream): the roofline denominator of the float64 rotation (bench.py). */
CRM_API int crm_fp64_tensor_peak(double* tflops, void* stream);
/* Worker threads of the host feeder. */
CRM_API int crm_host_threads(void);
/* Column blocks the feeder would cut a host matrix of p SNP columns into when the rotation contracts basis_cols columns (0: unknown):
 * starts[0..nblocks] (at most `capacity` entries are written; starts may be NULL).  Host-side logic only (tests): widths are whole
 * 256-SNP tiles chosen so that the persistent grid of the int8 contraction ends on full waves. */
CRM_API int crm_feeder_blocks(int64_t p, int64_t basis_cols, int64_t* starts, int32_t capacity, int32_t* nblocks);
/* The feeder's conversion on its own (host memory in, host memory out, no device involved): rows x cols elements of type `dtype`
 * (leading dimension ld) -> int8 (leading dimension ldd bytes); *bad = 1 when some entry is not an integer in [-127, 127] (such
 * entries are written as saturated / arbitrary values), *gmax = largest |entry| among the valid ones. */
CRM_API int crm_host_narrow(const void* src_host, int dtype, int64_t ld, int64_t rows, int64_t cols, int8_t* dst_host, int64_t ldd,
                            int32_t* bad, int32_t* gmax);

/* Replaces the tested-context matrix E0 (row-permuted contexts: idx_E of scan_interaction, _cellregmap.py:398-401). */
CRM_API int crm_set_test_contexts(crm_handle_t h, const double* E0, int64_t lde0, void* stream);

/*
 * Declares how the background half-covariance was built (reference get_L_values, _cellregmap.py:533-545, as called by
 * run_interaction :577-580 / estimate_betas :672-675 with E2 = E):  L[:, i q + c] = (E0 M)[:, i] * hK[:, c]  with hK the n x q
 * kinship factor (device, leading dimension ldhk) and M the k0 x r map from the contexts to the scaled left singular vectors
 * U S = E0 M of the context matrix (host, row-major).  The products L.E0_j that the rotation needs are then combinations of the
 * triple products hK_c.E0_l.E0_j, symmetric in (l, j): the rotation contracts only the k0 (k0 + 1) / 2 q distinct ones
 * (plus Hx itself and the products of E1, y and W) and rebuilds the rest with the k0 x r map -- 0.56 of the contraction work and of
 * the digit planes at k0 = 20.  The library verifies the claim against the basis it was set up with (one pass over L, one
 * synchronisation of `stream`); *accepted = 1 when the structure is used, 0 when it does not hold or does not apply (then the call
 * changes nothing).  Two ways to call it.  After crm_setup / crm_setup_finish: verified at once.  Before crm_setup (hK and M must
 * stay valid until the set-up returns): applied and verified inside the set-up without an extra synchronisation, *accepted stays 0 and
 * crm_rotation_rows tells afterwards whether the compact form is in use.  crm_set_test_contexts drops the declaration (tested
 * contexts that differ from the contexts inside L have no such symmetry).  CRM_KR=0 in the environment ignores it.
 */
CRM_API int crm_set_background_factors(crm_handle_t h, const double* hK, int64_t ldhk, int q, const double* M, int r, int* accepted,
                                       void* stream);

/* Hint before crm_setup: the scans of this model will most likely receive integer (or affine-integer) dosages at cell level.  The
 * set-up then builds the int8 digit planes of the basis on a side stream next to the latency-bound phases of its eigensolver instead
 * of leaving them to the first rotation (7 ms of a 147 ms step at BASELINE configs[2]); costs nothing but that kernel time if the
 * genotypes turn out to be real-valued.  CRM_EARLY_PLANES=0 ignores the hint. */
CRM_API int crm_hint_integer_genotypes(crm_handle_t h, int likely);

/* Columns of the basis operand that the rotation contracts per SNP: the expanded basis [Hx | Hx.E0_j] (full_rows = (1 + k0) * ld)
 * and what is contracted once a structure of the background has been accepted (used_rows; equal to full_rows otherwise).
 * bench.py scales the algorithmic flop of the rotation by used_rows / full_rows to report the work actually executed. */
CRM_API int crm_rotation_rows(crm_handle_t h, int64_t* full_rows, int64_t* used_rows);

/*
 * New phenotype y (n doubles, device) for the same cells, contexts, covariates and background (extension): refreshes only
 * the y-dependent state of the model; the Gram of the half-basis and the per-rho eigendecompositions are kept.  Equivalent
 * to crm_setup with the new y.  Typical use: one model per data set, one crm_update_phenotype + scan per gene.
 */
CRM_API int crm_update_phenotype(crm_handle_t h, const double* y, void* stream);

/*
 * Donor-level genotype ingress (extension).  The reference always receives genotypes expanded from donors to
 * cells, G_cells = G_donors[donor_of_cell].  After crm_set_donors the scan entry points also accept the d x p donor-level
 * matrix (flag bit 1 of `g_on_host`, see below): every contraction over cells is then done once per gene on the basis side
 * (sums over the cells of each donor) and the per-SNP work contracts over d donors instead of n cells.  Results equal the
 * expanded call up to summation order.  perm [n]: cell indices grouped by donor; offsets [d+1]: start of each donor's group
 * (both int32, device).  Must be called after crm_setup (and is refreshed by crm_set_test_contexts).
 */
CRM_API int crm_set_donors(crm_handle_t h, const int32_t* perm, const int32_t* offsets, int64_t d, void* stream);

/* Sizes fixed by crm_setup: [0]=n [1]=c [2]=k0 [3]=m (columns of H) [4]=R [5]=padded m [6]=max kept rank
 * [7]=1 when the pre-expanded basis [Hx | Hx.E0_j] of the float64 route is resident (rotation runs as a plain contraction), 0 when it
 * is not, -1 before the first float64 rotation of the model has decided. */
CRM_API int crm_get_dims(crm_handle_t h, int64_t* dims8);
/* Copies S0 of grid point r (padded to dims[5], zeros beyond the kept rank) into out (device). */
CRM_API int crm_get_spectrum(crm_handle_t h, int r, double* out, void* stream);

/*
 * Interaction scan: replaces CellRegMap.scan_interaction (cellregmap/_cellregmap.py:317-440) for p SNPs.
 *   G: n x p genotypes, leading dimension ldg; device memory, or pinned/pageable host memory when bit 0 of g_on_host is
 *      set (then the columns are streamed to the device in chunks, overlapped with compute).  Bit 1 of g_on_host set:
 *      G is the d x p donor-level matrix (see crm_set_donors).  The same two bits apply to crm_scan_association and
 *      crm_predict_interaction.
 *   Gtest: optional n x p genotypes used only in the tested design g.E0 (idx_G permutation, :410-413), else NULL.
 *   out_pv, out_rho1, out_e2, out_g2, out_eps2: p doubles each (device).
 *   Optional diagnostics (device, may be NULL): d_lml/d_delta/d_scale [p][R], d_Q [p], d_lam [p][k0],
 *   d_nlam [p] (int32), d_M [p][k0*k0], d_liu [p], d_ifault [p] (int32), d_flags [p] (int32: bit0 no eigenvalue
 *   > 0, bit1 rank-deficient design, bit2 det(H) <= 0 in a fit), d_nfev [p][R] (int32).
 *   Optional overrides (device, may be NULL): fix the selected grid index and variance components per SNP
 *   (stage-injected parity tests): ov_rho_idx [p] (int32), ov_v0 [p], ov_v1 [p].
 */
typedef struct {
    double* lml; double* delta; double* scale; double* Q; double* lam; int32_t* nlam; double* M; double* liu;
    int32_t* ifault; int32_t* flags; int32_t* nfev;
    const int32_t* ov_rho_idx; const double* ov_v0; const double* ov_v1;
} crm_scan_diag_t;

CRM_API int crm_scan_interaction(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, const double* Gtest,
                         int64_t ldgt, double* out_pv, double* out_rho1, double* out_e2, double* out_g2,
                         double* out_eps2, const crm_scan_diag_t* diag, void* stream);

/*
 * Association scans: replace CellRegMap.scan_association / scan_association_fast
 * (cellregmap/_cellregmap.py:246-281, 284-314) and lrt_pvalues (:443-469).  info4 (device) receives
 * rho1, e2, g2, eps2 of the null model; out_null_lml (device, 1 double, may be NULL).
 */
CRM_API int crm_scan_association(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, int fast,
                         double* out_pv, double* out_alt_lml, double* info4, double* out_null_lml, void* stream);

/*
 * Effect sizes: replaces CellRegMap.predict_interaction (cellregmap/_cellregmap.py:137-205) for p SNPs: per SNP and rho1 a
 * REML fit of y ~ [W g E0] under rho1 (g.E0)(g.E0)' + (1-rho1) sum_i L_i L_i' (+ noise), best rho1 by strict '>', then
 * beta_g = beta[c] and beta_gxe = v0 rho1 E0 (g.E0)' K^-1 (y - M beta) / sqrt(2 maf (1 - maf)).
 *   maf: p doubles (device).  use_background: 1 when the L passed to crm_setup is the Ls list (the only background this
 *   entry point uses, as in the reference), 0 to ignore it (model built from hK or without background).
 *   out_beta_g: p doubles; out_beta_gxe: n x p (leading dimension ldo >= p), element [i][s] = effect of SNP s in cell i,
 *   i.e. the reference's (1, n, p) array; out_rho1: selected rho1 per SNP (may be NULL).
 */
CRM_API int crm_predict_interaction(crm_handle_t h, const double* G, int64_t ldg, int64_t p, int g_on_host, const double* maf,
                                    int use_background, double* out_beta_g, double* out_beta_gxe, int64_t ldo, double* out_rho1,
                                    void* stream);

/* Number of CUDA kernels this library has launched in the process so far (bench.py's gpu_launches). */
CRM_API long long crm_launch_count(void);
/* Event timing of the rotation kernel (K1) on the launching stream: returns the totals accumulated since the last
 * call (ms, algorithmic flop = 2 n m (1+k0) per SNP, launches), resets them, and switches the timing on/off. */
CRM_API int crm_profile(crm_handle_t h, int enable, double* rot_ms, double* rot_flops, int64_t* rot_launches);

/* Same for the int8 tensor-core contraction of the exact int8 split of the rotation (taken for integer dosages unless
 * CRM_ROTATION=dmma): totals since the last call (ms, int8 multiply-add operations x 2, launches). */
CRM_API int crm_profile_int8(crm_handle_t h, double* gemm_ms, double* gemm_ops, int64_t* launches);

/* ---- stage-level entry points (used by the parity tests and by the Python mirror) ---- */

/* K1: out[n_count][m_count] (ldc) = B[:, n_begin:+n_count]' A[:, m_begin:+m_count] over K rows.
 * mode 0: B as given; 1: B .* B2; 2: B[k][s*kexp+j] = G[k][s] * Eext[k][j] with G = B, Eext = B2 (ldb2 = its width). */
CRM_API int crm_gemm(int mode, const double* A, int64_t lda, int64_t a_cols, const double* B, int64_t ldb, int64_t b_cols,
             const double* B2, int64_t ldb2, int64_t b2_cols, int64_t K, int m_begin, int m_count, int64_t n_begin,
             int64_t n_count, double* out, int64_t ldc, int kexp, void* stream);

/* Set-up eigensolver on caller-supplied matrices: `batch` symmetric n x n matrices A [batch][n][n] (device, both triangles) ->
 * eigenvalues W [batch][n] ascending and orthonormal eigenvectors V [batch][n][n] (V[b][t*n + i] = component i of vector t), the
 * batched replacement of numpy_sugar.economic_qs_linear's eigh per rho (cellregmap/_cellregmap.py:129).  quality_host [batch]
 * (optional) = largest residual |T z - lambda z|_inf / |T| of the tridiagonal eigenpairs (inf when a library step reported an
 * error); ms (optional) = CUDA-event time.  Synchronises the stream. */
CRM_API int crm_eigh_batched(const double* A, int n, int batch, double* W, double* V, double* quality_host, float* ms, void* stream);

/* K0 on caller-supplied operands: C[B][cols] (ldc) = G' X by the exact int8 split -- 8 digit planes of the real matrix X [n][cols]
 * (ldx) against the integer-valued matrix G [n][B] (ldg); replaces the same `Q0.T @` products as crm_gemm (cellregmap/_math.py:72-73)
 * when the genotypes are integer dosages.  route 0: hand-written tcgen05 kernel with fused fp64 recombination (the product path); 1: cuBLASLt int8
 * GEMM + recombination kernel; 2: the CTA-pair (cta_group::2) variant of the tcgen05 kernel -- all bit-identical.  flags2 (host) = {G not integer in [-127,127], max |g|}; contraction_ms (host, optional) =
 * CUDA-event time of the contraction alone.  Synchronises the stream. */
CRM_API int crm_int8_split_gemm(const double* X, int64_t ldx, int64_t cols, const double* G, int64_t ldg, int64_t B, int64_t n, int route,
                        double* C, int64_t ldc, int32_t* flags2, float* contraction_ms, void* stream);

/* K2 on caller-supplied rotated statistics: replaces glimix_core LMM(y, X, QS, restricted).fit() for p x R problems.
 * S, yr [R][mp]; Wr [R][c][mp]; gr [p][R*mp] (NULL: design is W only, p must be 1); gy [p], gW [p][c], gg [p];
 * stats = [y'y, W'y (c), W'W (c*c)].  Outputs [p][R] (beta [p][R][c + has_g]). */
CRM_API int crm_lmm_fit_rotated(const double* S, const double* yr, const double* Wr, const double* gr, const double* gy,
                        const double* gW, const double* gg, const double* stats, int m, int mp, int R, int c, int64_t p,
                        double n, int restricted, double* lml, double* delta, double* scale, double* beta,
                        int32_t* nfev, int32_t* flags, void* stream);

/* K4: Davies p-values with modified-Liu fall-backs: replaces chiscore.davies_pvalue's _pvalue_lambda for `count`
 * (Q, eigenvalue list) pairs.  lam [count][lam_ld], nlam [count] valid entries each.  Optional outputs may be NULL;
 * trace8 [count][8] = qfval, trace[0..6] of the published routine. */
CRM_API int crm_davies_pvalues(const double* Q, const double* lam, const int32_t* nlam, int lam_ld, int64_t count, int lim,
                       double acc, double* pv, double* liu, int32_t* ifault, int32_t* converged, double* trace8,
                       void* stream);

/* Batched score_statistic_liu_params (cellregmap/_math.py:163-180): modified-Liu parameters of `count` (Q, eigenvalue list)
 * pairs; out4 [count][4] = pv, mu_q, sigma_q, dof_x. */
CRM_API int crm_liu_params(const double* Q, const double* lam, const int32_t* nlam, int lam_ld, int64_t count, double* out4,
                           void* stream);
/* Batched qmin (cellregmap/_math.py:183-201): params4 [count][nrho][4] = pv, mu_q, sigma_q, dof_x -> out [count][nrho]. */
CRM_API int crm_qmin(const double* params4, int nrho, int64_t count, double* out, void* stream);

/* lrt_pvalues (cellregmap/_cellregmap.py:443-469) with dof = 1. */
CRM_API int crm_lrt_pvalues(const double* alt_lml, double null_lml, int64_t count, double* pv, void* stream);
/* The same for any dof > 0 (chi2(dof).sf; dof = 1 takes the erfc form above). */
CRM_API int crm_lrt_pvalues_dof(const double* alt_lml, double null_lml, int64_t count, double dof, double* pv, void* stream);

#ifdef __cplusplus
}
#endif
#endif
