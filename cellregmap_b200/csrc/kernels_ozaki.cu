// Translation unit: exact int8 split of the rotation (ozaki.cuh) -- slicing, genotype conversion, the int8 tensor-core
// contraction and the fp64 recombination.
#include <cublasLt.h>

#include <algorithm>

#include "launch.cuh"
#include "ozaki.cuh"

namespace crm {

static __global__ void oz_fill_kernel(int* p, long long count, int value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = value;
}

static __global__ void oz_fill_strided_kernel(int* p, long long stride, int count, int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[(long long)i * stride] = value;
}
int oz_launch_fill_exponents_strided(int* expo, long long row0, long long rstride, int count, cudaStream_t st) {
    if (count <= 0) return CRM_OK;
    oz_fill_strided_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(expo + row0, rstride, count, OZ_EXP_EMPTY);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int oz_launch_fill_exponents(int* expo, long long rows, cudaStream_t st) {
    oz_fill_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(expo, rows, OZ_EXP_EMPTY);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
// exponents and digit planes of the product columns F[:, j0 + jj] * X[:, a] at rows row0 + jj * rstride + a  (ozaki.cuh)
int oz_launch_product_exponents(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, int nj, long long n, int* expo, long long row0,
                                long long rstride, cudaStream_t st) {
    if (cols <= 0 || nj <= 0) return CRM_OK;
    for (int y0 = 0; y0 < nj; y0 += 32768) {
        const int ny = std::min(32768, nj - y0);
        dim3 grid((unsigned)((cols + 127) / 128), (unsigned)ny, (unsigned)std::max<long long>(1, std::min<long long>(64, n / 512)));
        oz_column_exponent_kernel<<<grid, 128, 0, st>>>(X, ldx, cols, F, ldf, j0 + y0, n, expo, row0 + (long long)y0 * rstride, rstride);
        CRM_CUDA(cudaGetLastError()); count_launch();
    }
    return CRM_OK;
}
int oz_launch_product_slices(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, int nj, long long n, const int* expo, int8_t* A8,
                             long long Mp, long long Kp, long long row0, long long rstride, cudaStream_t st) {
    if (cols <= 0 || nj <= 0) return CRM_OK;
    dim3 grid((unsigned)((Kp + OZ_ROWS - 1) / OZ_ROWS), (unsigned)((cols + OZ_TILE - 1) / OZ_TILE), 1);
    oz_slice_kernel<<<grid, 256, 0, st>>>(X, ldx, cols, F, ldf, j0, nj, n, expo, A8, Mp, Kp, row0, rstride);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int oz_launch_exponents(const double* Hx, int ldH, const double* Eext, int epitch, int kexp, long long n, int* expo, cudaStream_t st) {
    CRM_CHECK(oz_launch_fill_exponents(expo, (long long)kexp * ldH, st));
    return oz_launch_product_exponents(Hx, ldH, ldH, Eext, epitch, 0, kexp, n, expo, 0, ldH, st);
}

int oz_launch_slices(const double* Hx, int ldH, const double* Eext, int epitch, int j0, int nj, long long n, const int* expo, int8_t* A8, long long Mp,
                     long long Kp, cudaStream_t st) {
    return oz_launch_product_slices(Hx, ldH, ldH, Eext, epitch, j0, nj, n, expo, A8, Mp, Kp, (long long)j0 * ldH, ldH, st);
}

int oz_launch_matrix_planes(const double* X, long long ldx, int cols, long long n, int* expo, int8_t* P8, long long Mp, long long Kp, cudaStream_t st) {
    oz_fill_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(expo, cols, OZ_EXP_EMPTY);
    CRM_CUDA(cudaGetLastError()); count_launch();
    dim3 g1((unsigned)((cols + 127) / 128), (unsigned)std::max<long long>(1, std::min<long long>(256, n / 256)), 1);
    oz_matrix_exponent_kernel<<<g1, 128, 0, st>>>(X, ldx, cols, n, expo);
    CRM_CUDA(cudaGetLastError()); count_launch();
    dim3 g2((unsigned)((Kp + OZ_ROWS - 1) / OZ_ROWS), (unsigned)((cols + OZ_TILE - 1) / OZ_TILE), 1);
    oz_matrix_slice_kernel<<<g2, 256, 0, st>>>(X, ldx, cols, n, expo, P8, Mp, Kp);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int oz_launch_genotypes(const double* G, long long ldg, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp, long long Kp, int* flags, cudaStream_t st) {
    CRM_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    const long long blocks = ((Kp + OZ_ROWS - 1) / OZ_ROWS) * ((Bp + OZ_TILE - 1) / OZ_TILE);
    if (blocks > 2147483647LL) { set_error("genotype block too large for one conversion launch"); return CRM_ERR_UNSUPPORTED; }
    oz_genotype_kernel<<<(unsigned)blocks, 256, 0, st>>>(G, ldg, n, B, Gt8, G2t8, Bp, Kp, flags);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int oz_launch_transpose_i8(const int8_t* G8, long long ld8, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp, long long Kp, int* flags, cudaStream_t st) {
    if (flags) CRM_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    dim3 grid((unsigned)((Kp + 127) / 128), (unsigned)((Bp + 127) / 128), 1);
    oz_transpose_i8_kernel<<<grid, 256, 0, st>>>(G8, ld8, n, B, Gt8, G2t8, Bp, Kp, flags);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int oz_launch_widen_i8(const int8_t* G8, long long ld8, long long n, long long B, double* out, long long ldo, cudaStream_t st) {
    const long long total = n * B;
    if (total <= 0) return CRM_OK;
    oz_widen_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(G8, ld8, n, B, out, ldo);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int oz_launch_finite_check(const double* G, long long ldg, long long n, long long B, int* flags, cudaStream_t st) {
    const long long total = n * B;
    if (total <= 0) return CRM_OK;
    oz_finite_check_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(G, ldg, n, B, flags);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

// affine-integer columns (ozaki.cuh): statistics -> lattice (a, b, tolerance) per column -> int8 image of d; scratch = chunks * B OzColStat
size_t oz_affine_scratch_bytes(long long B) { return (size_t)OZ_AFFINE_CHUNKS * (size_t)B * sizeof(OzColStat); }
int oz_launch_affine_genotypes(const double* G, long long ldg, long long n, long long B, void* scratch, double* aff, long long lda, int8_t* Gt8, int8_t* G2t8,
                               long long Bp, long long Kp, int* flags, cudaStream_t st) {
    const int chunks = (int)std::max<long long>(1, std::min<long long>(OZ_AFFINE_CHUNKS, n / 256));
    const long long rows_per_chunk = (n + chunks - 1) / chunks;
    OzColStat* partial = static_cast<OzColStat*>(scratch);
    oz_colstat_kernel<<<dim3((unsigned)((B + 31) / 32), (unsigned)chunks), 256, 0, st>>>(G, ldg, n, B, rows_per_chunk, partial);
    CRM_CUDA(cudaGetLastError()); count_launch();
    oz_colstat_merge_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(partial, chunks, B, aff, lda);
    CRM_CUDA(cudaGetLastError()); count_launch();
    CRM_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    const long long blocks = ((Kp + OZ_ROWS - 1) / OZ_ROWS) * ((Bp + OZ_TILE - 1) / OZ_TILE);
    if (blocks > 2147483647LL) { set_error("genotype block too large for one conversion launch"); return CRM_ERR_UNSUPPORTED; }
    oz_affine_genotype_kernel<<<(unsigned)blocks, 256, 0, st>>>(G, ldg, n, B, aff, lda, Gt8, G2t8, Bp, Kp, flags);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int oz_launch_affine_fix(double* C, long long ldc, long long B, long long cols, const double* aff, long long lda, const double* colsum, cudaStream_t st) {
    oz_affine_fix_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)std::min<long long>(B, 65535)), 256, 0, st>>>(C, ldc, B, cols, aff, lda, colsum);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int oz_launch_affine_fix_square(double* sq, const double* lin, long long ld, long long B, int cols, const double* aff, long long lda, const double* colsum2, cudaStream_t st) {
    oz_affine_fix_square_kernel<<<dim3((unsigned)((cols + 127) / 128), (unsigned)std::min<long long>(B, 65535)), 128, 0, st>>>(sq, lin, ld, B, cols, aff, lda, colsum2);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int oz_launch_combine(const int* D, long long Mp, long long ldd, const int* expo, long long Mtot, long long B, double* C, long long ldc, cudaStream_t st) {
    dim3 grid((unsigned)((B + OZ_TILE - 1) / OZ_TILE), (unsigned)((Mtot + OZ_TILE - 1) / OZ_TILE), 1), block(OZ_TILE, 8, 1);
    oz_combine_kernel<<<grid, block, 0, st>>>(D, Mp, ldd, expo, Mtot, B, C, ldc);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

// D[Mrows][ldd] (int32, row-major) = A8[Mrows][Kp] (int8, K-major) x Gt8[Bp][Kp]^T (int8, K-major): the plain int8
// library GEMM (cuBLASLt "TN" layout, tcgen05 kernels on sm_100).
struct LtContext { cublasLtHandle_t lt = nullptr; void* workspace = nullptr; size_t workspace_bytes = 64ull << 20; };
static LtContext g_lt[16];

int oz_int8_gemm(const int8_t* A8, long long Mrows, const int8_t* Gt8, long long Bp, long long Kp, int* D, long long ldd, cudaStream_t st) {
    int dev = 0;
    CRM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { set_error("device index %d outside the supported range", dev); return CRM_ERR_UNSUPPORTED; }
    LtContext& c = g_lt[dev];
    if (!c.lt) {
        if (cublasLtCreate(&c.lt) != CUBLAS_STATUS_SUCCESS) { set_error("cublasLtCreate failed"); return CRM_ERR_SOLVER; }
        CRM_CUDA(cudaMalloc(&c.workspace, c.workspace_bytes));
    }
    // column-major view: D^T (Bp x Mrows, ld ldd) = op_T(Gt8 as Kp x Bp, ld Kp) * (A8 as Kp x Mrows, ld Kp)
    cublasLtMatmulDesc_t op = nullptr;
    cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
    cublasLtMatmulPreference_t pref = nullptr;
    cublasStatus_t s = cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32I, CUDA_R_32I);
    const cublasOperation_t opT = CUBLAS_OP_T, opN = CUBLAS_OP_N;
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &opT, sizeof(opT));
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &opN, sizeof(opN));
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatrixLayoutCreate(&la, CUDA_R_8I, (uint64_t)Kp, (uint64_t)Bp, Kp);
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatrixLayoutCreate(&lb, CUDA_R_8I, (uint64_t)Kp, (uint64_t)Mrows, Kp);
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatrixLayoutCreate(&lc, CUDA_R_32I, (uint64_t)Bp, (uint64_t)Mrows, ldd);
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatmulPreferenceCreate(&pref);
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &c.workspace_bytes, sizeof(c.workspace_bytes));
    cublasLtMatmulHeuristicResult_t heur{};
    int found = 0;
    if (s == CUBLAS_STATUS_SUCCESS) s = cublasLtMatmulAlgoGetHeuristic(c.lt, op, la, lb, lc, lc, pref, 1, &heur, &found);
    if (s == CUBLAS_STATUS_SUCCESS && found == 0) s = CUBLAS_STATUS_NOT_SUPPORTED;
    const int32_t alpha = 1, beta = 0;
    if (s == CUBLAS_STATUS_SUCCESS)
        s = cublasLtMatmul(c.lt, op, &alpha, Gt8, la, A8, lb, &beta, D, lc, D, lc, &heur.algo, c.workspace, c.workspace_bytes, st);
    if (pref) cublasLtMatmulPreferenceDestroy(pref);
    if (lc) cublasLtMatrixLayoutDestroy(lc);
    if (lb) cublasLtMatrixLayoutDestroy(lb);
    if (la) cublasLtMatrixLayoutDestroy(la);
    if (op) cublasLtMatmulDescDestroy(op);
    if (s != CUBLAS_STATUS_SUCCESS) { set_error("int8 cuBLASLt contraction failed with status %d (M=%lld N=%lld K=%lld)", (int)s, Mrows, Bp, Kp); return CRM_ERR_SOLVER; }
    return CRM_OK;      // a library kernel: not counted by count_launch()
}

}  // namespace crm
