// Translation unit: K3 (selection, grouping, gather/transpose, score statistic).
#include "score.cuh"
#include "launch.cuh"

namespace crm {

template <int P>
static int launch_score_t(const ScoreArgs& sa, long long count, cudaStream_t st) {
    const int NZ = 1 + P + sa.k;
    size_t smem = ((size_t)NZ * NZ + (size_t)NZ * (SCORE_CHUNK + 1) + SCORE_CHUNK + (size_t)sa.k * sa.k + (size_t)P * (1 + sa.k) + 2 * sa.k + 8) * sizeof(double);
    static bool attr = false;
    if (!attr) { CRM_CUDA(cudaFuncSetAttribute(crm_score_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
    if (smem > 160 * 1024) { set_error("score kernel needs %zu bytes of shared memory", smem); return CRM_ERR_UNSUPPORTED; }
    crm_score_kernel<P><<<(unsigned)count, SCORE_THREADS, smem, st>>>(sa);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int launch_score(const ScoreArgs& sa, long long count, cudaStream_t st) {
    const int P = sa.c + 1;
    if (1 + P + sa.k > SCORE_MAX_NZ) { set_error("1 + covariates + 1 + contexts = %d exceeds the compiled limit %d", 1 + P + sa.k, SCORE_MAX_NZ); return CRM_ERR_UNSUPPORTED; }
    if (count <= 0) return CRM_OK;
    switch (P) {
        case 2: return launch_score_t<2>(sa, count, st);
        case 3: return launch_score_t<3>(sa, count, st);
        case 4: return launch_score_t<4>(sa, count, st);
        case 5: return launch_score_t<5>(sa, count, st);
        case 6: return launch_score_t<6>(sa, count, st);
        case 7: return launch_score_t<7>(sa, count, st);
        case 8: return launch_score_t<8>(sa, count, st);
    }
    set_error("fixed-effect design with %d columns is outside the compiled range (2..8)", P);
    return CRM_ERR_UNSUPPORTED;
}
int launch_select(const double* lml, const double* delta, const double* scale, int p, int R, int* rho_idx, double* best_lml,
                  double* v0, double* v1, cudaStream_t st) {
    if (p <= 0) return CRM_OK;
    crm_select_kernel<<<(p + 255) / 256, 256, 0, st>>>(lml, delta, scale, p, R, rho_idx, best_lml, v0, v1);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int launch_group(const int* rho_idx, int p, int R, int* perm, int* offsets, cudaStream_t st) {
    crm_group_kernel<<<1, 1024, 0, st>>>(rho_idx, p, R, perm, offsets);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int launch_gather_transpose(const double* C, long long ldc, const int* perm, int kexp, int joff, int kcols, long long nq,
                            int na, double* out, long long ldo, cudaStream_t st) {
    if (nq == 0 || na == 0) return CRM_OK;
    dim3 grid((unsigned)((nq + 31) / 32), (unsigned)((na + 31) / 32), 1), block(32, 8, 1);
    crm_gather_transpose_kernel<<<grid, block, 0, st>>>(C, ldc, perm, kexp, joff, kcols, nq, na, out, ldo);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

}  // namespace crm
