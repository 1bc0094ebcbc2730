// Translation unit: K3 (selection, grouping, gather/transpose, score statistic).
#include <stdlib.h>
#include <string.h>

#include "score.cuh"
#include "launch.cuh"

namespace crm {

int launch_score(const ScoreArgs& sa, long long count, cudaStream_t st) {
    const int P = sa.c + 1, NZ = 1 + P + sa.k;
    if (NZ > SCORE_MAX_NZ) { set_error("1 + covariates + 1 + contexts = %d exceeds the compiled limit %d", NZ, SCORE_MAX_NZ); return CRM_ERR_UNSUPPORTED; }
    if (count <= 0) return CRM_OK;
    {   // warp-per-SNP kernel (tensor-core Gram) when the Gram fits 3 or 4 blocks of 8 columns; CRM_SCORE=cta forces the CTA-per-SNP kernel
        static const bool cta_only = [] { const char* v = getenv("CRM_SCORE"); return v && !strcmp(v, "cta"); }();
        const int nb = (NZ + 7) / 8;
        const size_t wsmem = SCOREW_WARPS * scorew_warp_doubles(NZ, P, sa.k) * sizeof(double);
        if (!cta_only && nb <= 4 && wsmem <= 200 * 1024) {
            static bool wattr = false;
            if (!wattr) {
                CRM_CUDA(cudaFuncSetAttribute(crm_score_warp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CRM_CUDA(cudaFuncSetAttribute(crm_score_warp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                wattr = true;
            }
            ScoreArgs wa = sa;
            wa.p = (int)count;
            const unsigned grid = (unsigned)((count + SCOREW_WARPS - 1) / SCOREW_WARPS);
            if (nb <= 3) crm_score_warp_kernel<3><<<grid, SCOREW_WARPS * 32, wsmem, st>>>(wa);
            else crm_score_warp_kernel<4><<<grid, SCOREW_WARPS * 32, wsmem, st>>>(wa);
            CRM_CUDA(cudaGetLastError()); count_launch();
            return CRM_OK;
        }
    }
    size_t smem = ((size_t)NZ * NZ + (size_t)NZ * (SCORE_CHUNK + 1) + SCORE_CHUNK + (size_t)sa.k * sa.k + 2 * (size_t)P * (1 + sa.k) + 2 * sa.k +
                   2 * (size_t)P * P + 8) * sizeof(double);
    static bool attr = false;
    if (!attr) { CRM_CUDA(cudaFuncSetAttribute(crm_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    if (smem > 200 * 1024) { set_error("score kernel needs %zu bytes of shared memory", smem); return CRM_ERR_UNSUPPORTED; }
    crm_score_kernel<<<(unsigned)count, SCORE_THREADS, smem, st>>>(sa);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
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
