// Translation unit: K5 (effect-size model of predict_interaction / estimate_betas).
#include <algorithm>
#include "betas.cuh"
#include "launch.cuh"

namespace crm {

int launch_beta_fit(const BetaArgs& ba, cudaStream_t st) {
    const int P = ba.c + (ba.has_g ? 1 : 0) + ba.k0, NZ = 1 + P + ba.k0, nzm = 1 + P;
    if (NZ > BETA_MAX_NZ || P > 64 || P < 1) { set_error("design of %d columns (+ %d random-effect columns) exceeds the compiled limit of %d in total", P, ba.k0, BETA_MAX_NZ - 1); return CRM_ERR_UNSUPPORTED; }
    if (ba.p <= 0) return CRM_OK;
    size_t doubles = (size_t)3 * NZ * NZ + (size_t)NZ * (BETA_CHUNK + 1) + BETA_CHUNK + (size_t)P * P + (size_t)ba.k0 * ba.k0 +
                     (size_t)ba.k0 * nzm + (size_t)nzm * nzm + (size_t)P * P + P + (P + (size_t)P * P) + 8;
    size_t smem = doubles * sizeof(double);
    static bool attr = false;
    if (!attr) { CRM_CUDA(cudaFuncSetAttribute(crm_beta_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    if (smem > 220 * 1024) { set_error("effect-size kernel needs %zu bytes of shared memory", smem); return CRM_ERR_UNSUPPORTED; }
    dim3 grid((unsigned)ba.p, (unsigned)ba.R, 1);
    crm_beta_fit_kernel<<<grid, BETA_THREADS, smem, st>>>(ba);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int launch_beta_gxe(const double* E0, long long lde0, const double* coef, int k0, long long n, long long p, double* out,
                    long long ldo, long long s0, cudaStream_t st) {
    if (p <= 0 || n <= 0) return CRM_OK;
    dim3 grid((unsigned)((p + 127) / 128), (unsigned)std::min<long long>(2048, (n + 1) / 2), 1);
    crm_beta_gxe_kernel<<<grid, 256, (size_t)128 * k0 * sizeof(double), st>>>(E0, lde0, coef, k0, n, p, out, ldo, s0);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

}  // namespace crm
