// Translation unit: K4 (Davies / modified Liu p-values, LRT p-values).
#include "pvalue.cuh"
#include "launch.cuh"

namespace crm {

int launch_pvalues(const PvalArgs& pa, cudaStream_t st) {
    if (pa.count <= 0) return CRM_OK;
    crm_pvalue_kernel<<<(pa.count + PV_WARPS - 1) / PV_WARPS, PV_WARPS * 32, 0, st>>>(pa);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int launch_lrt(const double* alt_lml, double null_lml, long long count, double* pv, cudaStream_t st) {
    if (count <= 0) return CRM_OK;
    crm_lrt_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(alt_lml, null_lml, (int)count, pv);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int launch_lrt_dof(const double* alt_lml, double null_lml, long long count, double dof, double* pv, cudaStream_t st) {
    if (count <= 0) return CRM_OK;
    crm_lrt_dof_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(alt_lml, null_lml, (int)count, dof, pv);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

int launch_liu_params(const double* Q, const double* lam, const int* nlam, int lam_ld, long long count, double* out, cudaStream_t st) {
    if (count <= 0) return CRM_OK;
    crm_liu_params_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(Q, lam, nlam, lam_ld, (int)count, out);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
int launch_qmin(const double* params, int nrho, long long count, double* out, cudaStream_t st) {
    if (count <= 0 || nrho <= 0) return CRM_OK;
    crm_qmin_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(params, nrho, (int)count, out);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

}  // namespace crm
