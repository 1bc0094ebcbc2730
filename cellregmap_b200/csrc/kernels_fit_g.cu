// Translation unit: K2 instantiations for the design [W g].
#include "fit.cuh"
#include "launch.cuh"

namespace crm {

template <int P, bool HAS_G>
int launch_fit_t(const FitArgs& fa, cudaStream_t st) {
    constexpr int C = HAS_G ? P - 1 : P;
    size_t smem = (size_t)(2 + C + (HAS_G ? FIT_WARPS : 0)) * fa.mp * sizeof(double);
    int use_smem = smem <= 200 * 1024 ? 1 : 0;
    static bool attr = false;
    if (!attr) { CRM_CUDA(cudaFuncSetAttribute(crm_fit_kernel<P, HAS_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    dim3 grid((fa.p + FIT_WARPS - 1) / FIT_WARPS, fa.R, 1);
    crm_fit_kernel<P, HAS_G><<<grid, FIT_WARPS * 32, use_smem ? smem : 0, st>>>(fa, use_smem);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

#ifndef CRM_FIT_NULL_TU
template <int C>
static int launch_fit_table_t(const FitArgs& fa, double* tab, cudaStream_t st) {
    crm_fit_table_kernel<C><<<dim3((unsigned)fa.tab_k, 2, (unsigned)fa.R), 32, 0, st>>>(fa, tab);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}
size_t fit_table_bytes(int R, int mp) { return (size_t)R * 2 * FIT_TAB_K * (size_t)fit_tab_record(mp) * sizeof(double); }
int fit_table_points() { return FIT_TAB_K; }
// fa.tab_k must be set (fit_table_points()); fills `tab` (fit_table_bytes) for the S / yr / Wr of fa
int launch_fit_table(const FitArgs& fa, double* tab, cudaStream_t st) {
    switch (fa.c) {
        case 1: return launch_fit_table_t<1>(fa, tab, st);
        case 2: return launch_fit_table_t<2>(fa, tab, st);
        case 3: return launch_fit_table_t<3>(fa, tab, st);
        case 4: return launch_fit_table_t<4>(fa, tab, st);
        case 5: return launch_fit_table_t<5>(fa, tab, st);
        case 6: return launch_fit_table_t<6>(fa, tab, st);
        case 7: return launch_fit_table_t<7>(fa, tab, st);
    }
    set_error("fit table: %d covariate columns outside the compiled range (1..7)", fa.c);
    return CRM_ERR_UNSUPPORTED;
}
int launch_fit_with_g(const FitArgs& fa, cudaStream_t st) {
    switch (fa.c + 1) {
        case 1: return launch_fit_t<1, true>(fa, st);
        case 2: return launch_fit_t<2, true>(fa, st);
        case 3: return launch_fit_t<3, true>(fa, st);
        case 4: return launch_fit_t<4, true>(fa, st);
        case 5: return launch_fit_t<5, true>(fa, st);
        case 6: return launch_fit_t<6, true>(fa, st);
        case 7: return launch_fit_t<7, true>(fa, st);
        case 8: return launch_fit_t<8, true>(fa, st);
    }
    set_error("fixed-effect design with %d columns is outside the compiled range (1..8)", fa.c + 1);
    return CRM_ERR_UNSUPPORTED;
}
#else
int launch_fit_null(const FitArgs& fa, cudaStream_t st) {
    switch (fa.c) {
        case 1: return launch_fit_t<1, false>(fa, st);
        case 2: return launch_fit_t<2, false>(fa, st);
        case 3: return launch_fit_t<3, false>(fa, st);
        case 4: return launch_fit_t<4, false>(fa, st);
        case 5: return launch_fit_t<5, false>(fa, st);
        case 6: return launch_fit_t<6, false>(fa, st);
        case 7: return launch_fit_t<7, false>(fa, st);
    }
    set_error("fixed-effect design with %d columns is outside the compiled range (1..7)", fa.c);
    return CRM_ERR_UNSUPPORTED;
}
#endif

}  // namespace crm
