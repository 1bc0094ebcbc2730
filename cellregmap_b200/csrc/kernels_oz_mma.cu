// Translation unit: tcgen05 int8 contraction with fused fp64 recombination (oz_mma.cuh) -- tensor maps + launch.
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "launch.cuh"
#include "oz_mma.cuh"

namespace crm {

typedef CUresult (*PFN_encodeTiledU8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// K-major int8 matrix (rows x Kp bytes, row stride Kp) -> tensor map with boxes of 128 bytes of K x box_rows rows, 128-byte
// swizzle, out-of-range rows / K bytes read as zero
static int make_map_k_major(CUtensorMap* map, const int8_t* ptr, long long rows, long long Kp, int box_rows) {
    static PFN_encodeTiledU8 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) { set_error("cuTensorMapEncodeTiled entry point not available"); return CRM_ERR_CUDA; }
        enc = reinterpret_cast<PFN_encodeTiledU8>(p);
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (Kp & 15) || box_rows > 256) { set_error("int8 TMA operand must be 16-byte aligned with a row stride that is a multiple of 16 (ptr=%p Kp=%lld)", (const void*)ptr, Kp); return CRM_ERR_INVALID; }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp};
    cuuint32_t box[2] = {(cuuint32_t)OZM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (int8, rows=%lld Kp=%lld box=%d) failed with CUresult %d", rows, Kp, box_rows, (int)r); return CRM_ERR_CUDA; }
    return CRM_OK;
}

// variant < 0: process default -- the single-CTA kernel; CRM_INT8_MMA=2cta selects the CTA-pair kernel, which keeps the tensor
// pipe busier (92 % vs 86 %) but runs into the power cap earlier: 118-143 ms against 101-110 ms at bench size
// (profiles/r01_int8_split_sweep.txt)
static bool oz_use_pairs(int variant) {
    static const bool env_pairs = [] { const char* v = getenv("CRM_INT8_MMA"); return v && !strcmp(v, "2cta"); }();
    return variant < 0 ? env_pairs : variant == 2;
}

int oz_launch_mma(const int8_t* A8, long long Mp, long long Mtot, const int* expo, const int8_t* Gt8, long long Bp, long long B, long long Kp, double* C,
                  long long ldc, cudaStream_t st, int variant) {
    if (B <= 0 || Mtot <= 0) return CRM_OK;
    if ((long long)OZ_SLICES * Mp > 2000000000LL || Kp > 2000000000LL) { set_error("int8 contraction: operand too large for 32-bit TMA coordinates"); return CRM_ERR_UNSUPPORTED; }
    CUtensorMap tmA, tmB;
    CRM_CHECK(make_map_k_major(&tmA, A8, (long long)OZ_SLICES * Mp, Kp, 128));
    const bool pairs = oz_use_pairs(variant);
    CRM_CHECK(make_map_k_major(&tmB, Gt8, Bp, Kp, pairs ? 128 : OZM_BN));
    OzMmaArgs a{};
    a.Mp = Mp; a.Mtot = Mtot; a.B = B; a.ldc = ldc;
    a.kblocks = (int)((Kp + OZM_BK - 1) / OZM_BK);
    a.m_tiles = (int)((Mtot + OZM_BM - 1) / OZM_BM);
    a.n_tiles = (int)((B + OZM_BN - 1) / OZM_BN);
    a.expo = expo; a.C = C;
    { static const int ng = [] { const char* v = getenv("CRM_OZ_NGROUP"); return v && atoi(v) > 0 ? atoi(v) : OZM_NGROUP; }(); a.ngroup = ng; }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        CRM_CUDA(cudaGetDevice(&dev));
        CRM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CRM_CUDA(cudaFuncSetAttribute(oz_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZM_SMEM_BYTES));
    }
    if (pairs) {
        static bool attr2 = false;
        static int max_pairs = 0;
        if (!attr2) {
            CRM_CUDA(cudaFuncSetAttribute(oz_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM_BYTES));
            // CTA pairs that can be resident at once (an SM whose TPC partner is fused off cannot host half a pair)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)sms, 1, 1); cfg.blockDim = dim3(OZM_THREADS, 1, 1); cfg.dynamicSmemBytes = OZ2_SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_pairs, oz_mma2_kernel, &cfg) != cudaSuccess || max_pairs <= 0) { cudaGetLastError(); max_pairs = sms / 2; }
            if (getenv("CRM_TRACE")) fprintf(stderr, "[crm trace] oz_mma2_kernel: %d SMs, %d resident CTA pairs\n", sms, max_pairs);
            attr2 = true;
        }
        a.m_tiles = (int)((Mtot + OZ2_BM - 1) / OZ2_BM);
        const long long units2 = (long long)a.m_tiles * a.n_tiles;
        const unsigned grid2 = 2u * (unsigned)std::min<long long>(units2, max_pairs);
        oz_mma2_kernel<<<grid2, OZM_THREADS, OZ2_SMEM_BYTES, st>>>(tmA, tmB, a);
        CRM_CUDA(cudaGetLastError()); count_launch();
        return CRM_OK;
    }
    const long long units = (long long)a.m_tiles * a.n_tiles;
    // Few tiles over a long contraction (the g^2 Grams: 2 column tiles x the SNP tiles of the batch): one CTA per tile would walk all of K
    // alone and leave SMs idle (2 tiles: 2 ms whatever the batch width; 80 tiles on 148 SMs: 46 % idle) -> cut K into chunks, one unit per
    // (tile, chunk), integer partial sums added in global memory; chunks of >= 16 K-blocks.
    static const bool splitk_on = [] { const char* v = getenv("CRM_OZ_SPLITK"); return !(v && atoi(v) == 0); }();
    int ks = 1;
    if (splitk_on && units < 2 * sms && a.kblocks >= 32) {
        // cost of a launch in K-blocks: waves x (chunk length + ~24 K-blocks' worth of pipeline fill and integer reductions per unit)
        double best = (double)((units + sms - 1) / sms) * (a.kblocks + 24.0);
        for (int cand = 2; cand <= a.kblocks / 16 && (long long)cand * units <= 16LL * sms; cand++) {
            const double cost = (double)((units * cand + sms - 1) / sms) * ((double)a.kblocks / cand + 24.0);
            if (cost < best) { best = cost; ks = cand; }
        }
    }
    if (ks > 1) {
        static bool attr3 = false;
        if (!attr3) { CRM_CUDA(cudaFuncSetAttribute(oz_mma_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZM_SMEM_BYTES)); attr3 = true; }
        a.kchunk = (a.kblocks + ks - 1) / ks;
        ks = (a.kblocks + a.kchunk - 1) / a.kchunk;
        a.ksplit = ks;
        a.Bp32 = B;
        const size_t bytes = (size_t)OZ_SLICES * (size_t)B * (size_t)Mp * sizeof(int);
        CRM_CUDA(pool_alloc_async((void**)&a.D32, bytes, st));
        CRM_CUDA(cudaMemsetAsync(a.D32, 0, bytes, st));
        const unsigned gridk = (unsigned)std::min<long long>(units * ks, sms);
        oz_mma_splitk_kernel<<<gridk, OZM_THREADS, OZM_SMEM_BYTES, st>>>(tmA, tmB, a);
        CRM_CUDA(cudaGetLastError()); count_launch();
        oz_combine_planes_kernel<<<dim3((unsigned)((Mtot + 127) / 128), (unsigned)std::min<long long>(B, 65535)), 128, 0, st>>>(a.D32, a.Bp32, Mp, expo, Mtot, B, C, ldc);
        CRM_CUDA(cudaGetLastError()); count_launch();
        CRM_CUDA(cudaFreeAsync(a.D32, st));
        return CRM_OK;
    }
    const unsigned grid = (unsigned)std::min<long long>(units, sms);
    oz_mma_kernel<<<grid, OZM_THREADS, OZM_SMEM_BYTES, st>>>(tmA, tmB, a);
    CRM_CUDA(cudaGetLastError()); count_launch();
    return CRM_OK;
}

}  // namespace crm
