// Host-side launcher of K1 (tensor-map construction + launch).
#include <stdlib.h>
#include <algorithm>
#include "gemm.cuh"
#include "launch.cuh"

namespace crm {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D row-major f64 matrix (rows x cols, leading dimension ld doubles) -> tiled tensor map with box (box_cols x box_rows).
inline int make_map_2d(CUtensorMap* map, const double* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return CRM_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 1)) {
        set_error("TMA operand must be 16-byte aligned with an even leading dimension (ptr=%p ld=%lld)", (const void*)ptr, ld);
        return CRM_ERR_INVALID;
    }
    if (box_cols > 256 || box_rows > 256 || (box_cols & 1)) { set_error("bad TMA box %d x %d", box_cols, box_rows); return CRM_ERR_INVALID; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)", (int)r, rows, cols, ld, box_cols, box_rows); return CRM_ERR_CUDA; }
    return CRM_OK;
}

// smallest even pitch >= w whose multiples {0,p,2p,3p} mod 16 are pairwise >= 2 apart (bank-conflict-free fragment reads)
inline int conflict_free_pitch(int w) {
    for (int p = (w + 1) & ~1;; p += 2) {
        int r[4] = {0, p % 16, (2 * p) % 16, (3 * p) % 16};
        bool ok = true;
        for (int i = 0; i < 4 && ok; i++)
            for (int j = i + 1; j < 4; j++) {
                int d = abs(r[i] - r[j]);
                d = d < 16 - d ? d : 16 - d;
                if (d < 2) { ok = false; break; }
            }
        if (ok) return p;
    }
}
inline int expand_box_width(int kexp) { return conflict_free_pitch((GEMM_BN - 1 + kexp - 1) / kexp + 2); }   // +1: even-aligned box start


// accumulator rows per warp in units of 8: 4 -> 16 consumer warps (4 per scheduler), 8 -> 8 consumer warps.
// CRM_GEMM_MT overrides the default at load time (tuning experiments only).
static int gemm_mt_choice() {
    static int mt = -1;
    if (mt < 0) { const char* e = getenv("CRM_GEMM_MT"); mt = (e && atoi(e) == 8) ? 8 : 4; }
    return mt;
}

static thread_local bool g_free_split = false, g_symmetric = false;
void gemm_set_free_split(bool on) { g_free_split = on; }
void gemm_set_symmetric(bool on) { g_symmetric = on; }

template <int MODE, int MT>
int launch_gemm_mode(const GemmOperands& op, GemmArgs args, cudaStream_t stream) {
    CUtensorMap tmA, tmB, tmB2;
    CRM_CHECK(make_map_2d(&tmA, op.A, args.K, op.a_cols, op.lda, GEMM_APITCH, GEMM_BK));
    int stage_bytes = GEMM_BK * GEMM_APITCH * 8;
    if (MODE == GEMM_EXPAND) {
        CRM_CHECK(make_map_2d(&tmB, op.B, args.K, op.b_cols, op.ldb, args.gpitch, GEMM_BK));
        CRM_CHECK(make_map_2d(&tmB2, op.B2, args.K, op.b2_cols, op.ldb2, args.epitch, GEMM_BK));
        stage_bytes += GEMM_BK * (args.gpitch + args.epitch) * 8;
    } else {
        CRM_CHECK(make_map_2d(&tmB, op.B, args.K, op.b_cols, op.ldb, GEMM_APITCH, GEMM_BK));
        stage_bytes += GEMM_BK * GEMM_APITCH * 8;
        if (MODE == GEMM_PRODUCT) {
            CRM_CHECK(make_map_2d(&tmB2, op.B2, args.K, op.b2_cols, op.ldb2, GEMM_APITCH, GEMM_BK));
            stage_bytes += GEMM_BK * GEMM_APITCH * 8;
        } else {
            tmB2 = tmB;
        }
    }
    const int max_smem = 227 * 1024;
    int stages = (max_smem - 256) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) { set_error("GEMM stage of %d bytes does not fit shared memory twice", stage_bytes); return CRM_ERR_UNSUPPORTED; }
    args.stages = stages;
    args.stage_bytes = stage_bytes;
    size_t smem = (size_t)stages * stage_bytes + 2 * stages * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        CRM_CUDA(cudaFuncSetAttribute(crm_gemm_kernel<MODE, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set = true;
    }
    const int m_span = args.m_count + (args.m_begin & 1), n_span = args.n_count + (MODE == GEMM_EXPAND ? 0 : (args.n_begin & 1));
    args.m_tiles = (m_span + GEMM_BM - 1) / GEMM_BM;
    args.n_tiles = (n_span + GEMM_BN - 1) / GEMM_BN;
    if (args.m_tiles == 0 || args.n_tiles == 0) return CRM_OK;
    // Gram of one matrix with itself (set-up): the tiles above the diagonal are mirror images
    const bool sym = g_symmetric && g_free_split && MODE == GEMM_PLAIN && op.A == op.B && op.lda == op.ldb && args.m_begin == 0 && args.n_begin == 0 &&
                     args.m_count == args.n_count && args.m_tiles == args.n_tiles && args.m_tiles > 1 && GEMM_BM == GEMM_BN;
    const long long ctas = sym ? (long long)args.m_tiles * (args.m_tiles + 1) / 2 : (long long)args.m_tiles * args.n_tiles;
    if (ctas > 2000000000LL) { set_error("GEMM of %d x %d tiles is too large for one launch", args.m_tiles, args.n_tiles); return CRM_ERR_UNSUPPORTED; }
    // Narrow outputs over a long contraction leave most SMs idle: cut K into chunks (partials summed in chunk order).
    // The chunking depends only on (m_tiles, K) -- never on the number of SNP columns -- so that a SNP's arithmetic is the
    // same whatever shard it is scanned in (multi-GPU results stay bit-identical to 1-GPU results); the set-up Gram
    // (free_split) has no SNP dimension and picks the split count by wave efficiency.
    args.k_splits = 1; args.k_chunk = args.K; args.partial = nullptr; args.partial_stride = 0;
    {
        const int max_splits = std::min(16, std::max(1, args.K / (16 * GEMM_BK)));
        int splits = 1;
        if (g_free_split && ctas < 2 * 148) {
            double best = 0.0;
            for (int z = sym ? 2 : 1; z <= max_splits; z++) {     // wave efficiency of ctas * z blocks on 148 SMs (sym needs the reduction pass)
                const long long blocks = ctas * z, waves = (blocks + 147) / 148;
                const double eff = (double)blocks / (double)(waves * 148);
                if (eff > best + 1e-9) { best = eff; splits = z; }
            }
        } else if (args.m_tiles <= 2 && MODE != GEMM_EXPAND) {
            splits = max_splits;
        }
        if (splits > 1) {
            const int chunk = (((args.K + splits - 1) / splits) + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
            splits = (args.K + chunk - 1) / chunk;
            args.k_splits = splits; args.k_chunk = chunk;
            args.partial_stride = (long long)args.n_count * args.m_count;
            if (splits > 1) CRM_CUDA(pool_alloc_async((void**)&args.partial, (size_t)splits * args.partial_stride * sizeof(double), stream));
        }
    }
    args.sym = (sym && args.k_splits > 1) ? 1 : 0;
    dim3 grid((unsigned)(args.sym ? ctas : (long long)args.m_tiles * args.n_tiles), (unsigned)args.k_splits, 1);
    crm_gemm_kernel<MODE, MT><<<grid, gemm_threads(MT), smem, stream>>>(tmA, tmB, tmB2, args);
    CRM_CUDA(cudaGetLastError()); count_launch();
    if (args.k_splits > 1) {
        const long long total = (long long)args.n_count * args.m_count;
        crm_gemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(args.partial, args.partial_stride, args.k_splits, args.n_count, args.m_count,
                                                                                 args.out, args.ldc, args.sym);
        CRM_CUDA(cudaGetLastError()); count_launch();
        CRM_CUDA(cudaFreeAsync(args.partial, stream));
    }
    return CRM_OK;
}

// C[n_count][m_count] (ldc) = B[:, n_begin:+n_count]^T A[:, m_begin:+m_count], B built according to `mode`.
int launch_gemm(int mode, const GemmOperands& op, int K, int m_begin, int m_count, int n_begin, int n_count,
                       double* out, long long ldc, int kexp, cudaStream_t stream) {
    GemmArgs a{};
    a.K = K; a.m_begin = m_begin; a.m_count = m_count; a.n_begin = n_begin; a.n_count = n_count;
    a.out = out; a.ldc = ldc; a.kexp = kexp > 0 ? kexp : 1;
    a.gpitch = GEMM_APITCH; a.epitch = GEMM_APITCH;
    if (mode == GEMM_EXPAND) {
        a.gpitch = expand_box_width(a.kexp);
        a.epitch = (int)op.ldb2;
        if (a.epitch < a.kexp + 1) { set_error("Eext needs at least kexp+1 columns"); return CRM_ERR_INVALID; }
        return gemm_mt_choice() == 8 ? launch_gemm_mode<GEMM_EXPAND, 8>(op, a, stream) : launch_gemm_mode<GEMM_EXPAND, 4>(op, a, stream);
    }
    if (mode == GEMM_PRODUCT) return gemm_mt_choice() == 8 ? launch_gemm_mode<GEMM_PRODUCT, 8>(op, a, stream) : launch_gemm_mode<GEMM_PRODUCT, 4>(op, a, stream);
    return gemm_mt_choice() == 8 ? launch_gemm_mode<GEMM_PLAIN, 8>(op, a, stream) : launch_gemm_mode<GEMM_PLAIN, 4>(op, a, stream);
}

// FP64 tensor-core issue-rate probe: register-resident chains of mma.sync m8n8k4 (SASS DMMA.8x8x4), 8 independent accumulator pairs
// per warp, 8 warps per CTA, 2 CTAs per SM -- the denominator of the fp64 rotation's roofline, measured on the device it runs on.
__global__ void __launch_bounds__(256) crm_dmma_rate_kernel(double* out, int iters, double seed) {
    const int lane = threadIdx.x & 31;
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + lane * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = seed - lane * 1e-3 - i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(acc[i][0], acc[i][1], a[i], b[i & 3]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int measure_fp64_tensor_peak(double* tflops, cudaStream_t st) {
    int dev = 0, sms = 0;
    CRM_CUDA(cudaGetDevice(&dev));
    CRM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    CRM_CUDA(pool_alloc_async((void**)&out, (size_t)sms * 2 * 256 * sizeof(double), st));
    const int iters = 20000, grid = sms * 2;
    cudaEvent_t e0, e1;
    CRM_CUDA(cudaEventCreate(&e0)); CRM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {          // first repetition = warm-up
        CRM_CUDA(cudaEventRecord(e0, st));
        crm_dmma_rate_kernel<<<grid, 256, 0, st>>>(out, iters, 1.0);
        CRM_CUDA(cudaGetLastError()); count_launch();
        CRM_CUDA(cudaEventRecord(e1, st));
        CRM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CRM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CRM_CUDA(cudaFreeAsync(out, st));
    const double flop = (double)grid * 8.0 * iters * 8.0 * (2.0 * 8 * 8 * 4);
    *tflops = flop / ((double)best * 1e-3) * 1e-12;
    return CRM_OK;
}

}  // namespace crm
