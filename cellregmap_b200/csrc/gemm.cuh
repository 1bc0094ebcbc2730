// K1: FP64 tensor-core contraction  C[N][M] = B^T A  over a long K (= cells) dimension.
//
// Replaces, for a whole batch of SNPs at once, the rotations the reference repeats per (SNP, rho1):
// glimix_core LMM.__init__ (Q0'y, Q0'X; call site cellregmap/_cellregmap.py:351) and the Q0 products of
// QSCov.solve / PMat.dot (cellregmap/_math.py:72-73,89,92-93) on g and g.E0 (_cellregmap.py:415).
//
//   A : K x M row-major (basis columns, e.g. [H | y | W]);       TMA 2D box (132 x BK), pitch 132 doubles
//   B : built per mode from K-outer row-major inputs:
//       PLAIN   B[k][n] = G[k][n]
//       PRODUCT B[k][n] = G[k][n] * G2[k][n]                      (g^2 and permuted-genotype products)
//       EXPAND  B[k][s*kexp + j] = G[k][s] * Eext[k][j]           (g and g.E_j formed on the fly; Eext = [1|E|0])
//   C : N-major, M contiguous:  C[(n - n_begin) * ldc + (m - m_begin)]
//
// CTA tile 128(M) x 128(N), BK = 32 rows of K per stage, 8 consumer warps (each 64 x 32 accumulators in
// registers, DMMA.8x8x4) + 1 TMA producer warp; full/empty mbarrier ring.
#pragma once
#include "common.cuh"
#include "launch.cuh"

namespace crm {


constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;
constexpr int GEMM_BK = 32;
constexpr int GEMM_APITCH = 132;      // doubles; 132 % 16 == 4 -> conflict-free DMMA fragment loads
// consumer warps: (128 / (8 MT)) along M x 4 along N, each owning MT m8-tiles x 4 n8-tiles of accumulators
__host__ __device__ constexpr int gemm_consumer_warps(int MT) { return (GEMM_BM / (8 * MT)) * 4; }
__host__ __device__ constexpr int gemm_threads(int MT) { return (gemm_consumer_warps(MT) + 1) * 32; }

struct GemmArgs {
    int K;         // contraction length (rows of A and of G)
    int m_begin;   // first column of A used
    int m_count;
    int n_begin;   // first flattened output row
    int n_count;
    double* out;
    long long ldc;
    int kexp;      // EXPAND: columns per SNP = 1 + k ; otherwise 1
    int gpitch;    // EXPAND: pitch (doubles) of the genotype tile rows = TMA box width
    int epitch;    // EXPAND: pitch (doubles) of the Eext tile rows (= padded width of Eext)
    int stages;
    int stage_bytes;
    int m_tiles, n_tiles;   // grid extent in tiles; the 1-D block index is rasterised in groups of GEMM_RASTER_N n-tiles
    int k_splits;           // > 1: the K range is cut into k_splits chunks of k_chunk rows (small grids); chunk `z` writes
    int k_chunk;            //      its partial tile to partial + z * partial_stride (packed, ld = m_count), summed afterwards
    double* partial;
    long long partial_stride;
    int sym;                // 1: A and B are the same matrix (Gram): only tiles with m_tile >= n_tile are computed (k_splits > 1 only; the
                            //    reduction mirrors them)
};
constexpr int GEMM_RASTER_N = 8;

template <int MODE, int MT>
__global__ void __launch_bounds__(gemm_threads(MT), 1)
crm_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmB2, const GemmArgs args) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int GEMM_CONSUMER_WARPS = gemm_consumer_warps(MT);
    constexpr int WARPS_M = GEMM_BM / (8 * MT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int stages = args.stages;
    const int split = blockIdx.y;
    const int k_begin = split * args.k_chunk;
    const int k_len = min(args.k_chunk, args.K - k_begin);
    const int numK = (k_len + GEMM_BK - 1) / GEMM_BK;
    // TMA needs the innermost box coordinate on a 16-byte boundary: tile origins sit on even columns, the odd
    // leading column of a range (if any) is computed and discarded.
    // Rasterisation: consecutive block indices walk GEMM_RASTER_N n-tiles for one m-tile, then the next m-tile, so the
    // ~148 CTAs in flight share each A panel among up to 8 of them and each B panel among ~18 (L2 reuse both ways).
    int m_tile, n_tile;
    if (args.sym) {         // lower triangle of the tile grid, row by row
        const int L = blockIdx.x;
        int r = (int)((sqrtf(8.0f * (float)L + 1.0f) - 1.0f) * 0.5f);
        while (r * (r + 1) / 2 > L) r--;
        while ((r + 1) * (r + 2) / 2 <= L) r++;
        m_tile = r; n_tile = L - r * (r + 1) / 2;
    } else {
        const int L = blockIdx.x, per_group = GEMM_RASTER_N * args.m_tiles;
        const int group = L / per_group, within = L - group * per_group;
        const int gn = min(GEMM_RASTER_N, args.n_tiles - group * GEMM_RASTER_N);
        m_tile = within / gn;
        n_tile = group * GEMM_RASTER_N + (within - m_tile * gn);
    }
    const int m_tile0 = (args.m_begin & ~1) + m_tile * GEMM_BM;
    const int n_tile0 = (MODE == GEMM_EXPAND ? args.n_begin : (args.n_begin & ~1)) + n_tile * GEMM_BN;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * args.stage_bytes);
    uint64_t* empty = full + stages;

    constexpr int A_BYTES = GEMM_BK * GEMM_APITCH * 8;
    const int b_bytes = (MODE == GEMM_EXPAND) ? GEMM_BK * args.gpitch * 8 : A_BYTES;
    const int b2_bytes = (MODE == GEMM_EXPAND) ? GEMM_BK * args.epitch * 8 : (MODE == GEMM_PRODUCT ? A_BYTES : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GEMM_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == GEMM_CONSUMER_WARPS) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            if (MODE != GEMM_PLAIN) tma_prefetch_desc(&tmB2);
            const int s_first = (MODE == GEMM_EXPAND) ? ((n_tile0 / args.kexp) & ~1) : n_tile0;
            int stage = 0;
            uint32_t phase = 0;
            for (int kt = 0; kt < numK; kt++) {
                mbar_wait_backoff(&empty[stage], phase ^ 1);
                unsigned char* base = smem + (size_t)stage * args.stage_bytes;
                mbar_expect_tx(&full[stage], (uint32_t)(A_BYTES + b_bytes + b2_bytes));
                tma_load_2d(base, &tmA, &full[stage], m_tile0, k_begin + kt * GEMM_BK);
                tma_load_2d(base + A_BYTES, &tmB, &full[stage], s_first, k_begin + kt * GEMM_BK);
                if (MODE == GEMM_PRODUCT) tma_load_2d(base + A_BYTES + b_bytes, &tmB2, &full[stage], s_first, k_begin + kt * GEMM_BK);
                if (MODE == GEMM_EXPAND) tma_load_2d(base + A_BYTES + b_bytes, &tmB2, &full[stage], 0, k_begin + kt * GEMM_BK);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------- consumers: warp tile 8 MT (M) x 32 (N) ----------------
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int a_off = t * GEMM_APITCH + wm * (8 * MT) + g;  // + 8*i, + 4*ks*APITCH
    int b_off[4], e_off[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int ncol = wn * 32 + 8 * j + g;   // column inside the CTA tile
        if (MODE == GEMM_EXPAND) {
            const int n = n_tile0 + ncol;
            const int s = n / args.kexp;
            int e = n - s * args.kexp;
            if (n >= args.n_begin + args.n_count) e = args.kexp;   // zero column of Eext
            b_off[j] = t * args.gpitch + (s - ((n_tile0 / args.kexp) & ~1));
            e_off[j] = t * args.epitch + e;
        } else {
            b_off[j] = t * GEMM_APITCH + ncol;
            e_off[j] = b_off[j];
        }
    }
    const int b_step = 4 * ((MODE == GEMM_EXPAND) ? args.gpitch : GEMM_APITCH);
    const int e_step = 4 * ((MODE == GEMM_EXPAND) ? args.epitch : GEMM_APITCH);

    double acc[MT][4][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    int stage = 0;
    uint32_t phase = 0;
    for (int kt = 0; kt < numK; kt++) {
        mbar_wait(&full[stage], phase);
        const unsigned char* base = smem + (size_t)stage * args.stage_bytes;
        const double* As = reinterpret_cast<const double*>(base) + a_off;
        const double* Bs = reinterpret_cast<const double*>(base + A_BYTES);
        const double* B2s = reinterpret_cast<const double*>(base + A_BYTES + b_bytes);
#pragma unroll
        for (int ks = 0; ks < GEMM_BK / 4; ks++) {
            double a[MT], b[4];
#pragma unroll
            for (int i = 0; i < MT; i++) a[i] = As[ks * 4 * GEMM_APITCH + 8 * i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double v = Bs[b_off[j] + ks * b_step];
                if (MODE != GEMM_PLAIN) v *= B2s[e_off[j] + ks * e_step];
                b[j] = v;
            }
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
    }

    // ---------------- epilogue: D[m][n] -> out[(n - n_begin) * ldc + (m - m_begin)] ----------------
    const int m_end = args.m_begin + args.m_count, n_end = args.n_begin + args.n_count;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int n = n_tile0 + wn * 32 + 8 * j + 2 * t + h;
            if (n >= n_end || n < args.n_begin) continue;
            double* row = (args.k_splits > 1 ? args.partial + (long long)split * args.partial_stride + (long long)(n - args.n_begin) * args.m_count
                                             : args.out + (long long)(n - args.n_begin) * args.ldc) - args.m_begin;
#pragma unroll
            for (int i = 0; i < MT; i++) {
                const int m = m_tile0 + wm * (8 * MT) + 8 * i + g;
                if (m < m_end && m >= args.m_begin) row[m] = acc[i][j][h];
            }
        }
    }
}

// out[n][m] = sum over the K chunks of partial[z][n][m], in chunk order (deterministic); sym: entries of tiles above the diagonal of
// the tile grid are taken from their mirror image (only the lower tiles were computed)
__global__ void crm_gemm_reduce_kernel(const double* partial, long long stride, int splits, int n_count, int m_count, double* out, long long ldc, int sym) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_count * m_count) return;
    const long long n = idx / m_count; const int m = (int)(idx - n * m_count);
    long long src = idx;
    if (sym && (m / GEMM_BM) < (int)(n / GEMM_BN)) src = (long long)m * m_count + n;
    double s = 0.0;
    for (int z = 0; z < splits; z++) s += partial[(long long)z * stride + src];
    out[n * ldc + m] = s;
}

}  // namespace crm
