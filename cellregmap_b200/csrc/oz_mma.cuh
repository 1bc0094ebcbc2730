// K0 hot kernel: the int8 contraction of the exact int8 split (ozaki.cuh) on the 5th-generation tensor cores, with the fp64
// recombination of the digit planes fused into its epilogue.
//
//   C[s][col] = 2^(e_col + 1) * sum_t 2^(-7 (t + 1)) * sum_i A8[t][col][i] * Gt8[s][i]
//
// One CTA owns a tile of 128 columns of [Hx | Hx.E_j] (TMEM lanes) x 256 SNPs (TMEM columns) and walks the 8 digit planes in
// 4 passes of 2 planes: both accumulators of a pass (2 x 256 columns = the whole TMEM) share every SNP tile staged in shared
// memory, so one stage of 64 KB (2 x 16 KB digit tiles + 32 KB dosages) feeds 8 tcgen05.mma (M=128, N=256, K=32, kind::i8).
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocation + MMA issue (one lane), warps 2-5 = epilogue
// (tcgen05.ld of both accumulators, Horner step in base 2^-7 in fp64 against the running value in C, final 2^(e+1) scaling).
// Operands are K-major with the 128-byte swizzle: a TMA box of (128 bytes of K) x rows lands as the canonical UMMA layout
// (8-row groups of 1024 bytes), the shared-memory descriptors below describe exactly that.
//
// Every partial sum is an exact integer (|sum| < 2^31 is checked by the caller) and the Horner order is the one of
// oz_combine_kernel, so the result is bit-identical to the cuBLASLt + oz_combine route.
#pragma once
#include "common.cuh"
#include "args.cuh"

namespace crm {

constexpr int OZM_BM = 128;                                   // [Hx|Hx.E] columns per tile  (TMEM lanes)
constexpr int OZM_BN = 256;                                   // SNPs per tile               (TMEM columns per accumulator)
constexpr int OZM_BK = 128;                                   // bytes of K per stage = one swizzle row
constexpr int OZM_UK = 32;                                    // K per tcgen05.mma of kind::i8
constexpr int OZM_STAGES = 3;
constexpr int OZM_THREADS = 192;
constexpr int OZM_A_BYTES = OZM_BM * OZM_BK;                  // 16 KB
constexpr int OZM_B_BYTES = OZM_BN * OZM_BK;                  // 32 KB
constexpr int OZM_STAGE_BYTES = 2 * OZM_A_BYTES + OZM_B_BYTES;
constexpr int OZM_SMEM_BYTES = OZM_STAGES * OZM_STAGE_BYTES + 1024 + 256;
constexpr int OZM_NGROUP = 8;                                 // default number of SNP tiles per rasterisation group
static_assert(OZ_SLICES == 8, "the pass structure below assumes 8 digit planes");

struct OzMmaArgs {
    long long Mp, Mtot, B, ldc;
    int kblocks, m_tiles, n_tiles, ngroup;
    const int* expo;
    double* C;
    // split-K variant (few tiles, long contraction): unit = (tile, K chunk); the int32 partial sums of every plane are added into
    // D32[plane][s][col] (exact: integer addition commutes) and recombined by oz_combine_planes_kernel
    int ksplit, kchunk;     // chunks per tile, K blocks per chunk
    int* D32; long long Bp32;
};

// ---- tcgen05 wrappers (PTX ISA: tcgen05.alloc / mma / commit / ld / fence) ----
__device__ __forceinline__ void tc_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive TMEM columns of this thread's lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile with 128-byte rows and the 128-byte swizzle (the layout a TMA box of
// 128 bytes x rows with CU_TENSOR_MAP_SWIZZLE_128B produces): start address and offsets in units of 16 bytes, 8-row groups
// 1024 bytes apart (stride byte offset), leading byte offset unused for swizzled K-major tiles, descriptor version 1
// (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor of kind::i8: D = s32 (bits 4-5 = 2), A and B signed 8-bit (bits 7-9, 10-12 = 1), both K-major
// (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28.
__device__ __forceinline__ uint32_t oz_instr_desc(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct OzTile { int m0, n0, ncols; };
// unit -> tile: groups of OZM_NGROUP SNP tiles x all column tiles, SNP tile fastest, so that the CTAs in flight (consecutive
// units) share a few dosage panels and a contiguous run of digit-plane panels through L2
__device__ __forceinline__ OzTile oz_unit_tile(const OzMmaArgs& a, int u) {
    const int per_group = a.ngroup * a.m_tiles;
    const int g = u / per_group, r = u - g * per_group;
    const int ng = min(a.ngroup, a.n_tiles - g * a.ngroup);
    const int mt = r / ng, nt = g * a.ngroup + (r - mt * ng);
    OzTile t;
    t.m0 = mt * OZM_BM; t.n0 = nt * OZM_BN;
    const long long left = a.B - (long long)t.n0;
    t.ncols = (int)min((long long)OZM_BN, (left + 15) / 16 * 16);
    return t;
}

template <bool SPLITK>
__device__ __forceinline__ void oz_mma_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const OzMmaArgs& a) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OZM_STAGES * OZM_STAGE_BYTES);
    uint64_t* full = bars;                         // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + OZM_STAGES;           // [STAGES]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * OZM_STAGES;   // MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 1;          // epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = SPLITK ? a.ksplit : 1;
    const int units = a.m_tiles * a.n_tiles * ks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < OZM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                            // ===== TMA producer =====
            uint32_t it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const OzTile t = oz_unit_tile(a, u / ks);
                const int kb0 = SPLITK ? (u % ks) * a.kchunk : 0, kb1 = SPLITK ? min(a.kblocks, kb0 + a.kchunk) : a.kblocks;
                for (int pass = 0; pass < 4; pass++) {
                    const int row1 = (int)((long long)(7 - 2 * pass) * a.Mp) + t.m0, row2 = (int)((long long)(6 - 2 * pass) * a.Mp) + t.m0;
                    for (int kb = kb0; kb < kb1; kb++, it++) {
                        const uint32_t s = it % OZM_STAGES, ph = (it / OZM_STAGES) & 1u;
                        mbar_wait_backoff(&empty[s], ph ^ 1u);
                        uint8_t* st = smem + s * OZM_STAGE_BYTES;
                        mbar_expect_tx(&full[s], OZM_STAGE_BYTES);
                        tma_load_2d(st, &tmA, &full[s], kb * OZM_BK, row1);
                        tma_load_2d(st + OZM_A_BYTES, &tmA, &full[s], kb * OZM_BK, row2);
                        tma_load_2d(st + 2 * OZM_A_BYTES, &tmB, &full[s], kb * OZM_BK, t.n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                            // ===== MMA issue =====
            uint32_t it = 0, acc_it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const OzTile t = oz_unit_tile(a, u / ks);
                const int kb0 = SPLITK ? (u % ks) * a.kchunk : 0, kb1 = SPLITK ? min(a.kblocks, kb0 + a.kchunk) : a.kblocks;
                const uint32_t idesc = oz_instr_desc(OZM_BM, t.ncols);
                for (int pass = 0; pass < 4; pass++, acc_it++) {
                    mbar_wait(tmem_empty, (acc_it & 1u) ^ 1u);
                    tc_fence_after();
                    for (int kb = kb0; kb < kb1; kb++, it++) {
                        const uint32_t s = it % OZM_STAGES, ph = (it / OZM_STAGES) & 1u;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + s * OZM_STAGE_BYTES);
                        const uint64_t d1 = oz_smem_desc(sa), d2 = oz_smem_desc(sa + OZM_A_BYTES), db = oz_smem_desc(sa + 2 * OZM_A_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < OZM_BK / OZM_UK; k4++) {
                            const uint64_t adv = (uint64_t)(k4 * OZM_UK >> 4);       // start-address field advances in 16-byte units
                            const uint32_t acc = ((kb - kb0) | k4) != 0;
                            tc_mma_i8(tmem_base, d1 + adv, db + adv, idesc, acc);
                            tc_mma_i8(tmem_base + OZM_BN, d2 + adv, db + adv, idesc, acc);
                        }
                        tc_commit(&empty[s]);        // frees the stage once these MMAs have read it
                    }
                    tc_commit(tmem_full);
                }
            }
        }
    } else {                                        // ===== epilogue: warps 2..5 own TMEM lane quadrants warp % 4 =====
        const int q = warp & 3;
        uint32_t acc_it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const OzTile t = oz_unit_tile(a, u / ks);
            const long long col = (long long)t.m0 + 32 * q + lane;
            const bool col_ok = col < a.Mtot;
            const int e = (col_ok && !SPLITK) ? a.expo[col] : OZ_EXP_EMPTY;
            for (int pass = 0; pass < 4; pass++, acc_it++) {
                mbar_wait(tmem_full, acc_it & 1u);
                tc_fence_after();
                const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
                for (int ch = 0; ch < OZM_BN / 32; ch++) {
                    const long long s_first = (long long)t.n0 + ch * 32;
                    if (s_first >= a.B) break;
                    uint32_t hi[32], lo[32];
                    tc_ld32(tlane + ch * 32, hi);
                    tc_ld32(tlane + OZM_BN + ch * 32, lo);
                    tc_wait_ld();
                    if (SPLITK) {
                        if (col_ok) {        // lanes = consecutive columns: coalesced integer reductions into D32[plane][s][col]
                            int* dhi = a.D32 + ((long long)(7 - 2 * pass) * a.Bp32 + s_first) * a.Mp + col;
                            int* dlo = a.D32 + ((long long)(6 - 2 * pass) * a.Bp32 + s_first) * a.Mp + col;
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                if (s_first + j < a.B) { atomicAdd(dhi + (long long)j * a.Mp, (int)hi[j]); atomicAdd(dlo + (long long)j * a.Mp, (int)lo[j]); }
                            }
                        }
                    } else if (col_ok) {
                        double* cp = a.C + s_first * a.ldc + col;
                        double acc[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) acc[j] = (pass > 0 && s_first + j < a.B) ? cp[(long long)j * a.ldc] : 0.0;
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            double v = (acc[j] + (double)(int)hi[j]) * 0.0078125;
                            v = (v + (double)(int)lo[j]) * 0.0078125;
                            if (pass == 3) v = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(v, e + 1);
                            if (s_first + j < a.B) cp[(long long)j * a.ldc] = v;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tmem_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_dealloc(tmem_base, 512);
}

__global__ void __launch_bounds__(OZM_THREADS, 1) oz_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OzMmaArgs a) {
    oz_mma_body<false>(tmA, tmB, a);
}
// few output tiles over a long contraction (the g^2 Grams: 232 columns): the K range is cut into chunks, one unit per (tile, chunk)
__global__ void __launch_bounds__(OZM_THREADS, 1) oz_mma_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OzMmaArgs a) {
    oz_mma_body<true>(tmA, tmB, a);
}
// C[s][col] = 2^(e_col + 1) sum_t 2^(-7 (t + 1)) D32[t][s][col]: the Horner order of the fused epilogue, so the result is bit-identical
__global__ void oz_combine_planes_kernel(const int* D32, long long Bp32, long long Mp, const int* expo, long long Mtot, long long B, double* C, long long ldc) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= Mtot) return;
    const int e = expo[col];
    for (long long s = blockIdx.y; s < B; s += gridDim.y) {
        double acc = 0.0;
#pragma unroll
        for (int t = OZ_SLICES - 1; t >= 0; t--) acc = (acc + (double)D32[((long long)t * Bp32 + s) * Mp + col]) * 0.0078125;
        C[s * ldc + col] = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(acc, e + 1);
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2): two CTAs of a cluster share one tile of 256 columns x 256 SNPs.  Each CTA stages its own
// 128 columns of both digit planes and its own half of the SNP tile (48 KB per stage instead of 64 KB: the single-CTA kernel
// sits on the L2 -> SM bandwidth cap, profiles/r01_ncu_oz_mma_kernel.txt), the leader CTA issues tcgen05.mma.cta_group::2
// (M = 256, N = 256) which reads both halves of the SNP tile and writes 128 accumulator lanes into the TMEM of each CTA.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int OZ2_BM = 256;
constexpr int OZ2_STAGES = 4;
constexpr int OZ2_A_BYTES = 128 * OZM_BK;                     // one digit plane, this CTA's 128 columns
constexpr int OZ2_B_BYTES = 128 * OZM_BK;                     // this CTA's half of the SNP tile
constexpr int OZ2_STAGE_BYTES = 2 * OZ2_A_BYTES + OZ2_B_BYTES;
constexpr int OZ2_SMEM_BYTES = OZ2_STAGES * OZ2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc2_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc2_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in both CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA load into this CTA's shared memory, completion counted on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ OzTile oz2_unit_tile(const OzMmaArgs& a, int u) {
    const int per_group = a.ngroup * a.m_tiles;
    const int g = u / per_group, r = u - g * per_group;
    const int ng = min(a.ngroup, a.n_tiles - g * a.ngroup);
    const int mt = r / ng, nt = g * a.ngroup + (r - mt * ng);
    OzTile t;
    t.m0 = mt * OZ2_BM; t.n0 = nt * OZM_BN;
    const long long left = a.B - (long long)t.n0;
    t.ncols = (int)min((long long)OZM_BN, (left + 31) / 32 * 32);
    return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZM_THREADS, 1)
oz_mma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OzMmaArgs a) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OZ2_STAGES * OZ2_STAGE_BYTES);
    uint64_t* full = bars;                         // [STAGES]  TMA (both CTAs) -> MMA; the leader's copy is the one in use
    uint64_t* empty = bars + OZ2_STAGES;           // [STAGES]  MMA -> TMA, one per CTA (multicast commit)
    uint64_t* tmem_full = bars + 2 * OZ2_STAGES;   // MMA -> epilogue, one per CTA (multicast commit)
    uint64_t* tmem_empty = tmem_full + 1;          // epilogues of both CTAs -> MMA (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int units = a.m_tiles * a.n_tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < OZ2_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 256);
        mbar_fence_init();
    }
    if (warp == 1) tc2_alloc(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                            // ===== TMA producer (both CTAs) =====
            uint32_t it = 0;
            for (int u = pair; u < units; u += npairs) {
                const OzTile t = oz2_unit_tile(a, u);
                const int brow = t.n0 + (int)rank * (t.ncols >> 1);
                for (int pass = 0; pass < 4; pass++) {
                    const int row1 = (int)((long long)(7 - 2 * pass) * a.Mp) + t.m0 + (int)rank * 128;
                    const int row2 = (int)((long long)(6 - 2 * pass) * a.Mp) + t.m0 + (int)rank * 128;
                    for (int kb = 0; kb < a.kblocks; kb++, it++) {
                        const uint32_t s = it % OZ2_STAGES, ph = (it / OZ2_STAGES) & 1u;
                        mbar_wait_backoff(&empty[s], ph ^ 1u);
                        uint8_t* st = smem + s * OZ2_STAGE_BYTES;
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * OZ2_STAGE_BYTES);
                        const uint32_t fb = cluster_map(smem_u32(&full[s]), 0);
                        tma2_load_2d(st, &tmA, fb, kb * OZM_BK, row1);
                        tma2_load_2d(st + OZ2_A_BYTES, &tmA, fb, kb * OZM_BK, row2);
                        tma2_load_2d(st + 2 * OZ2_A_BYTES, &tmB, fb, kb * OZM_BK, brow);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {               // ===== MMA issue (leader CTA) =====
            uint32_t it = 0, acc_it = 0;
            for (int u = pair; u < units; u += npairs) {
                const OzTile t = oz2_unit_tile(a, u);
                const uint32_t idesc = oz_instr_desc(OZ2_BM, t.ncols);
                for (int pass = 0; pass < 4; pass++, acc_it++) {
                    mbar_wait(tmem_empty, (acc_it & 1u) ^ 1u);
                    tc_fence_after();
                    for (int kb = 0; kb < a.kblocks; kb++, it++) {
                        const uint32_t s = it % OZ2_STAGES, ph = (it / OZ2_STAGES) & 1u;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + s * OZ2_STAGE_BYTES);
                        const uint64_t d1 = oz_smem_desc(sa), d2 = oz_smem_desc(sa + OZ2_A_BYTES), db = oz_smem_desc(sa + 2 * OZ2_A_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < OZM_BK / OZM_UK; k4++) {
                            const uint64_t adv = (uint64_t)(k4 * OZM_UK >> 4);
                            const uint32_t acc = (kb | k4) != 0;
                            tc2_mma_i8(tmem_base, d1 + adv, db + adv, idesc, acc);
                            tc2_mma_i8(tmem_base + OZM_BN, d2 + adv, db + adv, idesc, acc);
                        }
                        tc2_commit(&empty[s]);
                    }
                    tc2_commit(tmem_full);
                }
            }
        }
    } else {                                        // ===== epilogue (both CTAs): this CTA's 128 accumulator lanes =====
        const int q = warp & 3;
        const uint32_t te = cluster_map(smem_u32(tmem_empty), 0);
        uint32_t acc_it = 0;
        for (int u = pair; u < units; u += npairs) {
            const OzTile t = oz2_unit_tile(a, u);
            const long long col = (long long)t.m0 + (long long)rank * 128 + 32 * q + lane;
            const bool col_ok = col < a.Mtot;
            const int e = col_ok ? a.expo[col] : OZ_EXP_EMPTY;
            for (int pass = 0; pass < 4; pass++, acc_it++) {
                mbar_wait(tmem_full, acc_it & 1u);
                tc_fence_after();
                const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
                for (int ch = 0; ch < OZM_BN / 32; ch++) {
                    const long long s_first = (long long)t.n0 + ch * 32;
                    if (s_first >= a.B) break;
                    uint32_t hi[32], lo[32];
                    tc_ld32(tlane + ch * 32, hi);
                    tc_ld32(tlane + OZM_BN + ch * 32, lo);
                    tc_wait_ld();
                    if (col_ok) {
                        double* cp = a.C + s_first * a.ldc + col;
                        double acc[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) acc[j] = (pass > 0 && s_first + j < a.B) ? cp[(long long)j * a.ldc] : 0.0;
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            double v = (acc[j] + (double)(int)hi[j]) * 0.0078125;
                            v = (v + (double)(int)lo[j]) * 0.0078125;
                            if (pass == 3) v = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(v, e + 1);
                            if (s_first + j < a.B) cp[(long long)j * a.ldc] = v;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_cluster(te);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tc2_dealloc(tmem_base, 512);
}

}  // namespace crm
