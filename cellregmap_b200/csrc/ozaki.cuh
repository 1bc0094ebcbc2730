// Exact integer split of the fp64 rotation (Ozaki scheme) for integer dosages.
//
// The rotation C[s][col] = sum_n HxE[n][col] * G[n][s] (see launch_rotation in abi.cu) has one operand that is exactly
// representable in int8 whenever the genotypes are integer dosages (0/1/2, or any integer in [-127, 127]).  The other
// operand is split once per gene into SLICES int8 digit planes per column,
//     HxE[n][col] = 2^(e_col + 1) * sum_t q_t[n][col] 2^(-7 (t + 1))   (+ a remainder below 2^(e_col - 56)),
// so that every slice product q_t' g is an exact int8 x int8 -> int32 tensor-core contraction (tcgen05 on sm_100) and
//     C[s][col] = 2^(e_col + 1) * sum_t 2^(-7 (t + 1)) D_t[col][s]
// is assembled in fp64.  Eight 7-bit digits cover 56 bits below the column maximum: the result carries the rounding of
// one fp64 sum of eight terms, like the DMMA route carries the rounding of its fp64 accumulation.
#pragma once
#include "common.cuh"
#include "args.cuh"

namespace crm {

constexpr int OZ_TILE = 32;



// One balanced base-128 digit of x (|x| <= 1/2): d = rint(128 x), x <- 128 x - d.  rint and the float -> int conversion are both
// taken from one addition of 1.5 * 2^52 (round-to-nearest-even puts the integer into the low mantissa bits): conversion
// instructions run at a fraction of the FP64 add rate, and this loop is the whole cost of the digit planes.
__device__ __forceinline__ int oz_next_digit(double& x) {
    const double t = x * 128.0;                                  // exact
    const double s = __dadd_rn(t, 6755399441055744.0);
    const int q = __double2loint(s);
    x = t - __dadd_rn(s, -6755399441055744.0);                   // exact: |t - d| <= 1/2
    return q;
}

// exponent e with |x| < 2^e for the largest |x| of each product column F[:, j0 + jj] * X[:, a]  (a < cols, jj < gridDim.y), stored at
// expo[row0 + jj * rstride + a]; the expanded basis [Hx | Hx.E0_j] is X = Hx, F = Eext, rstride = ldH.  expo must be pre-filled with
// OZ_EXP_EMPTY; cells are split over blockIdx.z and merged with atomicMax
__global__ void oz_column_exponent_kernel(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, long long n, int* expo, long long row0,
                                          long long rstride) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int jj = blockIdx.y;
    if (a >= cols) return;
    const long long chunk = (n + gridDim.z - 1) / gridDim.z, i0 = (long long)blockIdx.z * chunk, i1 = min(n, i0 + chunk);
    double mx = 0.0;
    for (long long i = i0; i < i1; i++) mx = fmax(mx, fabs(F[i * ldf + j0 + jj] * X[i * ldx + a]));
    if (mx > 0.0) {
        int e = 0;
        frexp(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)  ->  mx < 2^e
        atomicMax(&expo[row0 + (long long)jj * rstride + a], e);
    }
}

constexpr int OZ_ROWS = 128;               // cells per block of the slicing kernel (4 per thread -> char4 stores)

// Shared-memory tile of OZ_ROWS cells x OZ_TILE columns used to turn row-major reads (lanes along columns) into K-major writes (each
// thread 4 consecutive cells of one column).  Row r of the tile is stored at slot (r % 4) * 32 + r / 4: the four cells 4 tx + u of
// lane tx then sit 32 slots apart and a warp's read of one u touches 32 consecutive slots of one column -- with a pitch of 33 doubles
// that is conflict-free, where the identity mapping is an 8-way bank conflict (ncu: 429 M conflicts per launch, profiles/r02_ncu_oz_slice_kernel.txt).
__device__ __forceinline__ int oz_tile_slot(int r) { return ((r & 3) << 5) | (r >> 2); }

// digit planes, K-major: A8[t][col][i] (plane stride = Mp * Kp, row stride = Kp, Kp a multiple of 4) of the product columns
// F[:, j0 + jj] * X[:, a], col = row0 + jj * rstride + a  (a < cols, jj < nj)
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* X, long long ldx, int cols, const double* F, long long ldf, int j0, int nj, long long n,
                                                       const int* expo, int8_t* A8, long long Mp, long long Kp, long long row0, long long rstride) {
    __shared__ double tile[OZ_ROWS][OZ_TILE + 1];
    const long long i0 = (long long)blockIdx.x * OZ_ROWS;
    const int a0 = blockIdx.y * OZ_TILE;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    // this thread's 16 entries of the X tile stay in registers for all factor columns (the basis is read once, not 1 + k times)
    double hx[OZ_ROWS / 8];
#pragma unroll
    for (int q = 0; q < OZ_ROWS / 8; q++) {
        const long long i = i0 + ty + 8 * q; const int a = a0 + tx;
        hx[q] = (i < n && a < cols) ? X[i * ldx + a] : 0.0;
    }
    for (int jj = 0; jj < nj; jj++) {
        const int j = j0 + jj;
#pragma unroll
        for (int q = 0; q < OZ_ROWS / 8; q++) {                   // rows i, columns a (coalesced along a); F[i][j] is a warp-wide broadcast
            const int r = ty + 8 * q; const long long i = i0 + r;
            tile[oz_tile_slot(r)][tx] = (i < n) ? F[i * ldf + j] * hx[q] : 0.0;
        }
        __syncthreads();
        for (int r = ty; r < OZ_TILE; r += 8) {                   // write: rows a, 4 consecutive cells per thread
            const int a = a0 + r; const long long i = i0 + 4 * tx;
            if (a < cols && i < Kp) {
                const long long col = row0 + (long long)jj * rstride + a;
                const int e = expo[col];
                const double scale = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(1.0, -(e + 1));     // exact power of two
                double x[4];
#pragma unroll
                for (int u = 0; u < 4; u++) x[u] = tile[(u << 5) | tx][r] * scale;           // cell 4 tx + u; |x| < 1/2
#pragma unroll
                for (int t = 0; t < OZ_SLICES; t++) {
                    char4 q;
                    signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
                    for (int u = 0; u < 4; u++) qq[u] = (signed char)oz_next_digit(x[u]);
                    *reinterpret_cast<char4*>(A8 + (long long)t * Mp * Kp + col * Kp + i) = q;
                }
            }
        }
        __syncthreads();
    }
}

// the same two steps for an explicit n x cols matrix X (row stride ldx): exponents, then digit planes P8[t][col][i]
__global__ void oz_matrix_exponent_kernel(const double* X, long long ldx, int cols, long long n, int* expo) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cols) return;
    const long long chunk = (n + gridDim.y - 1) / gridDim.y, i0 = (long long)blockIdx.y * chunk, i1 = min(n, i0 + chunk);
    double mx = 0.0;
    for (long long i = i0; i < i1; i++) mx = fmax(mx, fabs(X[i * ldx + a]));
    if (mx > 0.0) { int e = 0; frexp(mx, &e); atomicMax(&expo[a], e); }
}
__global__ void __launch_bounds__(256) oz_matrix_slice_kernel(const double* X, long long ldx, int cols, long long n, const int* expo, int8_t* P8, long long Mp,
                                                              long long Kp) {
    __shared__ double tile[OZ_ROWS][OZ_TILE + 1];
    const long long i0 = (long long)blockIdx.x * OZ_ROWS;
    const int a0 = blockIdx.y * OZ_TILE;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < OZ_ROWS; r += 8) {
        const long long i = i0 + r; const int a = a0 + tx;
        tile[oz_tile_slot(r)][tx] = (i < n && a < cols) ? X[i * ldx + a] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < OZ_TILE; r += 8) {
        const int a = a0 + r; const long long i = i0 + 4 * tx;
        if (a < cols && i < Kp) {
            const int e = expo[a];
            const double scale = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(1.0, -(e + 1));
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = tile[(u << 5) | tx][r] * scale;
#pragma unroll
            for (int t = 0; t < OZ_SLICES; t++) {
                char4 q;
                signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
                for (int u = 0; u < 4; u++) qq[u] = (signed char)oz_next_digit(x[u]);
                *reinterpret_cast<char4*>(P8 + (long long)t * Mp * Kp + (long long)a * Kp + i) = q;
            }
        }
    }
}

// Gt8[s][i] = (int8) G[i][s] for a block of SNP columns (K-major), zero padded to Bp x Kp; flags[0] |= 1 when some entry is
// not an integer in [-127, 127]; flags[1] = max |g|; flags[2] |= 1 when some entry is not finite
// G2t8 (may be null) receives the squares g^2 (valid when max |g| <= 11).  128 cells x 32 SNPs per block, char4 stores.
__global__ void __launch_bounds__(256) oz_genotype_kernel(const double* G, long long ldg, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp,
                                                          long long Kp, int* flags) {
    __shared__ double tile[OZ_ROWS][OZ_TILE + 1];
    __shared__ int s_bad, s_max;
    // 1-D grid, SNP tiles fastest: blocks in flight together read neighbouring 256-byte pieces of the same genotype rows (DRAM pages)
    const long long s_tiles = (Bp + OZ_TILE - 1) / OZ_TILE;
    const long long i0 = ((long long)blockIdx.x / s_tiles) * OZ_ROWS;
    const long long s0 = ((long long)blockIdx.x % s_tiles) * OZ_TILE;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_bad = 0; s_max = 0; }
    __syncthreads();
    int bad = 0, gmax = 0;
    {
        // all 16 loads of the thread are issued before the first value is looked at (the kernel is bound by bytes in flight, not by DRAM)
        const long long s = s0 + tx;
        double v[OZ_ROWS / 8];
#pragma unroll
        for (int q = 0; q < OZ_ROWS / 8; q++) { const long long i = i0 + ty + 8 * q; v[q] = (i < n && s < B) ? __ldcs(&G[i * ldg + s]) : 0.0; }
#pragma unroll
        for (int q = 0; q < OZ_ROWS / 8; q++) {
            const double x = v[q];
            if (!(x == rint(x)) || fabs(x) > 127.0) { bad = 1; if (!isfinite(x)) bad = 3; } else gmax = max(gmax, (int)fabs(x));
            tile[oz_tile_slot(ty + 8 * q)][tx] = x;
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (tx == 0) { if (bad) atomicOr(&s_bad, bad); atomicMax(&s_max, gmax); }
    __syncthreads();
    for (int r = ty; r < OZ_TILE; r += 8) {
        const long long s = s0 + r, i = i0 + 4 * tx;
        if (s < Bp && i < Kp) {
            char4 q, q2;
            signed char* a = reinterpret_cast<signed char*>(&q);
            signed char* b = reinterpret_cast<signed char*>(&q2);
#pragma unroll
            for (int u = 0; u < 4; u++) { const int gv = (int)tile[(u << 5) | tx][r]; a[u] = (signed char)gv; b[u] = (signed char)(gv * gv); }
            *reinterpret_cast<char4*>(Gt8 + s * Kp + i) = q;
            if (G2t8) *reinterpret_cast<char4*>(G2t8 + s * Kp + i) = q2;
        }
    }
    if (threadIdx.x == 0) {      // one (conditional) atomic per block
        if (s_bad) atomicOr(&flags[0], 1);
        if (s_bad & 2) atomicOr(&flags[2], 1);
        if (s_max > *reinterpret_cast<volatile int*>(&flags[1])) atomicMax(&flags[1], s_max);
    }
}

// The same from int8 dosages stored row-major (cells x SNPs, row stride ld8 bytes): Gt8[s][i] = G8[i][s], G2t8[s][i] = G8[i][s]^2.
// flags[0] |= 1 when some entry is -128 (outside the symmetric range the split uses); flags[1] = max |g|.  128 x 128 tiles.
__global__ void __launch_bounds__(256) oz_transpose_i8_kernel(const int8_t* G8, long long ld8, long long n, long long B, int8_t* Gt8, int8_t* G2t8, long long Bp,
                                                              long long Kp, int* flags) {
    __shared__ int8_t tile[128][129];          // row pitch = 1 (mod 32) bytes: the transposed reads of a warp hit 32 different banks
    __shared__ int s_bad, s_max;
    const long long i0 = (long long)blockIdx.x * 128, s0 = (long long)blockIdx.y * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_bad = 0; s_max = 0; }
    __syncthreads();
    int bad = 0, gmax = 0;
    for (int r = warp; r < 128; r += 8) {
        const long long i = i0 + r;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long s = s0 + 4 * lane + u;
            int v = 0;
            if (i < n && s < B) v = G8[i * ld8 + s];
            if (v == -128) { bad = 1; v = 0; }
            gmax = max(gmax, abs(v));
            tile[r][4 * lane + u] = (int8_t)v;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0) { if (bad) atomicOr(&s_bad, 1); atomicMax(&s_max, gmax); }
    __syncthreads();
    for (int r = warp; r < 128; r += 8) {
        const long long s = s0 + r, i = i0 + 4 * lane;
        if (s < Bp && i < Kp) {
            char4 q, q2;
            signed char* a = reinterpret_cast<signed char*>(&q);
            signed char* b = reinterpret_cast<signed char*>(&q2);
#pragma unroll
            for (int u = 0; u < 4; u++) { const int gv = tile[4 * lane + u][r]; a[u] = (signed char)gv; b[u] = (signed char)(gv * gv); }
            *reinterpret_cast<char4*>(Gt8 + s * Kp + i) = q;
            if (G2t8) *reinterpret_cast<char4*>(G2t8 + s * Kp + i) = q2;
        }
    }
    if (threadIdx.x == 0 && flags) {
        if (s_bad) atomicOr(&flags[0], 1);
        if (s_max > *reinterpret_cast<volatile int*>(&flags[1])) atomicMax(&flags[1], s_max);
    }
}

// float64 image of a block of int8 dosages (row-major in, row-major out)
__global__ void oz_widen_i8_kernel(const int8_t* G8, long long ld8, long long n, long long B, double* out, long long ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * B) return;
    const long long i = idx / B, s = idx - i * B;
    out[i * ldo + s] = (double)G8[i * ld8 + s];
}
// flags[2] |= 1 when some entry of the block is not finite
__global__ void oz_finite_check_kernel(const double* G, long long ldg, long long n, long long B, int* flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (idx < n * B) { const long long i = idx / B, s = idx - i * B; bad = !isfinite(G[i * ldg + s]); }
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(&flags[2], 1);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Affine-integer genotype columns: g = a d + b with small non-negative integers d (standardised dosages -- the reference's own
// simulator hands column_normalize(G) to the scan, cellregmap/_simulate.py:50-54,339 -- centred dosages, dosages / 2 ...).  The
// contraction is linear in g, so it runs on d with the exact int8 split and is mapped back with the column sums of the basis:
//     sum_i X[i][col] g_i = a sum_i X[i][col] d_i + b sum_i X[i][col].
// Detection per column: the two smallest distinct values give b and a, every entry must then sit on the lattice b + a {0..127}
// within a few ulps of the column's magnitude (what the rounding of (d - mean) / sd leaves).
// ---------------------------------------------------------------------------------------------------------------------------
struct OzColStat { double m1, m2, mx; };      // smallest, second smallest distinct (+inf: none), largest
__device__ __forceinline__ void oz_stat_push(OzColStat& s, double v) {
    if (v < s.m1) { s.m2 = s.m1; s.m1 = v; }
    else if (v > s.m1 && v < s.m2) s.m2 = v;
    if (v > s.mx) s.mx = v;
}
__device__ __forceinline__ void oz_stat_merge(OzColStat& s, const OzColStat& o) {
    oz_stat_push(s, o.m1);
    if (o.m2 < s.m2 && o.m2 > s.m1) s.m2 = o.m2;
    if (o.mx > s.mx) s.mx = o.mx;
}
// partial[chunk][s]: statistics of the rows of one chunk; grid (column tiles of 32, row chunks), 32 x 8 threads
__global__ void __launch_bounds__(256) oz_colstat_kernel(const double* G, long long ldg, long long n, long long B, long long rows_per_chunk, OzColStat* partial) {
    __shared__ OzColStat sh[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long s = (long long)blockIdx.x * 32 + tx;
    const long long i0 = (long long)blockIdx.y * rows_per_chunk, i1 = min(n, i0 + rows_per_chunk);
    OzColStat st{INFINITY, INFINITY, -INFINITY};
    if (s < B) for (long long i = i0 + ty; i < i1; i += 8) oz_stat_push(st, G[i * ldg + s]);
    sh[ty][tx] = st;
    __syncthreads();
    if (ty == 0 && s < B) {
        for (int r = 1; r < 8; r++) oz_stat_merge(st, sh[r][tx]);
        partial[(long long)blockIdx.y * B + s] = st;
    }
}
// aff[0][s] = a (lattice step), aff[1][s] = b (origin), aff[2][s] = tolerance
__global__ void oz_colstat_merge_kernel(const OzColStat* partial, int chunks, long long B, double* aff, long long lda) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    OzColStat st = partial[s];
    for (int c = 1; c < chunks; c++) oz_stat_merge(st, partial[(long long)c * B + s]);
    const double step = (st.m2 < INFINITY) ? st.m2 - st.m1 : 1.0;          // constant column: d = 0 everywhere
    aff[s] = step;
    aff[lda + s] = st.m1;
    aff[2 * lda + s] = 8.0 * 2.220446049250313e-16 * fmax(fabs(st.m1), fabs(st.mx));
}
// Gt8[s][i] = d, G2t8[s][i] = d^2 with d = rint((G[i][s] - b_s) / a_s); flags[0] |= 1 when some entry is off the lattice (or not
// finite, or d > 127); flags[1] = max d.  Same tiling as oz_genotype_kernel.
__global__ void __launch_bounds__(256) oz_affine_genotype_kernel(const double* G, long long ldg, long long n, long long B, const double* aff, long long lda, int8_t* Gt8,
                                                                 int8_t* G2t8, long long Bp, long long Kp, int* flags) {
    __shared__ double tile[OZ_ROWS][OZ_TILE + 1];
    __shared__ int s_bad, s_max;
    // 1-D grid, SNP tiles fastest: blocks in flight together read neighbouring 256-byte pieces of the same genotype rows (DRAM pages)
    const long long s_tiles = (Bp + OZ_TILE - 1) / OZ_TILE;
    const long long i0 = ((long long)blockIdx.x / s_tiles) * OZ_ROWS;
    const long long s0 = ((long long)blockIdx.x % s_tiles) * OZ_TILE;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_bad = 0; s_max = 0; }
    __syncthreads();
    int bad = 0, gmax = 0;
    const long long sc = s0 + tx;
    const double a = sc < B ? aff[sc] : 1.0, b = sc < B ? aff[lda + sc] : 0.0, tol = sc < B ? aff[2 * lda + sc] : 0.0;
    for (int r = ty; r < OZ_ROWS; r += 8) {
        const long long i = i0 + r;
        double d = 0.0;
        if (i < n && sc < B) {
            const double v = G[i * ldg + sc];
            d = rint((v - b) / a);
            if (!(fabs(v - fma(a, d, b)) <= tol) || !(d >= 0.0 && d <= 127.0)) { bad = 1; d = 0.0; } else gmax = max(gmax, (int)d);
        }
        tile[oz_tile_slot(r)][tx] = d;
    }
    bad = __any_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (tx == 0) { if (bad) atomicOr(&s_bad, 1); atomicMax(&s_max, gmax); }
    __syncthreads();
    for (int r = ty; r < OZ_TILE; r += 8) {
        const long long s = s0 + r, i = i0 + 4 * tx;
        if (s < Bp && i < Kp) {
            char4 q, q2;
            signed char* qa = reinterpret_cast<signed char*>(&q);
            signed char* qb = reinterpret_cast<signed char*>(&q2);
#pragma unroll
            for (int u = 0; u < 4; u++) { const int gv = (int)tile[(u << 5) | tx][r]; qa[u] = (signed char)gv; qb[u] = (signed char)(gv * gv); }
            *reinterpret_cast<char4*>(Gt8 + s * Kp + i) = q;
            if (G2t8) *reinterpret_cast<char4*>(G2t8 + s * Kp + i) = q2;
        }
    }
    if (threadIdx.x == 0) {
        if (s_bad) atomicOr(&flags[0], 1);
        if (s_max > *reinterpret_cast<volatile int*>(&flags[1])) atomicMax(&flags[1], s_max);
    }
}
// C[s][col] <- a_s C[s][col] + b_s colsum[col]      (rotation of g = a d + b from the rotation of d)
__global__ void oz_affine_fix_kernel(double* C, long long ldc, long long B, long long cols, const double* aff, long long lda, const double* colsum) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    for (long long s = blockIdx.y; s < B; s += gridDim.y) {
        const double a = aff[s], b = aff[lda + s];
        C[s * ldc + col] = fma(a, C[s * ldc + col], b * colsum[col]);
    }
}
// sq[s][c] <- a^2 sq[s][c] + 2 a b lin[s][c] + b^2 colsum2[c]      (Grams of g^2 from those of d^2 and d)
__global__ void oz_affine_fix_square_kernel(double* sq, const double* lin, long long ld, long long B, int cols, const double* aff, long long lda, const double* colsum2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    for (long long s = blockIdx.y; s < B; s += gridDim.y) {
        const double a = aff[s], b = aff[lda + s];
        sq[s * ld + c] = a * a * sq[s * ld + c] + 2.0 * a * b * lin[s * ld + c] + b * b * colsum2[c];
    }
}

// C[s][col] = 2^(e_col + 1) sum_t 2^(-7 (t + 1)) D[t * Mp + col][s]   (D int32, row stride ldd; C fp64, row stride ldc)
__global__ void oz_combine_kernel(const int* D, long long Mp, long long ldd, const int* expo, long long Mtot, long long B, double* C, long long ldc) {
    __shared__ double tile[OZ_TILE][OZ_TILE + 1];
    const long long s0 = (long long)blockIdx.x * OZ_TILE;
    const long long c0 = (long long)blockIdx.y * OZ_TILE;
    for (int r = threadIdx.y; r < OZ_TILE; r += blockDim.y) {           // read: rows col, columns s (coalesced along s)
        const long long col = c0 + r, s = s0 + threadIdx.x;
        double acc = 0.0;
        if (col < Mtot && s < B) {
#pragma unroll
            for (int t = OZ_SLICES - 1; t >= 0; t--)                   // smallest terms first; Horner in base 2^-7 (exact scalings)
                acc = (acc + (double)D[((long long)t * Mp + col) * ldd + s]) * 0.0078125;
            const int e = expo[col];
            acc = (e == OZ_EXP_EMPTY) ? 0.0 : ldexp(acc, e + 1);
        }
        tile[r][threadIdx.x] = acc;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < OZ_TILE; r += blockDim.y) {           // write: rows s, columns col (coalesced along col)
        const long long s = s0 + r, col = c0 + threadIdx.x;
        if (s < B && col < Mtot) C[s * ldc + col] = tile[threadIdx.x][r];
    }
}

}  // namespace crm
