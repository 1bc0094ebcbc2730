// Probe: FP64 tensor-core (DMMA) mma.sync shapes on sm_100a: fragment-layout check + peak rate,
// and plain DFMA rate. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---------------- layout checks: A is MxK row-major, B is KxN (element (k,n) at B[k*N+n]) -------------
template <int SHAPE> __global__ void layout_kernel(const double* A, const double* B, double* D) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    if (SHAPE == 0) {  // m8n8k4
        double d[2] = {0, 0};
        mma884(d, A[g * 4 + t], B[t * 8 + g]);
        D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
    } else if (SHAPE == 1) {  // m16n8k4
        double d[4] = {0, 0, 0, 0}; double a[2] = {A[g * 4 + t], A[(g + 8) * 4 + t]};
        mma1684(d, a, B[t * 8 + g]);
        D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
    } else if (SHAPE == 2) {  // m16n8k8
        double d[4] = {0, 0, 0, 0};
        double a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
        double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
        mma1688(d, a, b);
        D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
    } else {  // m16n8k16
        double d[4] = {0, 0, 0, 0}; double a[8]; double b[4];
        for (int i = 0; i < 8; i++) a[i] = A[(g + (i & 1) * 8) * 16 + (i >> 1) * 4 + t];
        for (int i = 0; i < 4; i++) b[i] = B[(i * 4 + t) * 8 + g];
        mma16816(d, a, b);
        D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
    }
}

template <int SHAPE> void check_layout(const char* name, int M, int K) {
    const int N = 8;
    double hA[16 * 16], hB[16 * 8], hD[16 * 8], ref[16 * 8];
    for (int i = 0; i < M * K; i++) hA[i] = 1.0 + 0.37 * i + 0.001 * i * i;
    for (int i = 0; i < K * N; i++) hB[i] = -2.0 + 0.11 * i - 0.003 * i * i;
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double s = 0; for (int k = 0; k < K; k++) s += hA[m * K + k] * hB[k * N + n]; ref[m * N + n] = s; }
    double *dA, *dB, *dD; CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(hD)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, sizeof(hD)));
    layout_kernel<SHAPE><<<1, 32>>>(dA, dB, dD); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < M * N; i++) err = fmax(err, fabs(hD[i] - ref[i]) / (1 + fabs(ref[i])));
    printf("LAYOUT %-10s max_rel_err %.3e  %s\n", name, err, err < 1e-12 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// ---------------- throughput ----------------
template <int SHAPE, int NACC> __global__ void __launch_bounds__(256) rate_kernel(double* out, int iters, double seed) {
    int lane = threadIdx.x & 31;
    double acc[NACC][4];
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = seed + lane * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = seed - lane * 1e-3 - i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) { double d2[2] = {acc[i][0], acc[i][1]}; mma884(d2, a[i & 7], b[i & 3]); acc[i][0] = d2[0]; acc[i][1] = d2[1]; }
            else if (SHAPE == 1) { double a2[2] = {a[0], a[(i & 3) + 1]}; mma1684(acc[i], a2, b[i & 3]); }
            else if (SHAPE == 2) { double a4[4] = {a[0], a[1], a[2], a[(i & 3) + 3]}; double b2[2] = {b[0], b[(i & 1) + 1]}; mma1688(acc[i], a4, b2); }
            else { mma16816(acc[i], a, b); }
        }
    }
    double s = 0; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> __global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double acc[NACC]; for (int i = 0; i < NACC; i++) acc[i] = i;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], b, a);
    }
    double s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

template <int SHAPE, int NACC> void rate(const char* name, double flop_per_mma, int nsm, int blocks_per_sm, int threads, double* out) {
    int iters = 20000;
    int grid = nsm * blocks_per_sm;
    float ms = time_it([&] { rate_kernel<SHAPE, NACC><<<grid, threads>>>(out, iters, 1.0); });
    double flops = (double)grid * (threads / 32) * iters * NACC * flop_per_mma;
    printf("RATE %-10s nacc=%2d grid=%4d thr=%3d  %.3f ms  %.2f TFLOP/s\n", name, NACC, grid, threads, ms, flops / ms * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("GPU %s  SMs %d  cc %d.%d  clock %d MHz\n", p.name, p.multiProcessorCount, p.major, p.minor, clk / 1000);
    check_layout<0>("m8n8k4", 8, 4);
    check_layout<1>("m16n8k4", 16, 4);
    check_layout<2>("m16n8k8", 16, 8);
    check_layout<3>("m16n8k16", 16, 16);
    double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
    int nsm = p.multiProcessorCount;
    for (int thr = 128; thr <= 256; thr *= 2) for (int bps = 1; bps <= 2; bps++) {
        rate<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, nsm, bps, thr, out);
        rate<1, 8>("m16n8k4", 2.0 * 16 * 8 * 4, nsm, bps, thr, out);
        rate<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, nsm, bps, thr, out);
        rate<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, nsm, bps, thr, out);
    }
    rate<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, nsm, 1, 256, out);
    rate<3, 16>("m16n8k16", 2.0 * 16 * 8 * 16, nsm, 1, 256, out);
    rate<3, 2>("m16n8k16", 2.0 * 16 * 8 * 16, nsm, 1, 256, out);
    rate<3, 1>("m16n8k16", 2.0 * 16 * 8 * 16, nsm, 1, 128, out);
    rate<0, 1>("m8n8k4", 2.0 * 8 * 8 * 4, nsm, 1, 128, out);
    for (int thr = 256; thr <= 1024; thr *= 2) {
        int iters = 20000; int grid = nsm * 2;
        float ms = time_it([&] { dfma_kernel<16><<<grid, thr>>>(out, iters, 1.0); });
        printf("RATE DFMA thr=%4d grid=%d %.3f ms %.2f TFLOP/s\n", thr, grid, ms, 2.0 * grid * thr * iters * 16 / ms * 1e-9);
    }
    return 0;
}
