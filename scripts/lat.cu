// lat.cu -- dependent-issue latencies on sm_100a that the FD kernel's schedule depends on:
// DFMA / DADD / DMUL chains, shared-memory and L1-hit global loads (pointer chase). One warp, clock64 around 4096 ops.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, const int* chase_g, int n) {
    __shared__ int chase_s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase_s[i] = (i + 32) & 1023;
    __syncthreads();
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = a + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = a * b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // two independent DFMA chains
    double a2 = a + 1.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a2 = fma(a2, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // four independent DFMA chains
    double a3 = a + 2.0, a4 = a + 3.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // shared-memory pointer chase
    int p = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) p = chase_s[p];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // global (L1-resident) pointer chase
    int q = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) q = __ldg(chase_g + q);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = a + a2 + a3 + a4 + p + q;
}
int main() {
    const int n = 4096;
    double* out; long long* cyc; int* cg;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8); cudaMalloc(&cg, 1024 * 4);
    int h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i + 32) & 1023;
    cudaMemcpy(cg, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, cg, n);
    long long c[8]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA dependent", "DADD dependent", "DMUL dependent", "2 independent DFMA chains (per pair)",
                           "4 independent DFMA chains (per quad)", "LDS pointer chase", "LDG (L1 hit) pointer chase"};
    for (int i = 0; i < 7; ++i) printf("%-42s %.2f cycles per step\n", names[i], (double)c[i] / n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
