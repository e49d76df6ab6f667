// wroof.cu -- write-bandwidth microbenchmarks (B200): what can a store-only kernel reach, as a function of the
// store pattern? Used to put the Jacobian kernels' 434 MB-per-launch write stream in context.
//   A  cudaMemsetAsync
//   B  grid-stride kernel, every warp writes 256 contiguous bytes per instruction (st.global.cs.f64), address order
//   C  as B with 16-byte stores (st.global.cs.v2.f64): 512 B per warp instruction
//   D  "instance pattern": one CTA per 101 KB region; warp w of 8 writes 256-byte pieces that advance by a column
//      stride of 2288 B (the k_rows_n store pattern: thread k stores row k of column l, l = 0..39, 6 states)
//   E  as D, but each CTA stages 8 KB in shared memory and writes it with one TMA bulk store (cp.async.bulk)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/wroof scripts/wroof.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void kB(double* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, 1.0);
}
__global__ void kC(double2* p, size_t n2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        __stcs(p + i, make_double2(1.0, 2.0));
}
// region of NNZ doubles per CTA: 240 columns of 45 (+8 for two of six) entries; thread (i,k) writes entry k of column (l,i)
__global__ void __launch_bounds__(256) kD(double* p, int nnz) {
    double* base = p + (size_t)blockIdx.x * nnz;
    const int tid = threadIdx.x;
    if (tid >= 240) return;
    const int i = tid / 40, k = tid % 40;
    // column (l,i) starts at 720 + l*286 + i*45 + (i > 0 ? 8 : 0) + (i > 1 ? 8 : 0)
    const int coff = 720 + i * 45 + (i > 0 ? 8 : 0) + (i > 1 ? 8 : 0);
    for (int l = 0; l < 40; ++l) {
        const int e = coff + l * 286 + k + (l < k ? 5 : 0);
        __stcs(base + e, (double)l);
    }
}
__device__ __forceinline__ uint32_t s32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__global__ void __launch_bounds__(256) kE(double* p, int nnz) {
    extern __shared__ __align__(16) double sm[];
    double* base = p + (size_t)blockIdx.x * nnz;
    const int chunk = 1024;  // doubles
    const int nch = nnz / chunk;
    for (int c = 0; c < nch; ++c) {
        double* buf = sm + (c & 1) * chunk;
        for (int e = threadIdx.x; e < chunk; e += blockDim.x) buf[e] = (double)e;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)c * chunk), "r"(s32(buf)), "r"(chunk * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main() {
    const int B = 4096, NNZ = 12654;
    const size_t n = (size_t)B * NNZ;  // 51.8 M doubles = 415 MB
    double *p, *fl;
    CK(cudaMalloc(&p, n * 8 + 64));
    const size_t nf = 256u * 1024 * 1024 / 8;
    CK(cudaMalloc(&fl, nf * 8));
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    auto flush = [&]() { cudaMemset(fl, 1, nf * 8); kB<<<1184, 256>>>(fl, 0); cudaDeviceSynchronize(); };
    float ms;
    for (int variant = 0; variant < 7; ++variant) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 8; ++it) {
            flush();
            cudaEventRecord(s);
            switch (variant) {
                case 0: cudaMemsetAsync(p, 0, n * 8); break;
                case 1: kB<<<148 * 8, 256>>>(p, n); break;
                case 2: kC<<<148 * 8, 256>>>((double2*)p, n / 2); break;
                case 3: kD<<<B, 256>>>(p, NNZ); break;
                case 4: kE<<<B, 256, 16384>>>(p, NNZ); break;
                case 5: kB<<<148 * 32, 256>>>(p, n); break;
                case 6: kB<<<B, 256>>>(p, n); break;
            }
            cudaEventRecord(e);
            cudaEventSynchronize(e);
            cudaEventElapsedTime(&ms, s, e);
            if (it >= 2) { sum += ms; if (ms < best) best = ms; }
        }
        const char* names[] = {"A memset", "B st.cs.f64 grid-stride 1184 CTAs", "C st.cs.v2.f64 grid-stride", "D instance pattern (256-B pieces, column stride)",
                               "E instance pattern via 8 KB TMA bulk stores", "B' grid-stride 4736 CTAs", "B'' grid-stride 4096 CTAs"};
        const double bytes = variant == 3 ? (double)B * 240 * 40 * 8 : variant == 4 ? (double)B * (NNZ / 1024) * 8192 : (double)n * 8;
        printf("%-55s avg %.4f ms  best %.4f ms  %.0f GB/s (avg)\n", names[variant], sum / 6, best, bytes / (sum / 6 * 1e-3) / 1e9);
    }
    CK(cudaGetLastError());
    return 0;
}
