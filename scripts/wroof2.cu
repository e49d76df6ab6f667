// wroof2.cu -- does the STORE PATTERN of a kernel decide the latency its own loads see? (B200)
// Every CTA writes its 101 KB instance region (12 654 doubles) with one of two patterns while warp 7 of the CTA runs
// a dependent-load chain over an L2-resident table (pointer chase, 64 steps) and records its duration:
//   S  scattered: thread (i,k) stores entry k of column (l,i) for l = 0..39 (256-byte pieces, 8-byte granularity,
//      pieces start at arbitrary 8-byte offsets), then the node-local leftovers in short runs -- the k_rows_n pattern
//   B  bulk: the same bytes staged in shared memory and written with 8 KB cp.async.bulk stores (whole lines)
//   N  no stores (the chase alone)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/wroof2 scripts/wroof2.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(double* p, int nnz, const int* __restrict__ chase, int nchase, long long* lat, int* sink) {
    extern __shared__ __align__(16) double sm[];
    double* base = p + (size_t)blockIdx.x * nnz;
    const int tid = threadIdx.x;
    if (tid >= 224) {  // warp 7: latency probe, lane 0 only
        if (tid == 224) {
            int j = (blockIdx.x * 97) % nchase;
            long long t0 = clock64();
            for (int s = 0; s < 64; ++s) j = __ldcg(chase + j);
            long long t1 = clock64();
            lat[blockIdx.x] = (t1 - t0) / 64;
            if (j == -1) *sink = j;
        }
        if (MODE != 2) return;
    }
    if (MODE == 0) {
        if (tid < 224) {  // 224 "row" threads: 5.6 states x 40 nodes; stores spread over the region like k_rows_n's
            const int i = tid / 40, k = tid % 40;
            const int coff = 720 + i * 45 + (i > 0 ? 8 : 0) + (i > 1 ? 8 : 0);
            for (int l = 0; l < 40; ++l) {
                const int e = coff + l * 286 + k + (l < k ? 5 : 0);
                if (e < nnz) __stcs(base + e, (double)l);
            }
            // leftovers (node-local entries, control columns): short runs at other offsets
            for (int e = tid * 13; e < nnz; e += 224 * 13)
                for (int r = 0; r < 3; ++r) if (e + r * 4 < nnz) __stcs(base + e + r * 4, 1.0);
        }
    } else if (MODE == 1) {
        const int chunk = 1024, nch = nnz / chunk;
        for (int c = 0; c < nch; ++c) {
            double* buf = sm + (c % 3) * chunk;
            for (int e = tid; e < chunk; e += 224) buf[e] = (double)e;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 224;" ::: "memory");
            if (tid == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)c * chunk), "r"(s32(buf)), "r"(chunk * 8) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
            asm volatile("bar.sync 1, 224;" ::: "memory");
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

int main() {
    const int B = 4096, NNZ = 12654;
    const size_t n = (size_t)B * NNZ;
    double *p, *fl; int *chase, *sink; long long* lat;
    CK(cudaMalloc(&p, n * 8 + 64));
    const size_t nf = 256u * 1024 * 1024 / 8;
    CK(cudaMalloc(&fl, nf * 8));
    const int NC = 1 << 20;  // 4 MB table: L2 resident, far larger than L1
    std::vector<int> h(NC);
    for (int i = 0; i < NC; ++i) h[i] = (int)(((long long)i * 40503 + 12345) % NC);
    CK(cudaMalloc(&chase, NC * 4)); CK(cudaMemcpy(chase, h.data(), NC * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&lat, B * 8)); CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 8192));
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    std::vector<long long> hl(B);
    const char* names[] = {"S scattered 8-byte stores (k_rows_n pattern)", "B bulk 8 KB stores from shared memory", "N no stores"};
    for (int mode = 0; mode < 3; ++mode) {
        float sum = 0; double lsum = 0; long long lmax = 0;
        for (int it = 0; it < 6; ++it) {
            cudaMemset(fl, 1, nf * 8);                       // evict
            k<2><<<64, 256>>>(p, NNZ, chase, NC, lat, sink);  // warm the table into L2 (a few lines) -- and below
            for (int w = 0; w < 4; ++w) k<2><<<B, 256>>>(p, NNZ, chase, NC, lat, sink);
            cudaDeviceSynchronize();
            cudaEventRecord(s);
            if (mode == 0) k<0><<<B, 256>>>(p, NNZ, chase, NC, lat, sink);
            if (mode == 1) k<1><<<B, 256, 3 * 8192>>>(p, NNZ, chase, NC, lat, sink);
            if (mode == 2) k<2><<<B, 256>>>(p, NNZ, chase, NC, lat, sink);
            cudaEventRecord(e); cudaEventSynchronize(e);
            float ms; cudaEventElapsedTime(&ms, s, e);
            cudaMemcpy(hl.data(), lat, B * 8, cudaMemcpyDeviceToHost);
            if (it >= 2) { sum += ms; double a = 0; for (int b = 0; b < B; ++b) { a += hl[b]; if (hl[b] > lmax) lmax = hl[b]; } lsum += a / B; }
        }
        printf("%-48s kernel %.4f ms   L2 load latency seen by the same CTAs: mean %.0f cycles, max %lld\n", names[mode], sum / 4, lsum / 4, lmax);
    }
    CK(cudaGetLastError());
    return 0;
}
