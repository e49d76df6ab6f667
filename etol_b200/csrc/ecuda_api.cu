// ecuda_api.cu -- __global__ kernels for sm_100a and the extern "C" entry points of include/ecuda.h.
//
// Kernel layout: one CTA per (VGP instance, phase). The instance's obstacle/track records are
// fetched into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier, SASS: UBLKCP) while
// the threads un-scale the decision vector into shared memory; then the barrier-free phases of
// ecuda_phases.cuh run. Results are written with streaming stores straight into the caller's
// f / g / Jacobian-triplet arrays (IPOPT layout), so the only HBM traffic is the algorithmic one:
// 8*(nvars + 1 + ncons + nnz) bytes per instance plus its obstacle records.
//
// There is no CPU fallback in this file: without a usable sm_100 device every compute entry point
// returns ECUDA_ERR_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "ecuda_rows.cuh"

namespace ecuda {

constexpr int kThreads = 256;
#ifndef ECUDA_MIN_CTAS
#define ECUDA_MIN_CTAS 2 /* resident CTAs per SM the specialised kernels are register-limited for */
#endif

// ---- TMA bulk copy helpers (PTX) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (SASS: UBLKCP), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest bulk group committed by this thread have finished READING their shared source
__device__ __forceinline__ void bulk_wait_read_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all bulk groups committed by this thread are complete (their global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// barrier over the first `count` threads of the CTA only (count a multiple of 32)
__device__ __forceinline__ void named_barrier(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- kernels ------------------------------------------------------------------------------------------
// NB > 0: every phase has exactly NB summation blocks (block sums in registers); NB == 0: generic
template <int M, int NB>
__global__ void __launch_bounds__(kThreads, NB > 0 ? ECUDA_MIN_CTAS : 1) k_eval(const __grid_constant__ ProbDev pb,
                                                   const __grid_constant__ EvalIO io) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.x / pb.nphases;
    const int p = blockIdx.x - b * pb.nphases;
    const PhaseDev& ph = pb.ph[p];
    const int tid = threadIdx.x, nthr = blockDim.x;
    CtaMem m;
    carve(m, smem, pb, ph, nthr);
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = static_cast<uint32_t>(pb.inst_stride) * 8u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, bytes, &bar);
    }
    stage_vars(pb, ph, io, m, b, tid, nthr, io.jac != nullptr && io.jac_mode == ECUDA_JAC_FD_INDEXSET);
    mbar_wait(&bar, 0);
    __syncthreads();
    // blockIdx.y: slice of the phase (instances whose phases have more defect rows than threads)
    phase_b<M>(pb, ph, p, io, m, b, tid, nthr, blockIdx.y, gridDim.y);
    __syncthreads();
    phase_c<M, NB>(pb, ph, p, io, m, b, tid, nthr, blockIdx.y, gridDim.y);
}

// Specialised kernel (ecuda_fast.cuh): one defect row per thread, NB summation blocks, separate
// instantiations for finite differences and for the exact Jacobian (which needs neither the
// perturbation data in shared memory nor the registers of the FD rows).
#ifndef ECUDA_MIN_CTAS_FD
#define ECUDA_MIN_CTAS_FD 3
#endif
#ifndef ECUDA_MIN_CTAS_EXACT
#define ECUDA_MIN_CTAS_EXACT 4
#endif
// Exact mode: the D-coupled triplets are instance independent, so an extra warp streams them from the
// per-problem template (L2 resident) to the instance's triplet array with TMA bulk copies,
// global -> shared ring -> global, while the 256 compute threads run stage and phase B. The ring has
// kCopySlots buffers of kCopyChunk doubles; one lane drives it. Used when nnz is even, so that the
// template element e and its destination b*nnz + e always agree modulo 16 bytes.
constexpr int kCopyWarpThreads = 32;
#ifndef ECUDA_COPY_SLOTS
#define ECUDA_COPY_SLOTS 4
#endif
#ifndef ECUDA_COPY_CHUNK
#define ECUDA_COPY_CHUNK 1024
#endif
constexpr int kCopySlots = ECUDA_COPY_SLOTS;
constexpr int kCopyChunk = ECUDA_COPY_CHUNK;  // doubles (1024 = 8 KB)

__device__ void copy_warp_template(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, int b, double* ring,
                                   uint64_t* bars, int lane) {
    const int e0 = __ldg(pb.colptr + ph.zoff + pb.nc * ph.N);             // first state column of the phase
    const int e1 = __ldg(pb.colptr + ph.zoff + (pb.nc + pb.ns) * ph.N);   // its t0 column
    const size_t g0 = static_cast<size_t>(b) * pb.nnz + e0;               // global element index of the first
    const int n = e1 - e0;
    const double* src = pb.jtmpl + e0;  // nnz even: e0 and g0 have the same parity
    double* dst = io.jac + g0;
    const int head = static_cast<int>(g0 & 1);        // first element sits at an odd index: copy it alone
    const int nal = (n - head) & ~1;                  // doubles in the 16-byte aligned interior
    if (lane == 1 && head && n > 0) __stcs(dst, __ldg(src));
    if (lane == 2 && head + nal < n) __stcs(dst + n - 1, __ldg(src + n - 1));
    if (lane != 0) return;
    src += head;
    dst += head;
    const int nchunks = (nal + kCopyChunk - 1) / kCopyChunk;
    auto chunk_bytes = [&](int c) { return static_cast<uint32_t>(min(kCopyChunk, nal - c * kCopyChunk)) * 8u; };
    for (int c = 0; c < nchunks && c < kCopySlots; ++c) {
        mbar_expect_tx(&bars[c], chunk_bytes(c));
        bulk_g2s(ring + c * kCopyChunk, src + static_cast<size_t>(c) * kCopyChunk, chunk_bytes(c), &bars[c]);
    }
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c % kCopySlots;
        mbar_wait(&bars[slot], (c / kCopySlots) & 1);
        bulk_s2g(dst + static_cast<size_t>(c) * kCopyChunk, ring + slot * kCopyChunk, chunk_bytes(c));
        bulk_commit();
        // refill the slot of the PREVIOUS chunk: its store has had one iteration to read shared memory
        // (bulk groups complete in order, so "all but the newest" covers it)
        const int cn = c - 1 + kCopySlots;
        if (c >= 1 && cn < nchunks) {
            bulk_wait_read_but_one();
            const int ps = (c - 1) % kCopySlots;
            mbar_expect_tx(&bars[ps], chunk_bytes(cn));
            bulk_g2s(ring + ps * kCopyChunk, src + static_cast<size_t>(cn) * kCopyChunk, chunk_bytes(cn), &bars[ps]);
        }
    }
    bulk_wait_all();  // writes performed before the CTA-wide barrier that precedes the node-local stores
}

template <int M, int NB, bool FD>
__global__ void __launch_bounds__(FD ? kThreads : kThreads + kCopyWarpThreads,
                                  FD ? ECUDA_MIN_CTAS_FD : ECUDA_MIN_CTAS_EXACT)
    k_eval_fast(const __grid_constant__ ProbDev pb, const __grid_constant__ EvalIO io) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t copy_bars[kCopySlots];
    const int b = blockIdx.x / pb.nphases;
    const int p = blockIdx.x - b * pb.nphases;
    const PhaseDev& ph = pb.ph[p];
    const int tid = threadIdx.x, nthr = kThreads;  // compute threads; exact mode launches one more warp
    const bool copy_warp = !FD && blockDim.x > kThreads;  // uniform over the CTA
    CtaMem m;
    carve(m, smem, pb, ph, nthr, FD ? CARVE_FD : 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        if (copy_warp)
            for (int c = 0; c < kCopySlots; ++c) mbar_init(&copy_bars[c], 1);
    }
    __syncthreads();
    if (!FD && tid >= kThreads) {
        // ---- copy warp: template -> triplet array, then wait at the barrier before phase C
        double* ring = smem + cta_doubles(pb, ph, nthr, 0);
        ring += (reinterpret_cast<uintptr_t>(ring) & 8) ? 1 : 0;  // 16-byte aligned
        if (io.jac) copy_warp_template(pb, ph, io, b, ring, copy_bars, tid - kThreads);
        __syncthreads();
        return;
    }
    if (tid == 0) {
        const uint32_t bytes = static_cast<uint32_t>(pb.inst_stride) * 8u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, bytes, &bar);
    }
    if (!FD && io.jac && !copy_warp) fast_copy_template(pb, ph, io, b, tid, nthr);
    stage_vars(pb, ph, io, m, b, tid, nthr, FD && io.jac != nullptr);
    mbar_wait(&bar, 0);
    if (copy_warp) named_barrier(1, kThreads); else __syncthreads();
    RowRegs<NB> rr;
    fast_phase_b<M, NB, FD>(pb, ph, p, io, m, b, tid, nthr, rr);
    __syncthreads();  // all threads: in exact mode the template has landed before the node-local stores
    fast_phase_c<M, NB, FD>(pb, ph, p, io, m, b, tid, nthr, rr);
    if (io.nranks > 0) {
        // fused summary + all-gather epilogue: {f, max bound violation} of this instance goes straight
        // into every rank's gathered buffer over NVLink (P2P stores), row rank*batch + b
        __shared__ double red[kThreads / 32 + 1];
        double v = rr.viol;
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((tid & 31) == 0) red[tid >> 5] = v;
        if (tid == nthr - 1) red[kThreads / 32] = rr.fval;
        named_barrier(2, kThreads);  // the copy warp (exact mode) has already left
        if (tid < 32) {
            double w = tid < kThreads / 32 ? red[tid] : 0.0;
            for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
            if (tid < io.nranks) {
                double2* dst = reinterpret_cast<double2*>(io.peer[tid] + (static_cast<size_t>(io.rank) * io.batch + b) * 2);
                // fire and forget: the kernel boundary orders the store before the caller's cross-GPU barrier
                // (a system-scope fence here would keep the CTA resident for an NVLink round trip)
                *dst = make_double2(red[kThreads / 32], w);
            }
        }
    }
}

// Row-owner kernel (ecuda_rows.cuh): one barrier after staging, then every thread writes whole rows.
// Same launch shape as k_eval_fast (exact mode with a Jacobian: a 9th warp streams the template).
#ifndef ECUDA_MIN_CTAS_ROWS_FD
#define ECUDA_MIN_CTAS_ROWS_FD 3
#endif
template <int M, int NB, bool FD>
__global__ void __launch_bounds__(FD ? kThreads : kThreads + kCopyWarpThreads,
                                  FD ? ECUDA_MIN_CTAS_ROWS_FD : ECUDA_MIN_CTAS_EXACT)
    k_eval_rows(const __grid_constant__ ProbDev pb, const __grid_constant__ EvalIO io) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t copy_bars[kCopySlots];
    const int b = blockIdx.x / pb.nphases;
    const int p = blockIdx.x - b * pb.nphases;
    const PhaseDev& ph = pb.ph[p];
    const int tid = threadIdx.x, nthr = kThreads;
    const bool copy_warp = !FD && blockDim.x > kThreads;  // uniform over the CTA
    CtaMem m;
    carve(m, smem, pb, ph, nthr, FD ? CARVE_FD : 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        if (copy_warp)
            for (int c = 0; c < kCopySlots; ++c) mbar_init(&copy_bars[c], 1);
    }
    __syncthreads();
    // fused summary: this phase's block of the instance's bounds is staged behind the work arrays
    const int nbnd = io.nranks > 0 ? phase_ncons(pb, ph) + (phase_ncons(pb, ph) & 1) : 0;
    double* bnd = smem + cta_doubles(pb, ph, nthr, FD ? CARVE_FD : 0);
    if (!FD && tid >= kThreads) {  // copy warp: template -> triplet array, then the barrier before the triplets
        double* ring = bnd + 2 * nbnd;
        ring += (reinterpret_cast<uintptr_t>(ring) & 8) ? 1 : 0;
        if (io.jac) copy_warp_template(pb, ph, io, b, ring, copy_bars, tid - kThreads);
        __syncthreads();
        return;
    }
    if (tid == 0) {
        const uint32_t bytes = static_cast<uint32_t>(pb.inst_stride) * 8u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, bytes, &bar);
    }
    if (!FD && io.jac && !copy_warp) fast_copy_template(pb, ph, io, b, tid, nthr);
    if (io.nranks > 0) {  // coalesced, off the critical path: the values are needed after the barrier
        const size_t o = static_cast<size_t>(b) * pb.ncons + ph.goff;
        const int ncp = phase_ncons(pb, ph);
        for (int c0 = tid; c0 < ncp; c0 += 4 * nthr) {  // loads of four strides in flight together
            double lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u * nthr < ncp) {
                    lo[u] = __ldg(io.bl + o + c0 + u * nthr);
                    hi[u] = __ldg(io.bu + o + c0 + u * nthr);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u * nthr < ncp) {
                    bnd[c0 + u * nthr] = lo[u];
                    bnd[nbnd + c0 + u * nthr] = hi[u];
                }
        }
        m.bl = bnd;
        m.bu = bnd + nbnd;
    }
    stage_vars(pb, ph, io, m, b, tid, nthr, FD && io.jac != nullptr);
    mbar_wait(&bar, 0);
    if (copy_warp) named_barrier(1, kThreads); else __syncthreads();
    RowState<M, NB> rs;
    rows_values<M, NB, FD>(pb, ph, io, m, b, tid, rs);
    if (FD) {
        rows_jacobian<M, NB, FD>(pb, ph, io, m, b, tid, rs);
        rows_other<M, NB, FD>(pb, ph, p, io, m, b, tid, nthr, rs, true, true);
    } else {
        rows_other<M, NB, FD>(pb, ph, p, io, m, b, tid, nthr, rs, true, false);
        __syncthreads();  // all threads: the template has landed before the node-local triplets overwrite it
        rows_jacobian<M, NB, FD>(pb, ph, io, m, b, tid, rs);
        rows_other<M, NB, FD>(pb, ph, p, io, m, b, tid, nthr, rs, false, true);
    }
    if (io.nranks > 0) {  // fused summary + all-gather epilogue (see k_eval_fast)
        __shared__ double red[kThreads / 32 + 1];
        double v = rs.viol;
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((tid & 31) == 0) red[tid >> 5] = v;
        if (tid == nthr - 1) red[kThreads / 32] = rs.fval;
        named_barrier(2, kThreads);
        if (tid < 32) {
            double w = tid < kThreads / 32 ? red[tid] : 0.0;
            for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
            if (tid < io.nranks) {
                double2* dst = reinterpret_cast<double2*>(io.peer[tid] + (static_cast<size_t>(io.rank) * io.batch + b) * 2);
                *dst = make_double2(red[kThreads / 32], w);
            }
        }
    }
}

// Write n doubles from the shared image `src` (element i at src[par + i], par = parity of the global
// element index of the first one, so that shared and global addresses are 16-byte aligned together)
// to dst[0..n): the aligned interior by one bulk copy issued by thread `lead`, the at most two
// boundary elements by plain stores from the next two threads.
__device__ __forceinline__ void flush_range(double* dst, const double* src, int par, int n, int tid, int lead) {
    const int start = par;             // par == 1: element 0 sits at an odd global index
    const int nal = (n - start) & ~1;  // doubles in the 16-byte aligned interior
    if (tid == lead) {
        if (nal > 0) bulk_s2g(dst + start, src + par + start, static_cast<uint32_t>(nal) * 8u);
        bulk_commit();
    } else if (tid == lead + 1) {
        if (start == 1 && n > 0) __stcs(dst, src[par]);
    } else if (tid == lead + 2) {
        if (start + nal < n) __stcs(dst + n - 1, src[par + n - 1]);
    }
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Exact mode, persistent: a CTA keeps a shared-memory IMAGE of its phase's whole triplet range. The
// instance-independent D-coupled triplets are loaded into it once from the per-problem template;
// for every instance the CTA works on, phase C overwrites the node-local triplets in the image (every
// one of them, the pattern is the same for all instances) and one bulk shared->global copy (TMA)
// writes the range out while the CTA already stages and evaluates its next instance. No triplet is
// stored by an LSU instruction, nothing is read back, and the template is read once per CTA instead
// of once per instance. Two CTAs per SM (the image of the benchmark shape is 101 KB).
template <int M, int NB>
__global__ void __launch_bounds__(kThreads, 2)
    k_eval_image(const __grid_constant__ ProbDev pb, const __grid_constant__ EvalIO io) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    const int p = blockIdx.x % pb.nphases;  // fixed per CTA: gridDim.x is a multiple of nphases
    const PhaseDev& ph = pb.ph[p];
    const int tid = threadIdx.x, nthr = blockDim.x;
    CtaMem m;
    carve(m, smem, pb, ph, nthr, 0);
    double* image = smem + cta_doubles(pb, ph, nthr, 0);
    image += (reinterpret_cast<uintptr_t>(image) & 8) ? 1 : 0;  // 16-byte aligned
    const int c0 = __ldg(pb.colptr + ph.zoff), c1 = __ldg(pb.colptr + ph.zoff + ph.nvars);
    const int n = c1 - c0;
    const int par = c0 & 1;  // nnz is even (checked on the host): triplet c0 of every instance has this parity
    double* vimage = image + par - c0;  // vimage[e] = slot of triplet e
    for (int e = c0 + tid; e < c1; e += nthr) vimage[e] = __ldg(pb.jtmpl + e);
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const int stride = gridDim.x / pb.nphases;
    for (int b = blockIdx.x / pb.nphases; b < io.batch; b += stride) {
        if (tid == 0) {
            const uint32_t bytes = static_cast<uint32_t>(pb.inst_stride) * 8u;
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(m.inst, io.inst + static_cast<size_t>(b) * pb.inst_stride, bytes, &bar);
        }
        stage_vars(pb, ph, io, m, b, tid, nthr, false);
        mbar_wait(&bar, parity);
        parity ^= 1;
        __syncthreads();
        RowRegs<NB> rr;
        fast_phase_b<M, NB, false>(pb, ph, p, io, m, b, tid, nthr, rr);
        if (tid == 0) bulk_wait_read_all();  // the previous instance's image has left shared memory
        __syncthreads();
        fast_phase_c<M, NB, false, true>(pb, ph, p, io, m, b, tid, nthr, rr, vimage);
        fence_async_smem();  // generic-proxy writes to the image -> visible to the bulk copy
        __syncthreads();
        flush_range(io.jac + static_cast<size_t>(b) * pb.nnz + c0, image, par, n, tid, 0);
    }
    if (tid == 0) bulk_wait_read_all();
}

template <int M>
__global__ void __launch_bounds__(kThreads) k_grad(const __grid_constant__ ProbDev pb,
                                                   const __grid_constant__ EvalIO io) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / pb.nphases;
    const int p = blockIdx.x - b * pb.nphases;
    const PhaseDev& ph = pb.ph[p];
    const int tid = threadIdx.x, nthr = blockDim.x;
    CtaMem m;
    carve(m, smem, pb, ph, nthr);
    stage_vars(pb, ph, io, m, b, tid, nthr, false);
    __syncthreads();
    cost_nodes<M>(pb, ph, m, tid, nthr);
    __syncthreads();
    gradient_phase<M>(pb, ph, io, m, b, tid, nthr);
}

// multi-phase objective: f = sf * (((f_0 + f_1) + f_2) + ...)
__global__ void k_sum_phases(const double* fpart, double* f, int batch, int nphases, double sf) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double tot = fpart[static_cast<size_t>(b) * nphases];
    for (int p = 1; p < nphases; ++p) tot = tot + fpart[static_cast<size_t>(b) * nphases + p];
    f[b] = sf * tot;
}

// per-instance summary {f, max bound violation}; one warp per instance, shuffle max-reduce
__global__ void k_summary(const double* f, const double* g, const double* gl, const double* gu, double* out,
                          int batch, int ncons) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= batch) return;
    const double* gb = g + static_cast<size_t>(warp) * ncons;
    const double* lb = gl + static_cast<size_t>(warp) * ncons;
    const double* ub = gu + static_cast<size_t>(warp) * ncons;
    double v = 0.0;
    for (int r = lane; r < ncons; r += 32) {
        double gv = gb[r];
        v = fmax(v, fmax(lb[r] - gv, gv - ub[r]));
    }
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) {
        out[2 * warp] = f[warp];
        out[2 * warp + 1] = v;
    }
}

// Fused per-instance summary + all-gather over NVLink peer memory: one warp reduces instance i to
// {f, max bound violation} (shuffle max) and lane r stores the 16-byte row straight into rank r's
// gathered buffer at row rank*batch + i (P2P stores; the buffers come from a symmetric-memory
// rendezvous). No staging buffer, no separate collective launch: the exchange of the sharded path
//  is the epilogue of the reduction. A cross-GPU barrier after the kernel
// (caller's stream) makes the rows visible everywhere.
struct PeerPtrs {
    double* p[16];
};
__global__ void k_summary_scatter(const double* f, const double* g, const double* gl, const double* gu, PeerPtrs peers,
                                  int nranks, int rank, int batch, int ncons) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= batch) return;
    const double* gb = g + static_cast<size_t>(warp) * ncons;
    const double* lb = gl + static_cast<size_t>(warp) * ncons;
    const double* ub = gu + static_cast<size_t>(warp) * ncons;
    double v = 0.0;
    for (int r = lane; r < ncons; r += 32) {
        double gv = gb[r];
        v = fmax(v, fmax(lb[r] - gv, gv - ub[r]));
    }
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const double fv = f[warp];
    if (lane < nranks) {
        double2* dst = reinterpret_cast<double2*>(peers.p[lane] + (static_cast<size_t>(rank) * batch + warp) * 2);
        *dst = make_double2(fv, v);  // ordered before the caller's cross-GPU barrier by the kernel boundary
    }
}

// Cross-GPU barrier for the fused exchange: one block, thread r talks to rank r. Each rank owns an array
// of counters in peer-visible memory; "I have finished step s" is a store of s into my slot of every
// peer's array, after a system-scope fence that orders this GPU's earlier peer stores (the gathered
// rows of the kernel before) ahead of it; then every thread waits until its peer's slot in the local
// array has reached s. Counters only grow, so nothing is ever reset. The wait is bounded (about 0.1 s
// of SM clocks): a lost peer turns into a wrong result the caller can detect, not into a hung GPU.
struct PeerFlags {
    unsigned long long* p[16];
};
__global__ void k_peer_barrier(PeerFlags flags, int nranks, int rank, unsigned long long step, int* timed_out) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    __threadfence_system();
    volatile unsigned long long* theirs = flags.p[r] + rank;  // my slot in rank r's array
    *theirs = step;
    volatile unsigned long long* mine = flags.p[rank] + r;    // rank r's slot in my array
    const long long t0 = clock64();
    while (*mine < step) {
        if (clock64() - t0 > 200000000ll) {
            if (timed_out) *timed_out = 1;
            break;
        }
    }
    __threadfence_system();
}

// FP64 FMA microbenchmark: 8 independent dependent-chains per thread, no memory traffic. Defines the
// FP64 roof the finite-difference kernel is compared with (MEASURED_PEAKS.json has no FP64 figure).
__global__ void __launch_bounds__(256) k_fp64_peak(double* sink, int iters, double a, double b) {
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    const double r = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (r == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = r;  // never true: keeps the chains alive
}

// ---- context ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace ecuda

using namespace ecuda;

struct ecuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool have_problem = false, have_inst = false, have_bounds = false;
    HostProblem hp;
    ProbDev pd{};
    DevBuf colptr, isz, sg, inst, gl, gu, fpart, jtmpl, bflag;
    DevBuf coll[ECUDA_MAX_PHASES];  // D | Dt | tau | w per phase
    DevBuf sx, sf_, sgv, sjac, sgrad, ssum;  // staging for host-memory calls
    size_t smem_bytes = 0, smem_fast_fd = 0, smem_fast_exact = 0;
    bool fast_ok = false;  // the specialised kernels (ecuda_fast.cuh) can run this problem
    bool no_fast = false;  // ECUDA_NO_FAST=1 in the environment: never use them (A/B runs and tests)
    bool no_image = true;   // ECUDA_IMAGE=1 opts in to the persistent image kernel (exact mode); measured
                            // slower than the copy-warp kernel on C2 (0.180 vs 0.157 ms: 2 CTAs/SM leave the
                            // latency-bound node-local phases exposed), kept for the next round's pipelining work
    bool image_ok = false;
    size_t smem_image = 0;
    int num_sms = 148;
    bool no_rows = false;   // ECUDA_NO_ROWS=1: use the column-owner kernels (k_eval_fast) instead of k_eval_rows
    bool no_copy_warp = false;  // ECUDA_NO_COPY_WARP=1: exact mode copies the template with plain loads/stores
    int64_t launches = 0;
    int ipopt_jac_mode = ECUDA_JAC_EXACT;
    bool force_generic = false;  // ECUDA_FORCE_GENERIC=1 in the environment: always run the generic kernel
    int nb_uniform = 0;  // summation-block count if all phases share it and it is <= 8, else 0
    std::vector<double> h_sz, h_sg;
};

static std::string g_create_err;

static int fail(ecuda_ctx* h, int code, const std::string& msg) {
    if (h)
        h->err = msg;
    else
        g_create_err = msg;
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(h, ECUDA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

static int ensure(ecuda_ctx* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return ECUDA_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    if (bytes == 0) bytes = 8;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(h, ECUDA_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    b.bytes = bytes;
    return ECUDA_OK;
}
static void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

// exact-mode template of the instance-independent triplets; depends on the scaling and on D
static int upload_template(ecuda_ctx* h) {
    if (h->hp.col.empty()) return ECUDA_OK;  // collocation data not built yet
    std::vector<double> isz(h->hp.dims.nvars), tmpl;
    for (int c = 0; c < h->hp.dims.nvars; ++c) isz[c] = 1.0 / h->h_sz[c];
    build_jac_template(h->hp, isz.data(), h->h_sg.data(), &tmpl);
    tmpl.resize(tmpl.size() + 2, 0.0);  // bulk copies read whole 16-byte units
    int rc;
    if ((rc = ensure(h, h->jtmpl, sizeof(double) * tmpl.size()))) return rc;
    CU(cudaMemcpy(h->jtmpl.p, tmpl.data(), sizeof(double) * tmpl.size(), cudaMemcpyHostToDevice));
    h->pd.jtmpl = static_cast<const double*>(h->jtmpl.p);
    return ECUDA_OK;
}

static int upload_scaling(ecuda_ctx* h) {
    const int nv = h->hp.dims.nvars, ng = h->hp.dims.ncons;
    std::vector<double> isz(nv);
    for (int c = 0; c < nv; ++c) isz[c] = 1.0 / h->h_sz[c];  // reciprocal rounded once, on the host
    int rc;
    if ((rc = ensure(h, h->isz, sizeof(double) * nv))) return rc;
    if ((rc = ensure(h, h->sg, sizeof(double) * ng))) return rc;
    CU(cudaMemcpy(h->isz.p, isz.data(), sizeof(double) * nv, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->sg.p, h->h_sg.data(), sizeof(double) * ng, cudaMemcpyHostToDevice));
    h->pd.isz = static_cast<const double*>(h->isz.p);
    h->pd.sg = static_cast<const double*>(h->sg.p);
    return upload_template(h);
}

static int upload_collocation(ecuda_ctx* h, int p) {
    const Collocation& c = h->hp.col[p];
    const size_t N = c.N;
    std::vector<double> pack(2 * N * N + 2 * N);
    for (size_t k = 0; k < N; ++k)
        for (size_t l = 0; l < N; ++l) {
            pack[k * N + l] = c.D[k * N + l];
            pack[N * N + l * N + k] = c.D[k * N + l];
        }
    std::memcpy(&pack[2 * N * N], c.tau.data(), sizeof(double) * N);
    std::memcpy(&pack[2 * N * N + N], c.w.data(), sizeof(double) * N);
    int rc;
    if ((rc = ensure(h, h->coll[p], sizeof(double) * pack.size()))) return rc;
    CU(cudaMemcpy(h->coll[p].p, pack.data(), sizeof(double) * pack.size(), cudaMemcpyHostToDevice));
    const double* base = static_cast<const double*>(h->coll[p].p);
    h->pd.ph[p].D = base;
    h->pd.ph[p].Dt = base + N * N;
    h->pd.ph[p].tau = base + 2 * N * N;
    h->pd.ph[p].w = base + 2 * N * N + N;
    return ECUDA_OK;
}

template <int M, int NB>
static int launch_keval(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    // opt in to more than 48 KB of dynamic shared memory. The attribute is per function and device
    // and only permits, so it is raised monotonically to the largest size any handle has needed.
    static std::mutex mu;
    static size_t configured[64] = {0};
    if (h->smem_bytes > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < h->smem_bytes) {
            CU(cudaFuncSetAttribute(k_eval<M, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
            cur = h->smem_bytes;
        }
    }
    int nslices = 1;  // the widest phase decides; extra slices of a narrower phase just own fewer nodes
    for (int p = 0; p < h->pd.nphases; ++p) nslices = std::max(nslices, generic_slices(h->pd.ns, h->pd.ph[p].N, kThreads));
    k_eval<M, NB><<<dim3(grid, nslices), kThreads, h->smem_bytes, st>>>(h->pd, io);
    return ECUDA_OK;
}

template <int M, int NB, bool FD>
static int launch_keval_fast(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    // exact mode with a Jacobian: one more warp and its shared-memory ring (see copy_warp_template)
    const bool copy_warp = !FD && io.jac != nullptr && !h->no_copy_warp && (h->pd.nnz & 1) == 0 &&
                           (reinterpret_cast<uintptr_t>(io.jac) & 15) == 0;
    const size_t smem = FD ? h->smem_fast_fd
                           : h->smem_fast_exact + (copy_warp ? (kCopySlots * kCopyChunk + 2) * sizeof(double) : 0);
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < smem) {
            CU(cudaFuncSetAttribute(k_eval_fast<M, NB, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    if (!h->no_rows) {
        static size_t configured_rows[64] = {0};
        size_t smem_rows = smem;
        if (io.nranks > 0) {  // staged bounds (fused summary), largest phase
            int ncp = 0;
            for (int p = 0; p < h->pd.nphases; ++p) ncp = std::max(ncp, phase_ncons(h->pd, h->pd.ph[p]));
            smem_rows += 2 * static_cast<size_t>(ncp + 2) * sizeof(double);
        }
        if (smem_rows > 48 * 1024) {
            std::lock_guard<std::mutex> lock(mu);
            size_t& cur = configured_rows[h->device & 63];
            if (cur < smem_rows) {
                CU(cudaFuncSetAttribute(k_eval_rows<M, NB, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
                cur = smem_rows;
            }
        }
        k_eval_rows<M, NB, FD><<<grid, kThreads + (copy_warp ? kCopyWarpThreads : 0), smem_rows, st>>>(h->pd, io);
        return ECUDA_OK;
    }
    k_eval_fast<M, NB, FD><<<grid, kThreads + (copy_warp ? kCopyWarpThreads : 0), smem, st>>>(h->pd, io);
    return ECUDA_OK;
}
template <int M, int NB>
static int launch_keval_image(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    const size_t smem = h->smem_image;
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < smem) {
            CU(cudaFuncSetAttribute(k_eval_image<M, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    // persistent: as many CTAs as fit (2 per SM while the image allows it), a multiple of nphases
    const int per_sm = smem <= 112 * 1024 ? 2 : 1;
    int grid = h->num_sms * per_sm;
    grid = std::min(grid, io.batch * h->pd.nphases);
    grid -= grid % h->pd.nphases;
    if (grid < h->pd.nphases) grid = h->pd.nphases;
    k_eval_image<M, NB><<<grid, kThreads, smem, st>>>(h->pd, io);
    return ECUDA_OK;
}

template <int M, int NB>
static int launch_keval_fast_mode(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    if (io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET) return launch_keval_fast<M, NB, true>(h, io, st, grid);
    if (io.jac && h->image_ok && (reinterpret_cast<uintptr_t>(io.jac) & 15) == 0) return launch_keval_image<M, NB>(h, io, st);
    return launch_keval_fast<M, NB, false>(h, io, st, grid);
}

template <int M>
static int launch_eval_t(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    const int grid = io.batch * h->pd.nphases;
    if (io.grad) {
        static std::mutex mu;
        static size_t configured[64] = {0};
        if (h->smem_bytes > 48 * 1024) {
            std::lock_guard<std::mutex> lock(mu);
            size_t& cur = configured[h->device & 63];
            if (cur < h->smem_bytes) {
                CU(cudaFuncSetAttribute(k_grad<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
                cur = h->smem_bytes;
            }
        }
        k_grad<M><<<grid, kThreads, h->smem_bytes, st>>>(h->pd, io);
        ++h->launches;
    }
    if (io.f || io.g || io.jac) {
        int rc = ECUDA_OK;
        if (h->fast_ok) {
            switch (h->nb_uniform) {
                case 3: rc = launch_keval_fast_mode<M, 3>(h, io, st, grid); break;
                case 4: rc = launch_keval_fast_mode<M, 4>(h, io, st, grid); break;
                default: rc = launch_keval_fast_mode<M, 5>(h, io, st, grid); break;
            }
        } else
        switch (h->nb_uniform) {  // block count shared by all phases, or 0
            // specialised for the node counts of the BASELINE configs: 17 -> 3, 30 -> 4, 33 / 40 -> 5
            case 3: rc = launch_keval<M, 3>(h, io, st, grid); break;
            case 4: rc = launch_keval<M, 4>(h, io, st, grid); break;
            case 5: rc = launch_keval<M, 5>(h, io, st, grid); break;
            default: rc = launch_keval<M, 0>(h, io, st, grid); break;
        }
        if (rc) return rc;
        ++h->launches;
        if (io.f && h->pd.nphases > 1) {
            k_sum_phases<<<(io.batch + 127) / 128, 128, 0, st>>>(io.fpart, io.f, io.batch, h->pd.nphases, h->pd.sf);
            ++h->launches;
        }
    }
    CU(cudaGetLastError());
    return ECUDA_OK;
}

static int launch_eval(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    switch (h->pd.model) {
        case ECUDA_MODEL_SI2D: return launch_eval_t<ECUDA_MODEL_SI2D>(h, io, st);
        case ECUDA_MODEL_PM3D: return launch_eval_t<ECUDA_MODEL_PM3D>(h, io, st);
        case ECUDA_MODEL_FW6: return launch_eval_t<ECUDA_MODEL_FW6>(h, io, st);
    }
    return fail(h, ECUDA_ERR_ARG, "unknown model");
}

// ---- host evaluation of the device models (callback verification in the plugin; no GPU) ------------------
template <int M>
static void host_model_eval_t(const double* x, const double* u, double t, double* f_out, double* cost_out) {
    Model<M>::f(x, u, t, f_out);
    *cost_out = Model<M>::cost(x, u, t);
}
extern "C" {

const char* ecuda_last_error(ecuda_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int ecuda_create(int device, ecuda_handle* out) {
    ecuda_ctx* h = nullptr;
    if (!out) return fail(nullptr, ECUDA_ERR_ARG, "null output handle");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, ECUDA_ERR_CUDA,
                    std::string("eCUDA needs an NVIDIA sm_100 GPU and has no CPU fallback: ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device present"));
    if (device < 0 || device >= n) return fail(nullptr, ECUDA_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, ECUDA_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, ECUDA_ERR_CUDA,
                    "eCUDA kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                        std::to_string(prop.minor));
    h = new (std::nothrow) ecuda_ctx;
    if (!h) return fail(nullptr, ECUDA_ERR_ALLOC, "out of host memory");
    h->device = device;
    {
        const char* fg = std::getenv("ECUDA_FORCE_GENERIC");
        h->force_generic = fg && fg[0] == '1';
        const char* nf = std::getenv("ECUDA_NO_FAST");
        h->no_fast = nf && nf[0] == '1';
        const char* ni = std::getenv("ECUDA_IMAGE");
        h->no_image = !(ni && ni[0] == '1');
        h->num_sms = prop.multiProcessorCount;
        const char* nr = std::getenv("ECUDA_NO_ROWS");
        h->no_rows = nr && nr[0] == '1';
        const char* nw = std::getenv("ECUDA_NO_COPY_WARP");
        h->no_copy_warp = nw && nw[0] == '1';
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        delete h;
        return fail(nullptr, ECUDA_ERR_CUDA, msg);
    }
    *out = h;
    return ECUDA_OK;
}

int ecuda_destroy(ecuda_handle h) {
    if (!h) return ECUDA_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (DevBuf* b : {&h->colptr, &h->isz, &h->sg, &h->inst, &h->gl, &h->gu, &h->fpart, &h->jtmpl, &h->bflag, &h->sx, &h->sf_, &h->sgv,
                      &h->sjac, &h->sgrad, &h->ssum})
        release(*b);
    for (auto& b : h->coll) release(b);
    cudaStreamDestroy(h->stream);
    delete h;
    return ECUDA_OK;
}

int ecuda_set_problem(ecuda_handle h, const ecuda_problem_desc* desc) {
    if (!h) return ECUDA_ERR_ARG;
    if (!desc) return fail(h, ECUDA_ERR_ARG, "null problem description");
    CU(cudaSetDevice(h->device));
    std::string err;
    HostProblem hp;
    if (!build_layout(*desc, &hp, &err)) return fail(h, ECUDA_ERR_ARG, err);
    build_structure(&hp);
    hp.col.resize(hp.nphases);
    for (int p = 0; p < hp.nphases; ++p)
        if (!build_collocation(desc->collocation, hp.N[p], &hp.col[p], &err)) return fail(h, ECUDA_ERR_ARG, err);
    h->hp = hp;
    h->have_problem = false;
    h->have_inst = false;
    h->have_bounds = false;
    ProbDev& pd = h->pd;
    fill_probdev(hp, &pd);
    size_t smem = 0, smem_fd = 0, smem_ex = 0;
    bool one_row_per_thread = true;
    for (int p = 0; p < hp.nphases; ++p) {
        PhaseDev& ph = pd.ph[p];
        int rc = upload_collocation(h, p);
        if (rc) return rc;
        smem = std::max(smem, cta_doubles(pd, ph, kThreads) * sizeof(double));
        smem_fd = std::max(smem_fd, cta_doubles(pd, ph, kThreads, CARVE_FD) * sizeof(double));
        smem_ex = std::max(smem_ex, cta_doubles(pd, ph, kThreads, 0) * sizeof(double));
        one_row_per_thread = one_row_per_thread && pd.ns * ph.N <= kThreads &&
                             (2 * ph.npath + pd.nc + 2) * ph.N < 65536;  // fast_div range
    }
    if (smem > 227 * 1024 - 64)
        return fail(h, ECUDA_ERR_ARG, "phase too large for one CTA's shared memory (" + std::to_string(smem) + " B)");
    h->smem_bytes = smem;
    h->smem_fast_fd = smem_fd;
    h->smem_fast_exact = smem_ex;
    // persistent image kernel: work area + the largest phase slice of the triplet array (+ alignment pad)
    size_t smem_img = 0;
    for (int p = 0; p < hp.nphases; ++p) {
        const size_t slice = hp.colptr[hp.zoff[p] + hp.nvars_p[p]] - hp.colptr[hp.zoff[p]];
        smem_img = std::max(smem_img, (cta_doubles(pd, pd.ph[p], kThreads, 0) + slice + 4) * sizeof(double));
    }
    h->smem_image = smem_img;
    h->nb_uniform = pd.ph[0].nb;
    for (int p = 1; p < hp.nphases; ++p)
        if (pd.ph[p].nb != h->nb_uniform) h->nb_uniform = 0;
    if (h->nb_uniform > 8) h->nb_uniform = 0;
    if (h->force_generic) h->nb_uniform = 0;
    // the specialised kernels: block counts of the BASELINE configs, one defect row per thread
    h->fast_ok = !h->no_fast && !h->force_generic && h->nb_uniform >= 3 && h->nb_uniform <= 5 && one_row_per_thread;
    h->image_ok = h->fast_ok && !h->no_image && (pd.nnz & 1) == 0 && smem_img <= 227 * 1024 - 1024;
    int rc;
    if ((rc = ensure(h, h->colptr, sizeof(int32_t) * (pd.nvars + 1)))) return rc;
    CU(cudaMemcpy(h->colptr.p, hp.colptr.data(), sizeof(int32_t) * (pd.nvars + 1), cudaMemcpyHostToDevice));
    pd.colptr = static_cast<const int*>(h->colptr.p);
    h->h_sz.assign(pd.nvars, 1.0);
    h->h_sg.assign(pd.ncons, 1.0);
    if ((rc = upload_scaling(h))) return rc;
    if ((rc = ensure(h, h->fpart, sizeof(double) * desc->batch * hp.nphases))) return rc;
    if ((rc = ensure(h, h->inst, sizeof(double) * desc->batch * pd.inst_stride))) return rc;
    CU(cudaMemset(h->inst.p, 0, sizeof(double) * desc->batch * pd.inst_stride));
    h->have_problem = true;
    // a problem with no obstacle data needs no upload
    bool any = desc->ntracks > 0;
    for (int p = 0; p < hp.nphases; ++p) any = any || hp.nstat[p] > 0;
    h->have_inst = !any;
    return ECUDA_OK;
}

int ecuda_get_dims(ecuda_handle h, ecuda_dims* out) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!out) return fail(h, ECUDA_ERR_ARG, "null output");
    *out = h->hp.dims;
    return ECUDA_OK;
}

int ecuda_get_structure(ecuda_handle h, int32_t* iRow, int32_t* jCol, int32_t* group_of_col) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    const int base = h->hp.desc.index_base;
    for (int e = 0; e < h->hp.dims.nnz; ++e) {
        if (iRow) iRow[e] = h->hp.irow[e] + base;
        if (jCol) jCol[e] = h->hp.jcol[e] + base;
    }
    if (group_of_col) std::memcpy(group_of_col, h->hp.group_of_col.data(), sizeof(int32_t) * h->hp.dims.nvars);
    return ECUDA_OK;
}

int ecuda_get_collocation(ecuda_handle h, int phase, double* tau, double* w, double* D) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (phase < 0 || phase >= h->hp.nphases) return fail(h, ECUDA_ERR_ARG, "phase out of range");
    const Collocation& c = h->hp.col[phase];
    if (tau) std::memcpy(tau, c.tau.data(), sizeof(double) * c.N);
    if (w) std::memcpy(w, c.w.data(), sizeof(double) * c.N);
    if (D) std::memcpy(D, c.D.data(), sizeof(double) * c.N * c.N);
    return ECUDA_OK;
}

int ecuda_set_collocation(ecuda_handle h, int phase, const double* tau, const double* w, const double* D) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (phase < 0 || phase >= h->hp.nphases || !tau || !w || !D) return fail(h, ECUDA_ERR_ARG, "bad argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    Collocation& c = h->hp.col[phase];
    std::memcpy(c.tau.data(), tau, sizeof(double) * c.N);
    std::memcpy(c.w.data(), w, sizeof(double) * c.N);
    std::memcpy(c.D.data(), D, sizeof(double) * c.N * c.N);
    int rc = upload_collocation(h, phase);
    return rc ? rc : upload_template(h);
}

int ecuda_set_scaling(ecuda_handle h, const double* sz, const double* sg, double sf) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < h->pd.nvars; ++c) {
        h->h_sz[c] = sz ? sz[c] : 1.0;
        if (!(h->h_sz[c] > 0.0)) return fail(h, ECUDA_ERR_ARG, "variable scale factors must be positive");
    }
    for (int r = 0; r < h->pd.ncons; ++r) h->h_sg[r] = sg ? sg[r] : 1.0;
    h->pd.sf = sf;
    return upload_scaling(h);
}

int ecuda_upload_instances(ecuda_handle h, const double* inst, int memkind) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!inst) return fail(h, ECUDA_ERR_ARG, "null instance data");
    CU(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * h->hp.desc.batch * h->pd.inst_stride;
    CU(cudaMemcpyAsync(h->inst.p, inst, bytes,
                       memkind == ECUDA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_inst = true;
    return ECUDA_OK;
}

int ecuda_upload_bounds(ecuda_handle h, const double* gl, const double* gu, int memkind) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!gl || !gu) return fail(h, ECUDA_ERR_ARG, "null bounds");
    CU(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * h->hp.desc.batch * h->pd.ncons;
    int rc;
    if ((rc = ensure(h, h->gl, bytes))) return rc;
    if ((rc = ensure(h, h->gu, bytes))) return rc;
    auto kind = memkind == ECUDA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpyAsync(h->gl.p, gl, bytes, kind, h->stream));
    CU(cudaMemcpyAsync(h->gu.p, gu, bytes, kind, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_bounds = true;
    return ECUDA_OK;
}

static int eval_common(ecuda_handle h, const double* x, double* f, double* g, double* jac, double* grad,
                       int jac_mode, int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_inst) return fail(h, ECUDA_ERR_STATE, "upload_instances has not been called");
    if (!x) return fail(h, ECUDA_ERR_ARG, "null decision vector");
    if (jac_mode != ECUDA_JAC_EXACT && jac_mode != ECUDA_JAC_FD_INDEXSET) return fail(h, ECUDA_ERR_ARG, "bad jac_mode");
    if (memkind != ECUDA_MEM_HOST && memkind != ECUDA_MEM_DEVICE) return fail(h, ECUDA_ERR_ARG, "bad memkind");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const size_t B = h->hp.desc.batch, nv = h->pd.nvars, ng = h->pd.ncons, nz = h->pd.nnz;
    EvalIO io{};
    io.inst = static_cast<const double*>(h->inst.p);
    io.fpart = static_cast<double*>(h->fpart.p);
    io.jac_mode = jac_mode;
    io.batch = static_cast<int>(B);
    if (memkind == ECUDA_MEM_DEVICE) {
        io.x = x;
        io.f = f;
        io.g = g;
        io.jac = jac;
        io.grad = grad;
        return launch_eval(h, io, st);
    }
    int rc;
    if ((rc = ensure(h, h->sx, sizeof(double) * B * nv))) return rc;
    if (f && (rc = ensure(h, h->sf_, sizeof(double) * B))) return rc;
    if (g && (rc = ensure(h, h->sgv, sizeof(double) * B * ng))) return rc;
    if (jac && (rc = ensure(h, h->sjac, sizeof(double) * B * nz))) return rc;
    if (grad && (rc = ensure(h, h->sgrad, sizeof(double) * B * nv))) return rc;
    CU(cudaMemcpyAsync(h->sx.p, x, sizeof(double) * B * nv, cudaMemcpyHostToDevice, st));
    io.x = static_cast<const double*>(h->sx.p);
    io.f = f ? static_cast<double*>(h->sf_.p) : nullptr;
    io.g = g ? static_cast<double*>(h->sgv.p) : nullptr;
    io.jac = jac ? static_cast<double*>(h->sjac.p) : nullptr;
    io.grad = grad ? static_cast<double*>(h->sgrad.p) : nullptr;
    if ((rc = launch_eval(h, io, st))) return rc;
    if (f) CU(cudaMemcpyAsync(f, io.f, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    if (g) CU(cudaMemcpyAsync(g, io.g, sizeof(double) * B * ng, cudaMemcpyDeviceToHost, st));
    if (jac) CU(cudaMemcpyAsync(jac, io.jac, sizeof(double) * B * nz, cudaMemcpyDeviceToHost, st));
    if (grad) CU(cudaMemcpyAsync(grad, io.grad, sizeof(double) * B * nv, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ECUDA_OK;
}

int ecuda_eval(ecuda_handle h, const double* x, double* f, double* g, double* jac, int jac_mode, int memkind,
               void* stream) {
    return eval_common(h, x, f, g, jac, nullptr, jac_mode, memkind, stream);
}

int ecuda_eval_allgather(ecuda_handle h, const double* x, double* f, double* g, double* jac, int jac_mode,
                         void* const* peer_out, int nranks, int rank, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_inst) return fail(h, ECUDA_ERR_STATE, "upload_instances has not been called");
    if (!h->have_bounds) return fail(h, ECUDA_ERR_STATE, "upload_bounds has not been called");
    if (!h->fast_ok || h->pd.nphases != 1)
        return fail(h, ECUDA_ERR_STATE, "the fused summary needs a single-phase problem on the specialised kernels; "
                                        "use ecuda_eval + ecuda_summarize_allgather");
    if (!x || !f || !g || !peer_out) return fail(h, ECUDA_ERR_ARG, "x, f, g and peer_out are required");
    if (jac_mode != ECUDA_JAC_EXACT && jac_mode != ECUDA_JAC_FD_INDEXSET) return fail(h, ECUDA_ERR_ARG, "bad jac_mode");
    if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return fail(h, ECUDA_ERR_ARG, "bad rank / nranks (1..16)");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    EvalIO io{};
    io.inst = static_cast<const double*>(h->inst.p);
    io.fpart = static_cast<double*>(h->fpart.p);
    io.jac_mode = jac_mode;
    io.batch = h->hp.desc.batch;
    io.x = x;
    io.f = f;
    io.g = g;
    io.jac = jac;
    io.bl = static_cast<const double*>(h->gl.p);
    io.bu = static_cast<const double*>(h->gu.p);
    io.nranks = nranks;
    io.rank = rank;
    for (int r = 0; r < nranks; ++r) {
        if (!peer_out[r] || (reinterpret_cast<uintptr_t>(peer_out[r]) & 15)) return fail(h, ECUDA_ERR_ARG, "peer buffer null or not 16-byte aligned");
        io.peer[r] = static_cast<double*>(peer_out[r]);
    }
    return launch_eval(h, io, st);
}

int ecuda_eval_grad_f(ecuda_handle h, const double* x, double* grad, int memkind, void* stream) {
    if (h && !grad) return fail(h, ECUDA_ERR_ARG, "null gradient output");
    return eval_common(h, x, nullptr, nullptr, nullptr, grad, ECUDA_JAC_EXACT, memkind, stream);
}

int ecuda_summary(ecuda_handle h, const double* x, double* out, int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_bounds) return fail(h, ECUDA_ERR_STATE, "upload_bounds has not been called");
    if (!out) return fail(h, ECUDA_ERR_ARG, "null output");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const size_t B = h->hp.desc.batch, ng = h->pd.ncons;
    int rc;
    if ((rc = ensure(h, h->sf_, sizeof(double) * B))) return rc;
    if ((rc = ensure(h, h->sgv, sizeof(double) * B * ng))) return rc;
    const double* xd = x;
    if (memkind == ECUDA_MEM_HOST) {
        if ((rc = ensure(h, h->sx, sizeof(double) * B * h->pd.nvars))) return rc;
        CU(cudaMemcpyAsync(h->sx.p, x, sizeof(double) * B * h->pd.nvars, cudaMemcpyHostToDevice, st));
        xd = static_cast<const double*>(h->sx.p);
    }
    rc = eval_common(h, xd, static_cast<double*>(h->sf_.p), static_cast<double*>(h->sgv.p), nullptr, nullptr,
                     ECUDA_JAC_EXACT, ECUDA_MEM_DEVICE, st);
    if (rc) return rc;
    double* od = out;
    if (memkind == ECUDA_MEM_HOST) {
        if ((rc = ensure(h, h->ssum, sizeof(double) * 2 * B))) return rc;
        od = static_cast<double*>(h->ssum.p);
    }
    const int warps_per_block = 4;
    k_summary<<<(unsigned)((B + warps_per_block - 1) / warps_per_block), 32 * warps_per_block, 0, st>>>(
        static_cast<const double*>(h->sf_.p), static_cast<const double*>(h->sgv.p),
        static_cast<const double*>(h->gl.p), static_cast<const double*>(h->gu.p), od, (int)B, (int)ng);
    ++h->launches;
    CU(cudaGetLastError());
    if (memkind == ECUDA_MEM_HOST) {
        CU(cudaMemcpyAsync(out, od, sizeof(double) * 2 * B, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ECUDA_OK;
}

int ecuda_summarize(ecuda_handle h, const double* f_dev, const double* g_dev, double* out_dev, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_bounds) return fail(h, ECUDA_ERR_STATE, "upload_bounds has not been called");
    if (!f_dev || !g_dev || !out_dev) return fail(h, ECUDA_ERR_ARG, "null pointer");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const size_t B = h->hp.desc.batch;
    const int warps_per_block = 4;
    k_summary<<<(unsigned)((B + warps_per_block - 1) / warps_per_block), 32 * warps_per_block, 0, st>>>(
        f_dev, g_dev, static_cast<const double*>(h->gl.p), static_cast<const double*>(h->gu.p), out_dev, (int)B,
        h->pd.ncons);
    ++h->launches;
    CU(cudaGetLastError());
    return ECUDA_OK;
}

int ecuda_summarize_allgather(ecuda_handle h, const double* f_dev, const double* g_dev, void* const* peer_out,
                              int nranks, int rank, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_bounds) return fail(h, ECUDA_ERR_STATE, "upload_bounds has not been called");
    if (!f_dev || !g_dev || !peer_out) return fail(h, ECUDA_ERR_ARG, "null pointer");
    if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return fail(h, ECUDA_ERR_ARG, "bad rank / nranks (1..16)");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    PeerPtrs pp{};
    for (int r = 0; r < nranks; ++r) {
        if (!peer_out[r] || (reinterpret_cast<uintptr_t>(peer_out[r]) & 15)) return fail(h, ECUDA_ERR_ARG, "peer buffer null or not 16-byte aligned");
        pp.p[r] = static_cast<double*>(peer_out[r]);
    }
    const size_t B = h->hp.desc.batch;
    const int warps_per_block = 4;
    k_summary_scatter<<<(unsigned)((B + warps_per_block - 1) / warps_per_block), 32 * warps_per_block, 0, st>>>(
        f_dev, g_dev, static_cast<const double*>(h->gl.p), static_cast<const double*>(h->gu.p), pp, nranks, rank,
        (int)B, h->pd.ncons);
    ++h->launches;
    CU(cudaGetLastError());
    return ECUDA_OK;
}

int ecuda_peer_barrier(ecuda_handle h, void* const* peer_flags, int nranks, int rank, uint64_t step, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!peer_flags || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || step == 0)
        return fail(h, ECUDA_ERR_ARG, "bad peer_flags / rank / nranks (1..16) / step (>= 1)");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    PeerFlags pf{};
    for (int r = 0; r < nranks; ++r) {
        if (!peer_flags[r]) return fail(h, ECUDA_ERR_ARG, "null peer flag array");
        pf.p[r] = static_cast<unsigned long long*>(peer_flags[r]);
    }
    int rc;
    if ((rc = ensure(h, h->bflag, sizeof(int)))) return rc;
    k_peer_barrier<<<1, 32, 0, st>>>(pf, nranks, rank, static_cast<unsigned long long>(step), static_cast<int*>(h->bflag.p));
    ++h->launches;
    CU(cudaGetLastError());
    return ECUDA_OK;
}

int ecuda_sync(ecuda_handle h) {
    if (!h) return ECUDA_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return ECUDA_OK;
}

int64_t ecuda_launch_count(ecuda_handle h) { return h ? h->launches : 0; }

int ecuda_host_model_eval(int model, const double* x, const double* u, double t, double* f_out, double* cost_out) {
    if (!x || !u || !f_out || !cost_out) return ECUDA_ERR_ARG;
    switch (model) {
        case ECUDA_MODEL_SI2D: host_model_eval_t<ECUDA_MODEL_SI2D>(x, u, t, f_out, cost_out); return ECUDA_OK;
        case ECUDA_MODEL_PM3D: host_model_eval_t<ECUDA_MODEL_PM3D>(x, u, t, f_out, cost_out); return ECUDA_OK;
        case ECUDA_MODEL_FW6: host_model_eval_t<ECUDA_MODEL_FW6>(x, u, t, f_out, cost_out); return ECUDA_OK;
    }
    return ECUDA_ERR_ARG;
}
int ecuda_host_path_eval(const ecuda_problem_desc* desc, const double* inst, double x, double y, double t, double* rows) {
    if (!desc || !inst || !rows) return ECUDA_ERR_ARG;
    HostProblem hp;
    std::string err;
    if (!build_layout(*desc, &hp, &err)) return ECUDA_ERR_ARG;
    const int nstat = hp.nstat[0], rec = hp.dims.rec_size;
    for (int q = 0; q < nstat; ++q) {
        const double* r = inst + hp.inst_off[0] + q * rec;
        rows[q] = desc->model == ECUDA_MODEL_SI2D ? edge_row(r, x, y) : cylinder_row(r, x, y);
    }
    for (int i = 0; i < desc->ntracks; ++i)
        rows[nstat + i] = track_row(inst + hp.track_off + i * hp.dims.track_size, desc->nwaypoints, x, y, t);
    return ECUDA_OK;
}

int ecuda_fp64_peak(ecuda_handle h, double* tflops) {
    if (!h) return ECUDA_ERR_ARG;
    if (!tflops) return fail(h, ECUDA_ERR_ARG, "null output");
    CU(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    int rc;
    if ((rc = ensure(h, h->ssum, sizeof(double) * blocks * threads))) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
        CU(cudaEventRecord(e0, h->stream));
        k_fp64_peak<<<blocks, threads, 0, h->stream>>>(static_cast<double*>(h->ssum.p), iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1, h->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CU(cudaGetLastError());
    ++h->launches;
    *tflops = best;
    return ECUDA_OK;
}

// ---- IPOPT TNLP-shaped shims (single instance) -----------------------------------------------------------
static int ipopt_guard(ecuda_handle h, int n) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (h->hp.desc.batch != 1) return fail(h, ECUDA_ERR_STATE, "IPOPT shims need batch == 1");
    if (n != h->pd.nvars) return fail(h, ECUDA_ERR_ARG, "n does not match nvars");
    return ECUDA_OK;
}
int ecuda_set_ipopt_jac_mode(ecuda_handle h, int jac_mode) {
    if (!h) return ECUDA_ERR_ARG;
    if (jac_mode != ECUDA_JAC_EXACT && jac_mode != ECUDA_JAC_FD_INDEXSET) return fail(h, ECUDA_ERR_ARG, "bad jac_mode");
    h->ipopt_jac_mode = jac_mode;
    return ECUDA_OK;
}
int ecuda_ipopt_eval_f(ecuda_handle h, int n, const double* x, int new_x, double* obj) {
    (void)new_x;
    int rc = ipopt_guard(h, n);
    if (rc) return rc;
    return eval_common(h, x, obj, nullptr, nullptr, nullptr, ECUDA_JAC_EXACT, ECUDA_MEM_HOST, nullptr);
}
int ecuda_ipopt_eval_grad_f(ecuda_handle h, int n, const double* x, int new_x, double* grad) {
    (void)new_x;
    int rc = ipopt_guard(h, n);
    if (rc) return rc;
    return eval_common(h, x, nullptr, nullptr, nullptr, grad, ECUDA_JAC_EXACT, ECUDA_MEM_HOST, nullptr);
}
int ecuda_ipopt_eval_g(ecuda_handle h, int n, const double* x, int new_x, int m, double* g) {
    (void)new_x;
    int rc = ipopt_guard(h, n);
    if (rc) return rc;
    if (m != h->pd.ncons) return fail(h, ECUDA_ERR_ARG, "m does not match ncons");
    return eval_common(h, x, nullptr, g, nullptr, nullptr, ECUDA_JAC_EXACT, ECUDA_MEM_HOST, nullptr);
}
int ecuda_ipopt_eval_jac_g(ecuda_handle h, int n, const double* x, int new_x, int m, int nele_jac, int32_t* iRow,
                           int32_t* jCol, double* values) {
    (void)new_x;
    int rc = ipopt_guard(h, n);
    if (rc) return rc;
    if (m != h->pd.ncons || nele_jac != h->pd.nnz) return fail(h, ECUDA_ERR_ARG, "m / nele_jac mismatch");
    if (!values) return ecuda_get_structure(h, iRow, jCol, nullptr);
    return eval_common(h, x, nullptr, nullptr, values, nullptr, h->ipopt_jac_mode, ECUDA_MEM_HOST, nullptr);
}

}  // extern "C"
