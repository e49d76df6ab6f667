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
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <tuple>
#include <string>
#include <vector>

#include "ecuda_kernels.cuh"

// This file is compiled as two translation units so that the build takes half the time: the main one, and (with
// ECUDA_TU_ROWSN defined, through ecuda_api_rowsn.cu) one that holds only the launchers -- and therefore the
// instantiations -- of the N-specialised kernel family (k_rows_n, k_rows_n_fd_persist, k_stream_exact).
namespace ecuda {
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};
}  // namespace ecuda

#ifndef ECUDA_TU_ROWSN
namespace ecuda {

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
// array has reached s. Counters only grow, so nothing is ever reset. The wait is bounded (timeout_ns of the
// global timer; default 10 s, ECUDA_PEER_TIMEOUT_MS): a lost peer does not hang the GPU. When it expires the kernel
// records {1, step, first late rank + 1} in the handle's status words, which are STICKY: every later ecuda_sync /
// ecuda_peer_barrier_status on the handle reports the failure until the status is read with reset.
struct PeerFlags {
    unsigned long long* p[16];
};
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_peer_barrier(PeerFlags flags, int nranks, int rank, unsigned long long step,
                               unsigned long long timeout_ns, unsigned long long* status) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    __threadfence_system();
    volatile unsigned long long* theirs = flags.p[r] + rank;  // my slot in rank r's array
    *theirs = step;
    volatile unsigned long long* mine = flags.p[rank] + r;    // rank r's slot in my array
    const unsigned long long t0 = global_timer_ns();
    while (*mine < step) {
        if (global_timer_ns() - t0 > timeout_ns) {
            if (atomicCAS(&status[0], 0ull, 1ull) == 0ull) {  // first failure wins, later ones keep it
                status[1] = step;
                status[2] = static_cast<unsigned long long>(r) + 1ull;
            }
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
// decision vectors interpolated onto another mesh (ecuda_resample): new node values = R * old node values
// per state / control, t0 and tf copied; input scaled with the problem's s_z, output with sz_new (or unscaled)
struct ResampleDev {
    const double* R;       // per phase [Nnew][N], back to back
    int moff[ECUDA_MAX_PHASES], Nnew[ECUDA_MAX_PHASES], zoff_new[ECUDA_MAX_PHASES];
    int nvars_new;
    const double* x;       // [B][nvars]
    double* x_new;         // [B][nvars_new]
    const double* sz_new;  // [nvars_new] or null
};
__global__ void k_resample(const __grid_constant__ ProbDev pb, const __grid_constant__ ResampleDev rd) {
    const int b = blockIdx.x / pb.nphases, p = blockIdx.x - b * pb.nphases;
    const PhaseDev& ph = pb.ph[p];
    const int N = ph.N, Nn = rd.Nnew[p], ns = pb.ns, nc = pb.nc, per = ns + nc;
    const double* x = rd.x + static_cast<size_t>(b) * pb.nvars + ph.zoff;
    const double* isz = pb.isz + ph.zoff;
    double* out = rd.x_new + static_cast<size_t>(b) * rd.nvars_new + rd.zoff_new[p];
    const double* R = rd.R + rd.moff[p];
    for (int it = threadIdx.x; it < Nn * per + 2; it += blockDim.x) {
        double v;
        int dst;
        if (it >= Nn * per) {  // t0, tf
            const int w = it - Nn * per;
            v = x[per * N + w] * isz[per * N + w];
            dst = per * Nn + w;
        } else {
            const int k = it / per, c = it - k * per;  // c < nc: control c, else state c - nc
            const int stride = c < nc ? nc : ns, base = c < nc ? c : nc * N + (c - nc);
            double acc = 0.0;
            for (int l = 0; l < N; ++l) acc = fma(__ldg(R + k * N + l), x[base + l * stride] * isz[base + l * stride], acc);
            v = acc;
            dst = c < nc ? k * nc + c : nc * Nn + k * ns + (c - nc);
        }
        if (rd.sz_new) v = v * rd.sz_new[rd.zoff_new[p] + dst];
        out[dst] = v;
    }
}

// ecuda_upload_bounds: do all instances share "defect bounds 0" and one pair of path-row bounds? (bitwise
// comparison, so that +-inf are ordinary values). flags[0] is cleared on the first counter-example.
__global__ void k_bounds_classify(const double* gl, const double* gu, int batch, int ncons, int ndef, int ne, int* flags) {
    const size_t n = static_cast<size_t>(batch) * ncons;
    const long long plo = __double_as_longlong(gl[ndef + ne]), phi = __double_as_longlong(gu[ndef + ne]);
    bool ok = true;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < n; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(e % ncons);
        const long long lo = __double_as_longlong(gl[e]), hi = __double_as_longlong(gu[e]);
        if (r < ndef) ok = ok && lo == 0 && hi == 0;  // +0.0 exactly
        else if (r >= ndef + ne && r < ncons - 1) ok = ok && lo == plo && hi == phi;
    }
    if (!ok) flags[0] = 0;
}
__global__ void k_bounds_compact(const double* gl, const double* gu, int batch, int ncons, int ndef, int ne, double* bev) {
    const int nev = ne + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * nev) return;
    const int b = i / nev, e = i - b * nev;
    const size_t src = static_cast<size_t>(b) * ncons + (e < ne ? ndef + e : ncons - 1);
    bev[static_cast<size_t>(b) * 2 * nev + e] = gl[src];
    bev[static_cast<size_t>(b) * 2 * nev + nev + e] = gu[src];
}

// ecuda_eval_compact: the per-instance triplets of an exact Jacobian gathered out of the full array.
// One CTA row per instance (grid.y) so the index list is read once per CTA column and stays in L1/L2;
// loads are streaming (the full array is scratch), stores are coalesced.
__global__ void __launch_bounds__(256) k_gather_local(const double* __restrict__ jac, const int32_t* __restrict__ idx,
                                                      double* __restrict__ out, int nnz, int nlocal) {
    const size_t b = blockIdx.y;
    const double* src = jac + b * static_cast<size_t>(nnz);
    double* dst = out + b * static_cast<size_t>(nlocal);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + __ldg(idx + i)));
}

// Pulls the inputs of a batch (decision vectors, instance records) into L2 before the evaluation kernel starts
// writing: a CTA's first loads then hit L2 instead of queueing in HBM behind the kernel's own write stream.
__global__ void __launch_bounds__(256) k_prefetch_inputs(const char* a, size_t abytes, const char* b, size_t bbytes) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 128;
    for (size_t o = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 128; o < abytes; o += stride)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a + o));
    for (size_t o = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 128; o < bbytes; o += stride)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(b + o));
}

}  // namespace ecuda
#endif  // !ECUDA_TU_ROWSN

using namespace ecuda;

constexpr int kMaxHostChunks = 8;

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
    size_t smem_generic_exact = 0;  // generic kernel without the finite-difference arrays
    size_t smem_isz = 0;            // extra shared memory of the exact row-owner kernel (CARVE_ISZ)
    bool fast_ok = false;  // the specialised kernels (ecuda_fast.cuh) can run this problem
    bool no_fast = false;  // ECUDA_NO_FAST=1 in the environment: never use them (A/B runs and tests)
    bool no_image = true;   // ECUDA_IMAGE=1 opts in to the persistent image kernel (exact mode); measured
                            // slower than the copy-warp kernel on C2 (0.180 vs 0.157 ms: 2 CTAs/SM leave the
                            // latency-bound node-local phases exposed), kept for the next round's pipelining work
    bool image_ok = false;
    size_t smem_image = 0;
    int num_sms = 148;
    bool rows_fill = false; // every phase has at least 7/8 * kThreads defect rows (see launch_keval_fast)
    bool no_rows = false;   // ECUDA_NO_ROWS=1: use the column-owner kernels (k_eval_fast) instead of k_eval_rows
    bool no_copy_warp = false;  // ECUDA_NO_COPY_WARP=1: exact mode copies the template with plain loads/stores
    int64_t launches = 0;
    bool barrier_used = false;  // ecuda_peer_barrier has been issued: ecuda_sync also reports its sticky status
    int ipopt_jac_mode = ECUDA_JAC_EXACT;
    bool force_generic = false;  // ECUDA_FORCE_GENERIC=1 in the environment: always run the generic kernel
    bool prefetch = false;    // ECUDA_PREFETCH=1: k_prefetch_inputs before the evaluation kernel
    bool persist = false;     // ECUDA_PERSIST=1: finite differences on the persistent kernel (k_rows_n_fd_persist)
    int exact_kernel = 0;     // ECUDA_EXACT_KERNEL: 0 k_eval_rows (default), 1 "ring" k_rows_n, 2 "stream" k_stream_exact
    DevBuf desc;              // exact-mode triplet descriptors
    int rowsn_N = 0;     // node count shared by all phases when the N-specialised kernels may run (else 0)
    int nb_uniform = 0;  // summation-block count if all phases share it and it is <= 8, else 0
    std::vector<double> h_sz, h_sg;
    // user model (ecuda_register_user_model): kernels compiled with NVRTC, loaded per handle
    const UserModel* um = nullptr;
    cudaLibrary_t ulib = nullptr;
    cudaKernel_t ukern[UserImage::NKERNELS] = {};
    size_t ukern_smem[UserImage::NKERNELS] = {};
    DevBuf slam, ssig, shess;  // staging for ecuda_eval_hess with host buffers
    DevBuf bcls;               // flag word of k_bounds_classify
    DevBuf bev;                // compact bounds for the fused summary (see EvalIO::bev); valid when bounds_compact
    bool bounds_compact = false;
    double plo = 0.0, phi = 0.0;
    // mesh-refinement support (ecuda_ode_error / ecuda_resample), built on first use
    DevBuf mesh[ECUDA_MAX_PHASES];  // E | dE | wq | tq per phase
    bool have_mesh = false;
    DevBuf serr, resmat, sxnew, ssznew;
    // compact exact output (ecuda_eval_compact): indices of the per-instance triplets, built on first use
    std::vector<int32_t> local_index;
    std::vector<int32_t> hess_ir, hess_jc;  // Hessian structure, built on first request
    DevBuf lidx, sjl;
    // HOST-buffer calls are pipelined over instance chunks (eval_host): copy streams and their events
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_done = nullptr, ev_x[kMaxHostChunks] = {}, ev_k[kMaxHostChunks] = {};
    int host_chunks = 0;  // ECUDA_HOST_CHUNKS=1..8 in the environment; 0 = automatic (4 when the results exceed 16 MB)
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

#ifndef ECUDA_TU_ROWSN
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
    if ((rc = ensure(h, h->isz, sizeof(double) * nv + 16))) return rc;
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

// CTAs per (instance, phase) of the generic kernel. At least enough for one defect row per thread (the widest
// phase decides; extra slices of a narrower phase just own fewer nodes). A small batch cannot fill the GPU
// that way -- a single 200-node instance would run on 5 of 148 SMs -- so the slices are made narrower until the
// grid covers the SMs (latency of one IPOPT-style evaluation), down to 4 nodes per slice.
static int slices_for(const ecuda_ctx* h, int grid) {
    int nslices = 1, nmin = 1 << 30;
    for (int p = 0; p < h->pd.nphases; ++p) {
        nslices = std::max(nslices, generic_slices(h->pd.ns, h->pd.ph[p].N, kThreads));
        nmin = std::min(nmin, h->pd.ph[p].N);
    }
    if (const char* e = std::getenv("ECUDA_SLICES")) return std::max(1, std::min(std::atoi(e), nmin));
    const int fill = h->num_sms / std::max(1, grid);
    return std::max(nslices, std::min(fill, std::max(1, nmin / 4)));
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
    const int nslices = slices_for(h, grid);
    const bool fd = io.jac != nullptr && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    k_eval<M, NB><<<dim3(grid, nslices), kThreads, fd ? h->smem_bytes : h->smem_generic_exact, st>>>(h->pd, io);
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
    // finite differences: the row-owner layout gives the heavy D-coupled work to the threads that own a defect
    // row, so it pays only when the defect rows (almost) fill the CTA -- C2: 240 of 256 threads, 7 % faster
    // than the column-owner layout; C0 (66 rows) 28 % slower, C4 (180 rows) 5 % slower. Exact mode: always.
    if (!h->no_rows && (!FD || h->rows_fill)) {
        static size_t configured_rows[64] = {0};
        size_t smem_rows = smem + (FD ? 0 : h->smem_isz);  // exact: 1/sz staged in shared memory (CARVE_ISZ)
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

#endif  // !ECUDA_TU_ROWSN

// defined in the ECUDA_TU_ROWSN translation unit: the N-specialised kernels for the handle's model and node count;
// returns 1 when no instantiation matches (the caller falls back)
int ecuda_launch_rows_n(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid);

#ifdef ECUDA_TU_ROWSN
// N-specialised row-owner kernels (ecuda_rowsn.cuh): instantiated ahead of time for the node counts of the
// BASELINE configurations; other shapes take the kernels above
template <int M, int N, bool FD, bool TRK, bool SUM>
static int launch_rows_n_t(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    constexpr bool RING = !FD || ECUDA_RN_FD_RING;  // see k_rows_n
    static std::mutex mu;
    static size_t configured[64] = {0};
    size_t ring = 0;  // the store ring: three buffers of the largest node group of any phase
    for (int p = 0; RING && p < h->pd.nphases; ++p) ring = std::max(ring, kRnBufs * rn_group_cap<M>(h->pd, h->pd.ph[p], N));
    size_t smem = (rn_doubles<M>(h->pd, N, FD) + ring) * sizeof(double);
    if (SUM && !io.bev) smem += 2 * static_cast<size_t>(phase_ncons(h->pd, h->pd.ph[0]) + 2) * sizeof(double);
    if (smem > 227 * 1024 - 1024) return 1;  // does not fit one CTA: the caller falls back
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < smem) {
            CU(cudaFuncSetAttribute(k_rows_n<M, N, FD, TRK, SUM, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    k_rows_n<M, N, FD, TRK, SUM, RING><<<grid, kThreads, smem, st>>>(h->pd, io);
    return ECUDA_OK;
}
// finite differences, persistent CTAs with TMA prefetch of the next instance (k_rows_n_fd_persist)
template <int M, int N, bool TRK>
static int launch_rows_n_persist(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    const size_t nv = static_cast<size_t>(rn_nv<M>(h->pd, N)), nve = nv + (nv & 1);
    const size_t smem = (4 * nve + 2 * static_cast<size_t>(h->pd.inst_stride) + 4 * nv + static_cast<size_t>(N) * N +
                         2 * static_cast<size_t>(N + (N & 1))) * sizeof(double);
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < smem) {
            CU(cudaFuncSetAttribute(k_rows_n_fd_persist<M, N, TRK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    const int grid = std::min(io.batch, h->num_sms * ECUDA_MIN_CTAS_ROWSN_FD);
    k_rows_n_fd_persist<M, N, TRK><<<grid, kThreads, smem, st>>>(h->pd, io);
    return ECUDA_OK;
}
// exact Jacobian / values only: the streaming kernel (ecuda_stream.cuh)
template <int M, int N, bool TRK, bool SUM>
static int launch_stream_t(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    size_t smem = 0;
    for (int p = 0; p < h->pd.nphases; ++p) smem = std::max(smem, st_doubles<M>(h->pd, h->pd.ph[p], N) * sizeof(double));
    if (SUM && !io.bev) smem += 2 * static_cast<size_t>(phase_ncons(h->pd, h->pd.ph[0]) + 2) * sizeof(double);
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < smem) {
            CU(cudaFuncSetAttribute(k_stream_exact<M, N, TRK, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    k_stream_exact<M, N, TRK, SUM><<<grid, kThreads, smem, st>>>(h->pd, io);
    return ECUDA_OK;
}
// TRK: instantiated with track rows (moving zones) or without; a problem whose model is only instantiated without
// them and has tracks falls back (returns 1)
template <int M, int N, bool TRK>
static int launch_rows_n_mn(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    if (!TRK && h->pd.ntracks > 0) return 1;
    const bool fd = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    if (fd) {
        // Round 1 / early round 2 kept phases with few defect rows (C0: 66 of 256 threads) on the column-owner kernel
        // (0.164 vs 0.207 ms). Since the path-row work items are grouped by kind (path_item) the N-specialised kernel
        // wins there too: C0 0.133 vs 0.147 ms, the 17-node variant 0.098 vs 0.120 ms. ECUDA_NO_ROWSN=1 switches back.
        if (h->persist && io.nranks == 0 && h->pd.nphases == 1 && (h->pd.nvars & 1) == 0 && (h->pd.inst_stride & 1) == 0 &&
            (reinterpret_cast<uintptr_t>(io.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(io.inst) & 15) == 0)
            return launch_rows_n_persist<M, N, TRK>(h, io, st);
        return io.nranks > 0 ? launch_rows_n_t<M, N, true, TRK, true>(h, io, st, grid)
                             : launch_rows_n_t<M, N, true, TRK, false>(h, io, st, grid);
    }
    // Exact mode. Measured on C2 (B200, same box, profiles/r2/README.md): round-1 k_eval_rows (TMA copy of the
    // template + node-local overwrite) 0.144 ms, k_rows_n with the store ring 0.154 ms, k_stream_exact 0.165 ms. The
    // fastest one stays the default; the other two are selected with ECUDA_EXACT_KERNEL=ring | stream (A/B, tests).
    if (h->exact_kernel == 1)
        return io.nranks > 0 ? launch_rows_n_t<M, N, false, TRK, true>(h, io, st, grid)
                             : launch_rows_n_t<M, N, false, TRK, false>(h, io, st, grid);
    if (h->exact_kernel == 2 && h->pd.ncons <= 65535 && h->pd.nvars <= 65535)  // 16-bit descriptor fields
        return io.nranks > 0 ? launch_stream_t<M, N, TRK, true>(h, io, st, grid)
                             : launch_stream_t<M, N, TRK, false>(h, io, st, grid);
    return 1;  // the caller falls back to k_eval_rows
}
// returns 1 when no instantiation matches (the caller falls back)
template <int M>
static int launch_rows_n(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    if (M == ECUDA_MODEL_PM3D) {
        if (h->rowsn_N == 40) return launch_rows_n_mn<M, 40, false>(h, io, st, grid);
        if (h->rowsn_N == 30) return launch_rows_n_mn<M, 30, false>(h, io, st, grid);
    }
    if (M == ECUDA_MODEL_SI2D) {
        if (h->rowsn_N == 33) return launch_rows_n_mn<M, 33, true>(h, io, st, grid);
        if (h->rowsn_N == 17) return launch_rows_n_mn<M, 17, true>(h, io, st, grid);
    }
    return 1;
}
int ecuda_launch_rows_n(ecuda_ctx* h, const EvalIO& io, cudaStream_t st, int grid) {
    switch (h->pd.model) {
        case ECUDA_MODEL_SI2D: return launch_rows_n<ECUDA_MODEL_SI2D>(h, io, st, grid);
        case ECUDA_MODEL_PM3D: return launch_rows_n<ECUDA_MODEL_PM3D>(h, io, st, grid);
        default: return 1;
    }
}
#else  // main translation unit from here on

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
        if (h->prefetch) {
            k_prefetch_inputs<<<h->num_sms * 2, 256, 0, st>>>(reinterpret_cast<const char*>(io.x),
                                                            sizeof(double) * io.batch * h->pd.nvars,
                                                            reinterpret_cast<const char*>(io.inst),
                                                            sizeof(double) * io.batch * h->pd.inst_stride);
            ++h->launches;
        }
        int rc = h->rowsn_N > 0 ? ecuda_launch_rows_n(h, io, st, grid) : 1;
        if (rc <= 0) {
        } else if (h->fast_ok) {
            rc = ECUDA_OK;
            switch (h->nb_uniform) {
                case 3: rc = launch_keval_fast_mode<M, 3>(h, io, st, grid); break;
                case 4: rc = launch_keval_fast_mode<M, 4>(h, io, st, grid); break;
                default: rc = launch_keval_fast_mode<M, 5>(h, io, st, grid); break;
            }
        } else {
        rc = ECUDA_OK;
        switch (h->nb_uniform) {  // block count shared by all phases, or 0
            // specialised for the node counts of the BASELINE configs: 17 -> 3, 30 -> 4, 33 / 40 -> 5
            case 3: rc = launch_keval<M, 3>(h, io, st, grid); break;
            case 4: rc = launch_keval<M, 4>(h, io, st, grid); break;
            case 5: rc = launch_keval<M, 5>(h, io, st, grid); break;
            default: rc = launch_keval<M, 0>(h, io, st, grid); break;
        }
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

// ---- user models: the same kernels, compiled at run time for Model<ECUDA_MODEL_USER> -------------------------
static int unload_user_kernels(ecuda_ctx* h) {
    if (h->ulib) cudaLibraryUnload(h->ulib);
    h->ulib = nullptr;
    for (int k = 0; k < UserImage::NKERNELS; ++k) {
        h->ukern[k] = nullptr;
        h->ukern_smem[k] = 0;
    }
    return ECUDA_OK;
}

static int load_user_kernels(ecuda_ctx* h) {
    // one compilation per (model, block count, kernel set) and process
    static std::mutex mu;
    static std::map<std::tuple<int, int, bool, int, bool>, std::shared_ptr<UserImage>> cache;
    const int nb = (h->nb_uniform >= 3 && h->nb_uniform <= 5) ? h->nb_uniform : 0;
    // the N-specialised finite-difference kernel (k_rows_n) under the rule of launch_rows_n_mn: every phase has N
    // nodes and one defect row per thread
    const int ns = h->pd.ns, rn = h->rowsn_N;
    int rowsn = (h->fast_ok && rn > 0 && ns * rn <= kThreads) ? rn : 0;
    if (rowsn) {  // its shared memory (launch_eval_user) must fit one CTA, else the round-1 kernels run
        const size_t nv = static_cast<size_t>(ns + h->pd.nc) * rn + 2;
        if ((static_cast<size_t>(h->pd.inst_stride) + nv + 1 + 4 * nv + 1) * sizeof(double) > 227 * 1024 - 1024) rowsn = 0;
    }
    bool trk = false;
    for (int p = 0; p < h->pd.nphases; ++p) trk = trk || h->pd.ph[p].npath > h->pd.ph[p].nstat;
    std::shared_ptr<UserImage> img;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto key = std::make_tuple(h->um->id, nb, h->fast_ok, rowsn, trk);
        auto it = cache.find(key);
        if (it == cache.end()) {
            std::shared_ptr<UserImage> fresh(new UserImage);
            std::string err;
            if (!user_model_compile(*h->um, nb, h->fast_ok, rowsn, trk, fresh.get(), &err)) return fail(h, ECUDA_ERR_CUDA, err);
            it = cache.emplace(key, fresh).first;
        }
        img = it->second;
    }
    unload_user_kernels(h);
    CU(cudaLibraryLoadData(&h->ulib, img->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    for (int k = 0; k < UserImage::NKERNELS; ++k)
        if (!img->name[k].empty()) CU(cudaLibraryGetKernel(&h->ukern[k], h->ulib, img->name[k].c_str()));
    return ECUDA_OK;
}

static int launch_user(ecuda_ctx* h, int which, dim3 grid, int threads, size_t smem, const EvalIO& io, cudaStream_t st,
                       const void* mesh = nullptr) {
    cudaKernel_t k = h->ukern[which];
    if (!k) return fail(h, ECUDA_ERR_STATE, "user-model kernel was not compiled");
    if (smem > 48 * 1024 && h->ukern_smem[which] < smem) {
        CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->ukern_smem[which] = smem;
    }
    void* args[] = {const_cast<ProbDev*>(&h->pd), const_cast<EvalIO*>(&io), const_cast<void*>(mesh)};
    CU(cudaLaunchKernel(reinterpret_cast<const void*>(k), grid, dim3(threads), args, smem, st));
    ++h->launches;
    return ECUDA_OK;
}

static int launch_eval_user(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    const int grid = io.batch * h->pd.nphases;
    int rc;
    if (io.grad && (rc = launch_user(h, UserImage::GRAD, dim3(grid), kThreads, h->smem_bytes, io, st))) return rc;
    if (io.f || io.g || io.jac) {
        if (h->fast_ok) {  // k_eval_rows, same launch rules as launch_keval_fast
            const bool fd = io.jac && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
            const bool copy_warp = !fd && io.jac != nullptr && !h->no_copy_warp && (h->pd.nnz & 1) == 0 &&
                                   (reinterpret_cast<uintptr_t>(io.jac) & 15) == 0;
            size_t smem = fd ? h->smem_fast_fd
                             : h->smem_fast_exact + h->smem_isz +
                                   (copy_warp ? (kCopySlots * kCopyChunk + 2) * sizeof(double) : 0);
            if (io.nranks > 0) smem += 2 * static_cast<size_t>(phase_ncons(h->pd, h->pd.ph[0]) + 2) * sizeof(double);
            if (fd && io.nranks == 0 && h->ukern[UserImage::ROWSN_FD]) {
                // k_rows_n<USER, N, FD>: shared memory as rn_doubles (instance records, z, one FdRec per variable)
                const size_t nv = static_cast<size_t>(h->pd.ns + h->pd.nc) * h->rowsn_N + 2, nve = nv + (nv & 1);
                size_t n = static_cast<size_t>(h->pd.inst_stride) + nve + 4 * nv;
                n += n & 1;
                rc = launch_user(h, UserImage::ROWSN_FD, dim3(grid), kThreads, n * sizeof(double), io, st);
            } else
            rc = launch_user(h, fd ? UserImage::ROWS_FD : UserImage::ROWS_EXACT, dim3(grid),
                             kThreads + (copy_warp ? kCopyWarpThreads : 0), smem, io, st);
        } else {
            const int nslices = slices_for(h, grid);
            const bool fdg = io.jac != nullptr && io.jac_mode == ECUDA_JAC_FD_INDEXSET;
            rc = launch_user(h, UserImage::GENERIC, dim3(grid, nslices), kThreads,
                             fdg ? h->smem_bytes : h->smem_generic_exact, io, st);
        }
        if (rc) return rc;
        if (io.f && h->pd.nphases > 1) {
            k_sum_phases<<<(io.batch + 127) / 128, 128, 0, st>>>(io.fpart, io.f, io.batch, h->pd.nphases, h->pd.sf);
            ++h->launches;
        }
    }
    CU(cudaGetLastError());
    return ECUDA_OK;
}

static int launch_eval(ecuda_ctx* h, const EvalIO& io, cudaStream_t st) {
    if (h->um) return launch_eval_user(h, io, st);
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
    {
        const char* hc = std::getenv("ECUDA_HOST_CHUNKS");
        const int v = hc ? std::atoi(hc) : 0;
        h->host_chunks = v >= 1 && v <= kMaxHostChunks ? v : 0;
    }
    auto make_events = [&]() {
        cudaError_t r = cudaEventCreateWithFlags(&h->ev_entry, cudaEventDisableTiming);
        if (r == cudaSuccess) r = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
        for (int c = 0; c < kMaxHostChunks && r == cudaSuccess; ++c) {
            r = cudaEventCreateWithFlags(&h->ev_x[c], cudaEventDisableTiming);
            if (r == cudaSuccess) r = cudaEventCreateWithFlags(&h->ev_k[c], cudaEventDisableTiming);
        }
        return r;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) != cudaSuccess || (e = make_events()) != cudaSuccess) {
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
    for (DevBuf* b : {&h->colptr, &h->isz, &h->sg, &h->inst, &h->gl, &h->gu, &h->fpart, &h->jtmpl, &h->bflag, &h->desc, &h->lidx, &h->sjl, &h->sx, &h->sf_, &h->sgv,
                      &h->sjac, &h->sgrad, &h->ssum})
        release(*b);
    for (auto& b : h->coll) release(b);
    for (auto& b : h->mesh) release(b);
    for (DevBuf* b : {&h->serr, &h->resmat, &h->sxnew, &h->ssznew, &h->slam, &h->ssig, &h->shess, &h->bev, &h->bcls}) release(*b);
    unload_user_kernels(h);
    cudaStreamDestroy(h->stream);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->ev_entry) cudaEventDestroy(h->ev_entry);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    for (int c = 0; c < kMaxHostChunks; ++c) {
        if (h->ev_x[c]) cudaEventDestroy(h->ev_x[c]);
        if (h->ev_k[c]) cudaEventDestroy(h->ev_k[c]);
    }
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
    h->local_index.clear();
    h->hess_ir.clear();
    h->hess_jc.clear();
    h->have_problem = false;
    h->have_inst = false;
    h->have_bounds = false;
    h->have_mesh = false;
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
        h->smem_isz = std::max(p == 0 ? size_t(0) : h->smem_isz,
                               (cta_doubles(pd, ph, kThreads, CARVE_ISZ) - cta_doubles(pd, ph, kThreads, 0)) * sizeof(double));
        h->smem_generic_exact = std::max(p == 0 ? size_t(0) : h->smem_generic_exact,
                                         cta_doubles(pd, ph, kThreads, CARVE_P) * sizeof(double));
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
    if (h->fast_ok) {
        // every launch variant of the specialised kernels must fit one CTA's shared memory, not only the generic
        // footprint checked above: exact mode adds the 1/sz copy, the template ring and (fused summary) the staged
        // bounds. A problem that does not fit (many obstacle records per instance) runs on the generic kernel instead
        // of failing at its first evaluation.
        int ncp = 0;
        for (int p = 0; p < hp.nphases; ++p) ncp = std::max(ncp, phase_ncons(pd, pd.ph[p]));
        const size_t bounds = 2 * static_cast<size_t>(ncp + 2) * sizeof(double);
        const size_t worst_exact = smem_ex + h->smem_isz + (kCopySlots * kCopyChunk + 2) * sizeof(double) + bounds;
        const size_t worst_fd = smem_fd + bounds;
        if (std::max(worst_exact, worst_fd) > 227 * 1024 - 1024) h->fast_ok = false;
    }
    h->image_ok = h->fast_ok && !h->no_image && (pd.nnz & 1) == 0 && smem_img <= 227 * 1024 - 1024;
    h->rowsn_N = 0;
    if (h->fast_ok && !std::getenv("ECUDA_NO_ROWSN")) {
        h->rowsn_N = pd.ph[0].N;
        for (int p = 1; p < hp.nphases; ++p)
            if (pd.ph[p].N != h->rowsn_N) h->rowsn_N = 0;
    }
    h->rows_fill = true;
    for (int p = 0; p < hp.nphases; ++p) h->rows_fill = h->rows_fill && 8 * pd.ns * pd.ph[p].N >= 7 * kThreads;
    if (std::getenv("ECUDA_ROWS_ALWAYS")) h->rows_fill = true;
    int rc;
    h->um = user_model(desc->model);
    if (h->um) {  // only the row-owner and the generic kernels are compiled for user models
        if (h->no_rows) h->fast_ok = false;
        h->image_ok = false;
        if ((rc = load_user_kernels(h))) return rc;
    } else {
        unload_user_kernels(h);
    }
    if ((rc = ensure(h, h->colptr, sizeof(int32_t) * (pd.nvars + 1) + 16))) return rc;  // + 16: bulk copies read whole 16-byte units
    CU(cudaMemcpy(h->colptr.p, hp.colptr.data(), sizeof(int32_t) * (pd.nvars + 1), cudaMemcpyHostToDevice));
    pd.colptr = static_cast<const int*>(h->colptr.p);
    if ((rc = ensure(h, h->desc, sizeof(uint64_t) * std::max<size_t>(1, hp.tdesc.size())))) return rc;
    CU(cudaMemcpy(h->desc.p, hp.tdesc.data(), sizeof(uint64_t) * hp.tdesc.size(), cudaMemcpyHostToDevice));
    pd.desc = static_cast<const unsigned long long*>(h->desc.p);
    h->persist = std::getenv("ECUDA_PERSIST") != nullptr;
    {
        const char* pf = std::getenv("ECUDA_PREFETCH");
        h->prefetch = pf && pf[0] == '1';
    }
    h->exact_kernel = 0;
    if (const char* ek = std::getenv("ECUDA_EXACT_KERNEL"))
        h->exact_kernel = std::strcmp(ek, "ring") == 0 ? 1 : std::strcmp(ek, "stream") == 0 ? 2 : 0;
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
    // compact form for the fused summary: single phase with at least one path row, uniform defect / path bounds
    h->bounds_compact = false;
    const int B = h->hp.desc.batch, ng = h->pd.ncons, ndef = h->pd.ns * h->pd.ph[0].N, ne = h->pd.ne;
    if (h->pd.nphases == 1 && h->pd.ph[0].npath > 0 && !std::getenv("ECUDA_DENSE_BOUNDS")) {
        if ((rc = ensure(h, h->bcls, sizeof(int)))) return rc;
        int one = 1;
        int* flag = static_cast<int*>(h->bcls.p);
        CU(cudaMemcpyAsync(flag, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        k_bounds_classify<<<h->num_sms * 4, 256, 0, h->stream>>>(static_cast<const double*>(h->gl.p),
                                                                static_cast<const double*>(h->gu.p), B, ng, ndef, ne, flag);
        double pl[2];
        CU(cudaMemcpyAsync(&one, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(&pl[0], static_cast<const double*>(h->gl.p) + ndef + ne, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(&pl[1], static_cast<const double*>(h->gu.p) + ndef + ne, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (one == 1) {
            if ((rc = ensure(h, h->bev, sizeof(double) * B * 2 * (ne + 1)))) return rc;
            k_bounds_compact<<<(B * (ne + 1) + 255) / 256, 256, 0, h->stream>>>(
                static_cast<const double*>(h->gl.p), static_cast<const double*>(h->gu.p), B, ng, ndef, ne,
                static_cast<double*>(h->bev.p));
            h->plo = pl[0];
            h->phi = pl[1];
            h->bounds_compact = true;
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    h->have_bounds = true;
    return ECUDA_OK;
}

// HOST-buffer evaluation: the batch goes through the device in chunks of instances on three streams -- host->device
// copies of x on h->s_h2d, kernels on the caller's stream, device->host copies on h->s_d2h -- so that the upload
// of chunk c+1 and the kernels of chunk c+1 overlap the download of chunk c (the download is > 95 % of the call:
// 106 KB of results per instance against 2.9 KB of input). jac_local != null: exact mode, the full triplet array
// stays in device scratch and only the per-instance triplets (k_gather_local) are downloaded.
static int eval_host(ecuda_ctx* h, const double* x, double* f, double* g, double* jac, double* grad, double* jac_local,
                     int jac_mode, cudaStream_t st) {
    const size_t B = h->hp.desc.batch, nv = h->pd.nvars, ng = h->pd.ncons, nz = h->pd.nnz;
    const size_t nl = jac_local ? h->local_index.size() : 0;
    const bool want_full = jac || jac_local;
    int rc;
    if ((rc = ensure(h, h->sx, sizeof(double) * B * nv))) return rc;
    if (f && (rc = ensure(h, h->sf_, sizeof(double) * B))) return rc;
    if (g && (rc = ensure(h, h->sgv, sizeof(double) * B * ng))) return rc;
    if (want_full && (rc = ensure(h, h->sjac, sizeof(double) * B * nz))) return rc;
    if (jac_local && (rc = ensure(h, h->sjl, sizeof(double) * B * nl))) return rc;
    if (grad && (rc = ensure(h, h->sgrad, sizeof(double) * B * nv))) return rc;
    const size_t out_bytes = 8 * B * ((f ? 1 : 0) + (g ? ng : 0) + (jac ? nz : 0) + nl + (grad ? nv : 0));
    int K = h->host_chunks > 0 ? h->host_chunks : (out_bytes >= (size_t(16) << 20) ? 4 : 1);
    if (static_cast<size_t>(K) > B) K = static_cast<int>(B);
    if (K > kMaxHostChunks) K = kMaxHostChunks;
    double* dx = static_cast<double*>(h->sx.p);
    double* df = f ? static_cast<double*>(h->sf_.p) : nullptr;
    double* dg = g ? static_cast<double*>(h->sgv.p) : nullptr;
    double* dj = want_full ? static_cast<double*>(h->sjac.p) : nullptr;
    double* djl = jac_local ? static_cast<double*>(h->sjl.p) : nullptr;
    double* dgr = grad ? static_cast<double*>(h->sgrad.p) : nullptr;
    // the copy streams start behind whatever the caller has queued on st
    CU(cudaEventRecord(h->ev_entry, st));
    CU(cudaStreamWaitEvent(h->s_h2d, h->ev_entry, 0));
    CU(cudaStreamWaitEvent(h->s_d2h, h->ev_entry, 0));
    for (int c = 0; c < K; ++c) {
        const size_t b0 = B * c / K, nb = B * (c + 1) / K - b0;
        CU(cudaMemcpyAsync(dx + b0 * nv, x + b0 * nv, sizeof(double) * nb * nv, cudaMemcpyHostToDevice, h->s_h2d));
        CU(cudaEventRecord(h->ev_x[c], h->s_h2d));
    }
    for (int c = 0; c < K; ++c) {
        const size_t b0 = B * c / K, nb = B * (c + 1) / K - b0;
        CU(cudaStreamWaitEvent(st, h->ev_x[c], 0));
        EvalIO io{};
        io.inst = static_cast<const double*>(h->inst.p) + b0 * h->pd.inst_stride;
        io.fpart = static_cast<double*>(h->fpart.p) + b0 * h->pd.nphases;
        io.jac_mode = jac_mode;
        io.batch = static_cast<int>(nb);
        io.x = dx + b0 * nv;
        io.f = df ? df + b0 : nullptr;
        io.g = dg ? dg + b0 * ng : nullptr;
        io.jac = dj ? dj + b0 * nz : nullptr;
        io.grad = dgr ? dgr + b0 * nv : nullptr;
        if ((rc = launch_eval(h, io, st))) return rc;
        if (jac_local) {
            const dim3 grid(static_cast<unsigned>((nl + 255) / 256), static_cast<unsigned>(nb));
            k_gather_local<<<grid, 256, 0, st>>>(io.jac, static_cast<const int32_t*>(h->lidx.p), djl + b0 * nl, (int)nz, (int)nl);
            ++h->launches;
        }
        CU(cudaEventRecord(h->ev_k[c], st));
        CU(cudaStreamWaitEvent(h->s_d2h, h->ev_k[c], 0));
        if (f) CU(cudaMemcpyAsync(f + b0, df + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, h->s_d2h));
        if (g) CU(cudaMemcpyAsync(g + b0 * ng, dg + b0 * ng, sizeof(double) * nb * ng, cudaMemcpyDeviceToHost, h->s_d2h));
        if (jac) CU(cudaMemcpyAsync(jac + b0 * nz, dj + b0 * nz, sizeof(double) * nb * nz, cudaMemcpyDeviceToHost, h->s_d2h));
        if (jac_local) CU(cudaMemcpyAsync(jac_local + b0 * nl, djl + b0 * nl, sizeof(double) * nb * nl, cudaMemcpyDeviceToHost, h->s_d2h));
        if (grad) CU(cudaMemcpyAsync(grad + b0 * nv, dgr + b0 * nv, sizeof(double) * nb * nv, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    // the caller's stream continues only after the results have landed; so does the host
    CU(cudaEventRecord(h->ev_done, h->s_d2h));
    CU(cudaStreamWaitEvent(st, h->ev_done, 0));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
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
    if (memkind == ECUDA_MEM_DEVICE) {
        EvalIO io{};
        io.inst = static_cast<const double*>(h->inst.p);
        io.fpart = static_cast<double*>(h->fpart.p);
        io.jac_mode = jac_mode;
        io.batch = h->hp.desc.batch;
        io.x = x;
        io.f = f;
        io.g = g;
        io.jac = jac;
        io.grad = grad;
        return launch_eval(h, io, st);
    }
    return eval_host(h, x, f, g, jac, grad, nullptr, jac_mode, st);
}

int ecuda_eval(ecuda_handle h, const double* x, double* f, double* g, double* jac, int jac_mode, int memkind,
               void* stream) {
    return eval_common(h, x, f, g, jac, nullptr, jac_mode, memkind, stream);
}

static int ensure_compact(ecuda_ctx* h) {
    if (!h->local_index.empty()) return ECUDA_OK;
    build_local_index(h->hp, &h->local_index);
    int rc;
    if ((rc = ensure(h, h->lidx, sizeof(int32_t) * h->local_index.size()))) return rc;
    CU(cudaMemcpy(h->lidx.p, h->local_index.data(), sizeof(int32_t) * h->local_index.size(), cudaMemcpyHostToDevice));
    return ECUDA_OK;
}

int ecuda_get_compact_structure(ecuda_handle h, int32_t* nlocal, int32_t* local_index, double* shared_vals) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    std::vector<int32_t> tmp;
    const std::vector<int32_t>* li = &h->local_index;
    if (li->empty()) {
        build_local_index(h->hp, &tmp);
        li = &tmp;
    }
    if (nlocal) *nlocal = static_cast<int32_t>(li->size());
    if (local_index) std::memcpy(local_index, li->data(), sizeof(int32_t) * li->size());
    if (shared_vals) {
        std::vector<double> isz(h->hp.dims.nvars), tmpl;
        for (int c = 0; c < h->hp.dims.nvars; ++c) isz[c] = 1.0 / h->h_sz[c];
        build_jac_template(h->hp, isz.data(), h->h_sg.data(), &tmpl);
        std::memcpy(shared_vals, tmpl.data(), sizeof(double) * h->hp.dims.nnz);
    }
    return ECUDA_OK;
}

int ecuda_eval_compact(ecuda_handle h, const double* x, double* f, double* g, double* jac_local, int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!jac_local) return fail(h, ECUDA_ERR_ARG, "jac_local is required (use ecuda_eval for values only)");
    if (h->hp.desc.batch > 65535) return fail(h, ECUDA_ERR_ARG, "ecuda_eval_compact: batch > 65535");
    if (memkind != ECUDA_MEM_HOST && memkind != ECUDA_MEM_DEVICE) return fail(h, ECUDA_ERR_ARG, "bad memkind");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_compact(h))) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const size_t B = h->hp.desc.batch, nz = h->pd.nnz, nl = h->local_index.size();
    // the full triplet array is handle-owned scratch on the device; only the per-instance part leaves it
    if ((rc = ensure(h, h->sjac, sizeof(double) * B * nz))) return rc;
    double* full = static_cast<double*>(h->sjac.p);
    const dim3 grid(static_cast<unsigned>((nl + 255) / 256), static_cast<unsigned>(B));
    if (memkind == ECUDA_MEM_DEVICE) {
        if ((rc = eval_common(h, x, f, g, full, nullptr, ECUDA_JAC_EXACT, ECUDA_MEM_DEVICE, stream))) return rc;
        k_gather_local<<<grid, 256, 0, st>>>(full, static_cast<const int32_t*>(h->lidx.p), jac_local, (int)nz, (int)nl);
        ++h->launches;
        CU(cudaGetLastError());
        return ECUDA_OK;
    }
    if (!h->have_inst) return fail(h, ECUDA_ERR_STATE, "upload_instances has not been called");
    if (!x) return fail(h, ECUDA_ERR_ARG, "null decision vector");
    return eval_host(h, x, f, g, nullptr, nullptr, jac_local, ECUDA_JAC_EXACT, st);
}

int ecuda_splice_jacobian(const double* shared_vals, const int32_t* local_index, int32_t nnz, int32_t nlocal,
                          const double* jac_local, int32_t batch, double* jac_full) {
    if (!shared_vals || !local_index || !jac_local || !jac_full || nnz < 0 || nlocal < 0 || nlocal > nnz || batch < 0)
        return ECUDA_ERR_ARG;
    for (int32_t i = 0; i < nlocal; ++i)
        if (local_index[i] < 0 || local_index[i] >= nnz) return ECUDA_ERR_ARG;
    for (int32_t b = 0; b < batch; ++b) {
        double* dst = jac_full + static_cast<size_t>(b) * nnz;
        const double* src = jac_local + static_cast<size_t>(b) * nlocal;
        std::memcpy(dst, shared_vals, sizeof(double) * nnz);
        for (int32_t i = 0; i < nlocal; ++i) dst[local_index[i]] = src[i];
    }
    return ECUDA_OK;
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
    if (h->bounds_compact) {
        io.bev = static_cast<const double*>(h->bev.p);
        io.plo = h->plo;
        io.phi = h->phi;
    }
    for (int r = 0; r < nranks; ++r) {
        if (!peer_out[r] || (reinterpret_cast<uintptr_t>(peer_out[r]) & 15)) return fail(h, ECUDA_ERR_ARG, "peer buffer null or not 16-byte aligned");
        io.peer[r] = static_cast<double*>(peer_out[r]);
    }
    return launch_eval(h, io, st);
}

}  // extern "C"

// ---- mesh refinement support ---------------------------------------------------------------------------------
template <int M>
static int launch_ode_error_t(ecuda_ctx* h, const EvalIO& io, const MeshDev& mesh, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    if (h->smem_bytes > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < h->smem_bytes) {
            CU(cudaFuncSetAttribute(k_ode_error<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
            cur = h->smem_bytes;
        }
    }
    k_ode_error<M><<<io.batch * h->pd.nphases, kThreads, h->smem_bytes, st>>>(h->pd, io, mesh);
    ++h->launches;
    return ECUDA_OK;
}

template <int M>
static int launch_hess_t(ecuda_ctx* h, const EvalIO& io, const HessIO& hio, cudaStream_t st) {
    static std::mutex mu;
    static size_t configured[64] = {0};
    if (h->smem_bytes > 48 * 1024) {
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = configured[h->device & 63];
        if (cur < h->smem_bytes) {
            CU(cudaFuncSetAttribute(k_hess<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
            cur = h->smem_bytes;
        }
    }
    k_hess<M><<<io.batch * h->pd.nphases, kThreads, h->smem_bytes, st>>>(h->pd, io, hio);
    ++h->launches;
    return ECUDA_OK;
}

static int ensure_mesh(ecuda_ctx* h) {
    if (h->have_mesh) return ECUDA_OK;
    for (int p = 0; p < h->hp.nphases; ++p) {
        MeshHost mh;
        build_error_mesh(h->hp.col[p], &mh);
        const size_t nE = mh.E.size(), nq = mh.wq.size();
        int rc;
        if ((rc = ensure(h, h->mesh[p], sizeof(double) * (2 * nE + 2 * nq)))) return rc;
        double* base = static_cast<double*>(h->mesh[p].p);
        CU(cudaMemcpy(base, mh.E.data(), sizeof(double) * nE, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(base + nE, mh.dE.data(), sizeof(double) * nE, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(base + 2 * nE, mh.wq.data(), sizeof(double) * nq, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(base + 2 * nE + nq, mh.tq.data(), sizeof(double) * nq, cudaMemcpyHostToDevice));
    }
    h->have_mesh = true;
    return ECUDA_OK;
}

extern "C" {

int ecuda_ode_error(ecuda_handle h, const double* x, double* err, int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!x || !err) return fail(h, ECUDA_ERR_ARG, "x and err are required");
    if (memkind != ECUDA_MEM_HOST && memkind != ECUDA_MEM_DEVICE) return fail(h, ECUDA_ERR_ARG, "bad memkind");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    int rc;
    if ((rc = ensure_mesh(h))) return rc;
    const size_t B = h->hp.desc.batch, nv = h->pd.nvars;
    MeshDev mesh{};
    int nint = 0;
    for (int p = 0; p < h->hp.nphases; ++p) {
        const size_t N = h->hp.N[p], nE = (N - 1) * ECUDA_MESH_Q * N, nq = (N - 1) * ECUDA_MESH_Q;
        const double* base = static_cast<const double*>(h->mesh[p].p);
        mesh.E[p] = base;
        mesh.dE[p] = base + nE;
        mesh.wq[p] = base + 2 * nE;
        mesh.tq[p] = base + 2 * nE + nq;
        mesh.eoff[p] = nint;
        nint += static_cast<int>(N) - 1;
    }
    mesh.nint = nint;
    EvalIO io{};
    io.inst = static_cast<const double*>(h->inst.p);
    io.batch = static_cast<int>(B);
    io.x = x;
    mesh.out = err;
    if (memkind == ECUDA_MEM_HOST) {
        if ((rc = ensure(h, h->sx, sizeof(double) * B * nv))) return rc;
        if ((rc = ensure(h, h->serr, sizeof(double) * B * nint))) return rc;
        CU(cudaMemcpyAsync(h->sx.p, x, sizeof(double) * B * nv, cudaMemcpyHostToDevice, st));
        io.x = static_cast<const double*>(h->sx.p);
        mesh.out = static_cast<double*>(h->serr.p);
    }
    if (h->um) {
        rc = launch_user(h, UserImage::ODE_ERROR, dim3(io.batch * h->pd.nphases), kThreads, h->smem_bytes, io, st, &mesh);
    } else {
        switch (h->pd.model) {
            case ECUDA_MODEL_SI2D: rc = launch_ode_error_t<ECUDA_MODEL_SI2D>(h, io, mesh, st); break;
            case ECUDA_MODEL_PM3D: rc = launch_ode_error_t<ECUDA_MODEL_PM3D>(h, io, mesh, st); break;
            default: rc = launch_ode_error_t<ECUDA_MODEL_FW6>(h, io, mesh, st); break;
        }
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    if (memkind == ECUDA_MEM_HOST) {
        CU(cudaMemcpyAsync(err, mesh.out, sizeof(double) * B * nint, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ECUDA_OK;
}

int ecuda_get_hess_structure(ecuda_handle h, int32_t* nnz_h, int32_t* iRow, int32_t* jCol) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (h->hess_ir.empty()) build_hess_structure(h->hp, &h->hess_ir, &h->hess_jc);  // once per problem (IPOPT asks every iteration)
    const std::vector<int32_t>&ir = h->hess_ir, &jc = h->hess_jc;
    if (nnz_h) *nnz_h = static_cast<int32_t>(ir.size());
    const int base = h->hp.desc.index_base;
    for (size_t e = 0; e < ir.size(); ++e) {
        if (iRow) iRow[e] = ir[e] + base;
        if (jCol) jCol[e] = jc[e] + base;
    }
    return ECUDA_OK;
}

int ecuda_eval_hess(ecuda_handle h, const double* x, const double* sigma, double sigma0, const double* lambda,
                    double* vals, int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!h->have_inst) return fail(h, ECUDA_ERR_STATE, "upload_instances has not been called");
    if (!x || !lambda || !vals) return fail(h, ECUDA_ERR_ARG, "x, lambda and vals are required");
    if (memkind != ECUDA_MEM_HOST && memkind != ECUDA_MEM_DEVICE) return fail(h, ECUDA_ERR_ARG, "bad memkind");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const size_t B = h->hp.desc.batch, nv = h->pd.nvars, ng = h->pd.ncons;
    HessIO hio{};
    int nnz_h = 0;
    for (int p = 0; p < h->hp.nphases; ++p) {
        hio.hoff[p] = nnz_h;
        nnz_h += hess_phase_nnz(h->hp.ns, h->hp.nc, h->hp.N[p]);
    }
    hio.nnz_h = nnz_h;
    hio.sigma0 = sigma0;
    hio.sigma = sigma;
    hio.lambda = lambda;
    hio.vals = vals;
    EvalIO io{};
    io.inst = static_cast<const double*>(h->inst.p);
    io.batch = static_cast<int>(B);
    io.x = x;
    int rc;
    if (memkind == ECUDA_MEM_HOST) {
        if ((rc = ensure(h, h->sx, sizeof(double) * B * nv))) return rc;
        if ((rc = ensure(h, h->slam, sizeof(double) * B * ng))) return rc;
        if ((rc = ensure(h, h->shess, sizeof(double) * B * nnz_h))) return rc;
        CU(cudaMemcpyAsync(h->sx.p, x, sizeof(double) * B * nv, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->slam.p, lambda, sizeof(double) * B * ng, cudaMemcpyHostToDevice, st));
        io.x = static_cast<const double*>(h->sx.p);
        hio.lambda = static_cast<const double*>(h->slam.p);
        hio.vals = static_cast<double*>(h->shess.p);
        if (sigma) {
            if ((rc = ensure(h, h->ssig, sizeof(double) * B))) return rc;
            CU(cudaMemcpyAsync(h->ssig.p, sigma, sizeof(double) * B, cudaMemcpyHostToDevice, st));
            hio.sigma = static_cast<const double*>(h->ssig.p);
        }
    }
    if (h->um) {
        rc = launch_user(h, UserImage::HESS, dim3(io.batch * h->pd.nphases), kThreads, h->smem_bytes, io, st, &hio);
    } else {
        switch (h->pd.model) {
            case ECUDA_MODEL_SI2D: rc = launch_hess_t<ECUDA_MODEL_SI2D>(h, io, hio, st); break;
            case ECUDA_MODEL_PM3D: rc = launch_hess_t<ECUDA_MODEL_PM3D>(h, io, hio, st); break;
            default: rc = launch_hess_t<ECUDA_MODEL_FW6>(h, io, hio, st); break;
        }
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    if (memkind == ECUDA_MEM_HOST) {
        CU(cudaMemcpyAsync(vals, hio.vals, sizeof(double) * B * nnz_h, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ECUDA_OK;
}

int ecuda_resample(ecuda_handle h, const double* x, const int32_t* nnodes_new, const double* sz_new, double* x_new,
                   int memkind, void* stream) {
    if (!h) return ECUDA_ERR_ARG;
    if (!h->have_problem) return fail(h, ECUDA_ERR_STATE, "set_problem has not succeeded");
    if (!x || !x_new || !nnodes_new) return fail(h, ECUDA_ERR_ARG, "x, nnodes_new and x_new are required");
    if (memkind != ECUDA_MEM_HOST && memkind != ECUDA_MEM_DEVICE) return fail(h, ECUDA_ERR_ARG, "bad memkind");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    const HostProblem& hp = h->hp;
    // interpolation matrices of all phases, back to back, and the layout of the new decision vector
    ResampleDev rd{};
    std::vector<double> mats;
    int zo = 0;
    for (int p = 0; p < hp.nphases; ++p) {
        Collocation to;
        std::string err;
        if (!build_collocation(hp.desc.collocation, nnodes_new[p], &to, &err)) return fail(h, ECUDA_ERR_ARG, err);
        std::vector<double> R;
        build_resample(hp.col[p], to, &R);
        rd.moff[p] = static_cast<int>(mats.size());
        rd.Nnew[p] = nnodes_new[p];
        rd.zoff_new[p] = zo;
        zo += (hp.ns + hp.nc) * nnodes_new[p] + 2;
        mats.insert(mats.end(), R.begin(), R.end());
    }
    rd.nvars_new = zo;
    const size_t B = hp.desc.batch, nv = h->pd.nvars;
    int rc;
    if ((rc = ensure(h, h->resmat, sizeof(double) * mats.size()))) return rc;
    CU(cudaMemcpyAsync(h->resmat.p, mats.data(), sizeof(double) * mats.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // mats is a local
    rd.R = static_cast<const double*>(h->resmat.p);
    rd.x = x;
    rd.x_new = x_new;
    rd.sz_new = sz_new;
    if (memkind == ECUDA_MEM_HOST) {
        if ((rc = ensure(h, h->sx, sizeof(double) * B * nv))) return rc;
        if ((rc = ensure(h, h->sxnew, sizeof(double) * B * zo))) return rc;
        CU(cudaMemcpyAsync(h->sx.p, x, sizeof(double) * B * nv, cudaMemcpyHostToDevice, st));
        rd.x = static_cast<const double*>(h->sx.p);
        rd.x_new = static_cast<double*>(h->sxnew.p);
        if (sz_new) {
            if ((rc = ensure(h, h->ssznew, sizeof(double) * zo))) return rc;
            CU(cudaMemcpyAsync(h->ssznew.p, sz_new, sizeof(double) * zo, cudaMemcpyHostToDevice, st));
            rd.sz_new = static_cast<const double*>(h->ssznew.p);
        }
    }
    k_resample<<<static_cast<unsigned>(B) * hp.nphases, 256, 0, st>>>(h->pd, rd);
    ++h->launches;
    CU(cudaGetLastError());
    if (memkind == ECUDA_MEM_HOST) {
        CU(cudaMemcpyAsync(x_new, rd.x_new, sizeof(double) * B * zo, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ECUDA_OK;
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
    if (!h->bflag.p) {  // status words {timed out, step, late rank + 1}: zeroed once, when they are allocated
        if ((rc = ensure(h, h->bflag, 4 * sizeof(unsigned long long)))) return rc;
        CU(cudaMemsetAsync(h->bflag.p, 0, 4 * sizeof(unsigned long long), st));
    }
    unsigned long long timeout_ms = 10000;
    if (const char* e = std::getenv("ECUDA_PEER_TIMEOUT_MS")) {
        const long long v = std::atoll(e);
        if (v > 0) timeout_ms = static_cast<unsigned long long>(v);
    }
    k_peer_barrier<<<1, 32, 0, st>>>(pf, nranks, rank, static_cast<unsigned long long>(step), timeout_ms * 1000000ull,
                                     static_cast<unsigned long long*>(h->bflag.p));
    h->barrier_used = true;
    ++h->launches;
    CU(cudaGetLastError());
    return ECUDA_OK;
}

// reads the sticky status of the peer barriers issued through this handle (after the work on `stream` has finished)
static int read_barrier_status(ecuda_ctx* h, cudaStream_t st, unsigned long long out[3], bool reset) {
    out[0] = out[1] = out[2] = 0;
    if (!h->barrier_used || !h->bflag.p) return ECUDA_OK;
    CU(cudaMemcpyAsync(out, h->bflag.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (reset && out[0]) {
        CU(cudaMemsetAsync(h->bflag.p, 0, 4 * sizeof(unsigned long long), st));
        CU(cudaStreamSynchronize(st));
    }
    return ECUDA_OK;
}

int ecuda_peer_barrier_status(ecuda_handle h, void* stream, int32_t* timed_out, uint64_t* step, int32_t* late_rank, int reset) {
    if (!h) return ECUDA_ERR_ARG;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    CU(cudaStreamSynchronize(st));
    unsigned long long w[3];
    int rc = read_barrier_status(h, st, w, reset != 0);
    if (rc) return rc;
    if (timed_out) *timed_out = w[0] ? 1 : 0;
    if (step) *step = w[1];
    if (late_rank) *late_rank = w[0] ? static_cast<int32_t>(w[2]) - 1 : -1;
    return ECUDA_OK;
}

int ecuda_sync(ecuda_handle h) {
    if (!h) return ECUDA_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    unsigned long long w[3];
    int rc = read_barrier_status(h, h->stream, w, false);
    if (rc) return rc;
    if (w[0])
        return fail(h, ECUDA_ERR_PEER, "ecuda_peer_barrier timed out at step " + std::to_string(w[1]) + " waiting for rank " +
                                           std::to_string(static_cast<long long>(w[2]) - 1) +
                                           ": the gathered rows of that step are incomplete (read and clear with "
                                           "ecuda_peer_barrier_status)");
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
    if (const UserModel* um = user_model(model)) {
        user_model_eval(*um, x, u, t, f_out, cost_out);
        return ECUDA_OK;
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
        rows[q] = rec == 6 ? edge_row(r, x, y) : cylinder_row(r, x, y);
    }
    for (int i = 0; i < desc->ntracks; ++i)
        rows[nstat + i] = track_row(inst + hp.track_off + i * hp.dims.track_size, desc->nwaypoints, x, y, t);
    if (const UserModel* um = desc->model >= ECUDA_MODEL_USER_BASE ? user_model(desc->model) : nullptr)
        user_model_rows(*um, x, y, t, rows + nstat + desc->ntracks);  // traced path rows
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
int ecuda_ipopt_eval_h(ecuda_handle h, int n, const double* x, int new_x, double obj_factor, int m, const double* lambda,
                       int new_lambda, int nele_hess, int32_t* iRow, int32_t* jCol, double* values) {
    (void)new_x;
    (void)new_lambda;
    int rc = ipopt_guard(h, n);
    if (rc) return rc;
    int32_t nnz_h = 0;
    if ((rc = ecuda_get_hess_structure(h, &nnz_h, nullptr, nullptr))) return rc;
    if (m != h->pd.ncons || nele_hess != nnz_h) return fail(h, ECUDA_ERR_ARG, "m / nele_hess mismatch");
    if (!values) return ecuda_get_hess_structure(h, nullptr, iRow, jCol);
    return ecuda_eval_hess(h, x, nullptr, obj_factor, lambda, values, ECUDA_MEM_HOST, nullptr);
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
#endif  // ECUDA_TU_ROWSN
