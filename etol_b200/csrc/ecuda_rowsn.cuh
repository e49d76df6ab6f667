// ecuda_rowsn.cuh -- row-owner evaluation with the node count N as a COMPILE-TIME constant.
//
// Same arithmetic as ecuda_rows.cuh / ecuda_phases.cuh (every value is produced by the same FP64 operation
// sequence, so results are bit-identical to the generic kernels and to the oracle); what changes is what the
// instruction stream spends its issue slots on. The round-1 kernel (k_eval_rows<M,NB,FD>, N a kernel
// parameter) executed 3000 warp instructions per warp of which only 1068 were FP64: every D[k][l] load
// needed a 64-bit multiply-add for its address, every triplet two compares, two selects and a four-instruction
// 64-bit address computation, and the per-column finite-difference data came from five separate shared arrays.
// Here
//   * N, the block count and every offset derived from them are constants: D[k][l], X(l,i) and the
//     per-column records are addressed as  per-thread base + immediate;
//   * the finite-difference data of a column live in ONE 32-byte shared-memory record {z+d, z-d, 1/(2d),
//     first triplet}, fetched with two 128-bit loads;
//   * triplet addresses are  (jac + k) + 8 * (record.cp + select) : one compare, one select, one add, one
//     wide multiply-add.
//   * STORES: the D-coupled triplets (74 % of the Jacobian) are not stored from registers. A store-only kernel with
//     the row-owner pattern (every warp instruction writes one 256-byte piece, consecutive instructions advance by a
//     column) reaches 2.2 TB/s on a B200; 8 KB bulk stores from shared memory reach 5.9 TB/s (scripts/wroof.cu,
//     profiles/r2). So the threads write their D-coupled values into a shared-memory ring -- one buffer per group of
//     ECUDA_RN_GROUP nodes, whose state columns are one contiguous range of the (col,row)-sorted triplet array --
//     and every finished group leaves with one cp.async.bulk shared -> global (TMA) while the next one is computed.
//     The few node-local triplets inside those ranges are written afterwards, into lines the bulk stores have
//     just put into L2.
// Reference counterparts: PSOPT's defect assembly and index-set finite differences entered at
// src/ePSOPT/ePSOPT.cpp:84 (mode chosen at :64), callbacks src/ePSOPT/ePSOPT.cpp:186-306.
#ifndef ECUDA_ROWSN_CUH_
#define ECUDA_ROWSN_CUH_

#include "ecuda_rows.cuh"

namespace ecuda {

// finite-difference data of one decision variable (one Jacobian column), staged per instance
struct alignas(16) FdRec {
    double xp;  // (z~ + delta) / sz
    double xm;  // (z~ - delta) / sz
    double ri;  // 1 / (2 delta)
    int cp;     // first triplet of the column
    int pad_;
};
// the two halves of a record as 128-bit shared-memory loads
struct FdVals {
    double xp, xm, ri;
    unsigned cp;
};
ECUDA_HD FdVals rn_load(const FdRec* r) {
    FdVals v;
#if defined(__CUDA_ARCH__)
    const double2 a = *reinterpret_cast<const double2*>(r);
    const double2 c = *(reinterpret_cast<const double2*>(r) + 1);
    v.xp = a.x;
    v.xm = a.y;
    v.ri = c.x;
    v.cp = static_cast<unsigned>(__double2loint(c.y));
#else
    v.xp = r->xp;
    v.xm = r->xm;
    v.ri = r->ri;
    v.cp = static_cast<unsigned>(r->cp);
#endif
    return v;
}
// byte a (0..3) of w
ECUDA_HD unsigned rn_byte(unsigned w, int a) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, 0x4440u + static_cast<unsigned>(a));
#else
    return (w >> (8 * a)) & 0xffu;
#endif
}

// exact mode: what a triplet of a column needs -- 1 / sz of the variable and the column's first triplet
struct alignas(16) ExRec {
    double isz;
    int cp;
    int pad_;
};
struct ExVals {
    double isz;
    unsigned cp;
};
ECUDA_HD ExVals rn_load(const ExRec* r) {
    ExVals v;
#if defined(__CUDA_ARCH__)
    const double2 a = *reinterpret_cast<const double2*>(r);
    v.isz = a.x;
    v.cp = static_cast<unsigned>(__double2loint(a.y));
#else
    v.isz = r->isz;
    v.cp = static_cast<unsigned>(r->cp);
#endif
    return v;
}

// shared memory of one CTA of the N-specialised kernels
struct RnMem {
    double* inst;  // [inst_stride] obstacle / track records (bulk-copied)
    double* z;     // [nv]          unscaled variables of the phase
    FdRec* rec;    // [nv]          FD mode
    ExRec* erec;   // [nv]          exact mode
    // collocation data of the phase, copied into shared memory while the decision vector is on its way (FD mode):
    // after the staging barrier the threads read nothing from global memory -- under the kernel's own write traffic
    // even an L1 hit queues behind the stores, and a miss costs thousands of cycles
    const double* dt;   // [N*N] D^T (dt[l*N + k] = D[k][l]); exact mode: the global array
    const double* tau;  // [N]
    const double* w;    // [N]
    // ECUDA_RN_TMAZ: raw copies of x[b], 1/sz and the column pointers (bulk copies; 16-byte aligned, padded)
    double* rawz;
    double* rawis;
    int* rawcp;
};
// bytes of the column-pointer bulk copy (multiple of 16)
ECUDA_HD unsigned rn_cp_bytes(int nv) { return (static_cast<unsigned>(nv) * 4u + 15u) & ~15u; }

template <int M>
ECUDA_HD int rn_nv(const ProbDev& pb, int N) { return (Model<M>::NS + pb.nc) * N + 2; }

// ECUDA_RN_DSMEM: the one-shot FD kernel copies D^T, tau and w into shared memory per CTA. Measured on C2: 0.165 ms
// with the copy, 0.159 ms reading them through L1 (the copy costs more than the shorter load latency saves), so it is
// off; the persistent kernel, which copies once per CTA lifetime, always has them in shared memory.
// ECUDA_RN_TMAZ: the one-shot FD kernel fetches the decision vector, 1/sz and the column pointers with TMA bulk copies
// (the mbarrier that already brings the instance records) instead of per-thread loads, and stages from shared memory.
#ifndef ECUDA_RN_TMAZ
#define ECUDA_RN_TMAZ 0
#endif
// ECUDA_RN_NOSCATTER (experiment, WRONG RESULTS): the finite-difference kernel computes the node-local triplets, the
// g values and the item rows but does not store them -- an upper bound for what removing the scattered 8-byte stores
// can buy (scripts/wroof2.cu: under that store pattern an L2 hit costs 3300 cycles instead of 800).
#ifndef ECUDA_RN_NOSCATTER
#define ECUDA_RN_NOSCATTER 0
#endif
#if ECUDA_RN_NOSCATTER && defined(__CUDA_ARCH__)
#define RN_LOCAL_STORE(p, v)                         \
    do {                                             \
        const double v_ = (v);                       \
        if (v_ == 1.2345678e-300) __stcs((p), v_);   \
    } while (0)
#else
#define RN_LOCAL_STORE(p, v) ECUDA_STREAM_STORE(p, v)
#endif
// ECUDA_RN_FD_RING: finite differences through the shared-memory store ring as well (measured slower, see k_rows_n)
#ifndef ECUDA_RN_FD_RING
#define ECUDA_RN_FD_RING 0
#endif
#ifndef ECUDA_RN_DSMEM
#define ECUDA_RN_DSMEM 0
#endif
#ifndef ECUDA_RN_OBJWARP
#define ECUDA_RN_OBJWARP 0
#endif
// ECUDA_RN_INTERLEAVE: the node-local triplets of a defect row as separate pieces that run between the D-coupled node
// groups instead of as one function after the last one. Measured on C2: 0.167 ms as pieces (interleaved or all at
// the end: each piece reloads the node's variables), 0.159 ms as one function, so it is off.
#ifndef ECUDA_RN_INTERLEAVE
#define ECUDA_RN_INTERLEAVE 0
#endif
// store ring: kRnBufs buffers, each holds the triplets of the state columns of kRnGroup consecutive nodes
constexpr int kRnGroup = 4;  // nodes per group (half a summation block)
constexpr int kRnBufs = 3;   // one barrier per group needs three buffers (see k_rows_n)
// doubles of one ring buffer: upper bound of a group's triplet range (+ 2: parity pad, even size)
template <int M>
ECUDA_HD size_t rn_group_cap(const ProbDev& pb, const PhaseDev& ph, int N) {
    constexpr int NS = Model<M>::NS;
    int per_node = NS * (N - 1) + 2 * ph.npath + 2 * NS;  // D-coupled + path rows (states 0, 1) + event + linkage
    for (int j = 0; j < NS; ++j) per_node += pb.xcnt[j];
    const size_t n = static_cast<size_t>(kRnGroup) * per_node + 2;
    return n + (n & 1);
}

// doubles of shared memory, without the exact-mode ring and the fused-summary bounds
template <int M>
ECUDA_HD size_t rn_doubles(const ProbDev& pb, int N, bool fd) {
    const size_t nv = static_cast<size_t>(rn_nv<M>(pb, N)), nve = nv + (nv & 1);
    size_t n = static_cast<size_t>(pb.inst_stride) + nve;
    n += fd ? 4 * nv : 2 * nv;
    if (fd && ECUDA_RN_DSMEM) n += static_cast<size_t>(N) * N + 2 * static_cast<size_t>(N + (N & 1));
    if (fd && ECUDA_RN_TMAZ) n += 2 * nve + rn_cp_bytes(static_cast<int>(nv)) / 8;
    return n + (n & 1);
}

template <int M>
ECUDA_HD void rn_carve(RnMem& m, double* base, const ProbDev& pb, int N, bool fd) {
    const size_t nv = static_cast<size_t>(rn_nv<M>(pb, N)), nve = nv + (nv & 1);
    m.inst = base;  // first: 16-byte aligned destination of the bulk copy (inst_stride is even)
    base += pb.inst_stride;
    m.z = base;
    base += nve;
    m.rec = nullptr;
    m.erec = nullptr;
    m.dt = m.tau = m.w = nullptr;  // exact mode: rn_stage points them at the global arrays
    if (fd) {
        m.rec = reinterpret_cast<FdRec*>(base);
        base += 4 * nv;
        if (ECUDA_RN_TMAZ) {
            m.rawz = base;
            base += nve;
            m.rawis = base;
            base += nve;
            m.rawcp = reinterpret_cast<int*>(base);
            base += rn_cp_bytes(static_cast<int>(nv)) / 8;
        }
        m.dt = base;
        base += static_cast<size_t>(N) * N;
        m.tau = base;
        base += N + (N & 1);
        m.w = base;
    } else {
        m.erec = reinterpret_cast<ExRec*>(base);
    }
}

// ---- stage: same arithmetic as stage_vars ----------------------------------------------------------------------
template <int M, int N, bool FD>
ECUDA_HD void rn_stage(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, RnMem& m, int b, int tid, int nthr) {
    const int nv = rn_nv<M>(pb, N);
    const double* xs = io.x + static_cast<size_t>(b) * pb.nvars + ph.zoff;
    const double* is = pb.isz + ph.zoff;
    const int* cpg = pb.colptr + ph.zoff;
    // Every global load of the thread is issued before the first dependent store: the decision vector comes from HBM
    // (thousands of cycles under the kernel's own write traffic), the rest from L2 -- one wait for all of them.
    double zt[2], s[2];
    int cp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int c = tid + u * nthr;
        if (c < nv) {
            if (FD && ECUDA_RN_TMAZ) {
                zt[u] = m.rawz[c];
                s[u] = m.rawis[c];
                cp[u] = m.rawcp[c];
            } else {
                zt[u] = ECUDA_LDG(xs + c);
                s[u] = ECUDA_LDG(is + c);
                cp[u] = ECUDA_LDG(cpg + c);
            }
        }
    }
    constexpr bool DSM = FD && ECUDA_RN_DSMEM;
    constexpr int ND = DSM ? (N * N + 255) / 256 : 1;
    double dv[ND], tv = 0.0, wv = 0.0;
    if (DSM) {
#pragma unroll
        for (int u = 0; u < ND; ++u)
            if (tid + u * nthr < N * N) dv[u] = ECUDA_LDG(ph.Dt + tid + u * nthr);
        if (tid < N) {
            tv = ECUDA_LDG(ph.tau + tid);
            wv = ECUDA_LDG(ph.w + tid);
        }
    } else {
        m.dt = ph.Dt;
        m.tau = ph.tau;
        m.w = ph.w;
    }
    for (int c0 = tid; c0 < nv; c0 += 2 * nthr) {
        if (c0 != tid) {  // further batches (more than 2 * nthr variables)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = c0 + u * nthr;
                if (c < nv) {
                    if (FD && ECUDA_RN_TMAZ) {
                        zt[u] = m.rawz[c];
                        s[u] = m.rawis[c];
                        cp[u] = m.rawcp[c];
                    } else {
                        zt[u] = ECUDA_LDG(xs + c);
                        s[u] = ECUDA_LDG(is + c);
                        cp[u] = ECUDA_LDG(cpg + c);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c = c0 + u * nthr;
            if (c < nv) {
                m.z[c] = zt[u] * s[u];
                if (FD) {
                    const double delta = ECUDA_SQRT_EPS * (1.0 + fabs(zt[u]));
                    FdRec r;
                    r.xp = (zt[u] + delta) * s[u];
                    r.xm = (zt[u] - delta) * s[u];
                    r.ri = 1.0 / (2.0 * delta);
                    r.cp = cp[u];
                    r.pad_ = 0;
                    m.rec[c] = r;
                } else {
                    ExRec r;
                    r.isz = s[u];
                    r.cp = cp[u];
                    r.pad_ = 0;
                    m.erec[c] = r;
                }
            }
        }
    }
    if (DSM) {  // collocation data of the phase into shared memory
        double* dt = const_cast<double*>(m.dt);
#pragma unroll
        for (int u = 0; u < ND; ++u)
            if (tid + u * nthr < N * N) dt[tid + u * nthr] = dv[u];
        if (tid < N) {
            const_cast<double*>(m.tau)[tid] = tv;
            const_cast<double*>(m.w)[tid] = wv;
        }
    }
}

// ---- D X: block sums and total (canonical blocked order, see dot_row) ----------------------------------------
template <int NS, int N>
ECUDA_HD double rn_dot(const double* __restrict__ Dtk, const double* __restrict__ Xi, double (&P)[(N + 7) / 8]) {
    constexpr int BL = ECUDA_DOT_BLOCK, NB = (N + BL - 1) / BL;
    double total = 0.0;
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
        double p = 0.0;
#pragma unroll
        for (int a = 0; a < BL; ++a)
            if (bi * BL + a < N) p = fma(Dtk[(bi * BL + a) * N], Xi[(bi * BL + a) * NS], p);
        P[bi] = p;
        total = (bi == 0) ? p : total + p;
    }
    return total;
}

// perturbed dots of the row's own diagonal column (models whose f_i reads x_i): sequence of fast_diag
template <int NS, int N>
ECUDA_HD void rn_diag(const double* __restrict__ Dtk, const double* __restrict__ Xi, double xpk, double xmk, int k,
                      const double (&P)[(N + 7) / 8], double& dp, double& dm) {
    constexpr int BL = ECUDA_DOT_BLOCK, NB = (N + BL - 1) / BL;
    const int bk = k / BL, l0 = bk * BL, krel = k - l0;
    double q = 0.0, sp = 0.0, sm = 0.0;
#pragma unroll
    for (int a = 0; a < BL; ++a) {
        if (l0 + a < N) {
            const double di = Dtk[(l0 + a) * N];
            const double xi = Xi[(l0 + a) * NS];
            if (a < krel) {
                q = fma(di, xi, q);
            } else if (a == krel) {
                sp = fma(di, xpk, q);
                sm = fma(di, xmk, q);
            } else {
                sp = fma(di, xi, sp);
                sm = fma(di, xi, sm);
            }
        }
    }
    double pre = 0.0;
#pragma unroll
    for (int t = 0; t < NB; ++t)
        if (t < bk) pre = pre + P[t];
    double tp = pre + sp, tm = pre + sm;
#pragma unroll
    for (int t = 0; t < NB; ++t)
        if (t > bk) {
            tp = tp + P[t];
            tm = tm + P[t];
        }
    dp = tp;
    dm = tm;
}

// ---- per-thread state of a defect row between the phases of the kernel (registers) ----------------------------------
// Triplet of row k in column X(l,i): rows k < l sit at position k of the column; rows k > l come after its
// node-local block, dlt = (number of node-local defect rows of column X(.,i)) - 1 places further; l == k is the
// row's own diagonal triplet at kdo = its rank among the node-local rows (final for DIAG_FREE models, otherwise
// overwritten afterwards). For one group of kRnGroup nodes these offsets are one byte per node: `all` (dlt in every
// byte) for the groups below the row's own, `mix` inside it, 0 above.
template <int N>
struct RnRow {
    double P[(N + 7) / 8];  // block sums of (D X)[k][i]
    double d[8], xv[8];     // FD: D[k][l] and X(l,i) of the current summation block
    double pre, q;          // FD: sum of the earlier block sums, unperturbed in-block prefix
    double sgr, hfv;        // row scale, h * f_i(node k)
    int i, k, kg;           // state, node, group of the node
    unsigned all, mix;
    bool row;               // this thread owns a defect row
};
template <int N>
ECUDA_HD void rn_row_offsets(const ProbDev& pb, RnRow<N>& st) {
    const int kr = st.k & (kRnGroup - 1);
    const unsigned dlt = static_cast<unsigned>(pb.xcnt[st.i] - 1), kdo = static_cast<unsigned>(pb.xrank[st.i][st.i]);
    st.kg = st.k / kRnGroup;
    st.all = dlt * 0x01010101u;
    st.mix = (st.all & ((1u << (8 * kr)) - 1u)) | (kdo << (8 * kr));
}

// ---- finite differences, part 1: value of defect row (k,i)                     [rows_values]
template <int M, int N, bool SUM>
ECUDA_HD void rn_fd_begin(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, const CtaMem& cm, int b,
                          int tid, RnRow<N>& st, double& viol) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    st.row = tid < NS * N;
    if (!st.row) return;
    const int nc = pb.nc;
    const int i = tid / N, k = tid - i * N;
    st.i = i;
    st.k = k;
    rn_row_offsets<N>(pb, st);
    const double* zx = m.z + nc * N;  // X(l,j) = zx[l*NS + j]
    const int r = ph.goff + k * NS + i;
    st.sgr = ECUDA_LDG(pb.sg + r);
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double t = h * m.tau[k] + mid;
    double x[NS], u[NCU], f[NS];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = zx[k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    Model<M>::f(x, u, t, f);
    double fi = 0.0;
#pragma unroll
    for (int a = 0; a < NS; ++a)
        if (a == i) fi = f[a];
    st.hfv = h * fi;
    const double dv = rn_dot<NS, N>(m.dt + k, zx + i, st.P);
    if (io.g) {
        const double val = st.sgr * (dv - st.hfv);
        RN_LOCAL_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, val);
        if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, 0));
    }
}

// ---- finite differences, part 2: the D-coupled triplets of row (k,i) in the state columns of node group G (nodes
// kRnGroup*G ...), by index-set central differences, row-restricted (operation sequence of fast_fd_block).
// out[e] is the slot of triplet e of the instance (a shared-memory ring buffer, or the global array).
// RING: `out` is a shared-memory ring buffer (plain stores); otherwise the global triplet array (streaming stores)
template <int NS, int N, int G, bool RING>
ECUDA_HD void rn_fd_group(const ProbDev& pb, const PhaseDev& ph, const RnMem& m, RnRow<N>& st, double* __restrict__ out) {
    constexpr int BL = ECUDA_DOT_BLOCK, NB = (N + BL - 1) / BL, BI = (G * kRnGroup) / BL, l0 = BI * BL;
    constexpr int a0 = G * kRnGroup - l0;                  // first node of the group inside its summation block
    constexpr int nin = (N - l0) < BL ? (N - l0) : BL;     // nodes of the block
    constexpr int a1 = (a0 + kRnGroup) < nin ? (a0 + kRnGroup) : nin;
    if (!st.row) return;
    const double* Dtk = m.dt + st.k;
    const double* Xi = m.z + pb.nc * N + st.i;
    const FdRec* Ri = m.rec + pb.nc * N + st.i;
    if (a0 == 0) {  // first group of a summation block: its D entries and node values, the prefix of block sums
#pragma unroll
        for (int a = 0; a < nin; ++a) {
            st.d[a] = Dtk[(l0 + a) * N];
            st.xv[a] = Xi[(l0 + a) * NS];
        }
        st.pre = 0.0;
#pragma unroll
        for (int t = 0; t < BI; ++t) st.pre = (t == 0) ? st.P[0] : st.pre + st.P[t];
        st.q = 0.0;
    }
    const unsigned w = G < st.kg ? st.all : (G == st.kg ? st.mix : 0u);
    double* ok = out + st.k;
#pragma unroll
    for (int a = a0; a < a1; ++a) {
        const FdVals rc = rn_load(Ri + (l0 + a) * NS);
        double sp = fma(st.d[a], rc.xp, st.q);
        double sm = fma(st.d[a], rc.xm, st.q);
#pragma unroll
        for (int e = a + 1; e < nin; ++e) {
            sp = fma(st.d[e], st.xv[e], sp);
            sm = fma(st.d[e], st.xv[e], sm);
        }
        double tp = (BI > 0) ? st.pre + sp : sp;
        double tm = (BI > 0) ? st.pre + sm : sm;
#pragma unroll
        for (int t = BI + 1; t < NB; ++t) {
            tp = tp + st.P[t];
            tm = tm + st.P[t];
        }
        const double gp = st.sgr * (tp - st.hfv);
        const double gm = st.sgr * (tm - st.hfv);
        if (RING)
            ok[rc.cp + rn_byte(w, a - a0)] = (gp - gm) * rc.ri;
        else
            ECUDA_STREAM_STORE(ok + (rc.cp + rn_byte(w, a - a0)), (gp - gm) * rc.ri);
        st.q = fma(st.d[a], st.xv[a], st.q);
    }
}

// ---- finite differences, part 3: the node-local triplets of row (k,i), all columns in one function
// [rows_jacobian<FD>, node-local part]
// Runs after every D-coupled group has been stored (the bulk stores are complete): a slot inside a group's range that
// this part writes overwrites what the ring buffer held there.
template <int M, int N>
ECUDA_HD void rn_fd_local_all(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b, const RnRow<N>& st) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU, NB = (N + 7) / 8;
    if (!st.row || !io.jac) return;
    const int nc = pb.nc, i = st.i, k = st.k;
    const double sgr = st.sgr;
    const double (&P)[NB] = st.P;
    const double* zx = m.z + nc * N;
    const double* Dtk = m.dt + k;
    const double* Xi = zx + i;
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    constexpr bool DS = Model<M>::DIAG_FREE;
    const FdRec* rx = m.rec + nc * N;  // record of X(l,j) = rx[l*NS + j]
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double tau = m.tau[k];
    const double t = h * tau + mid;
    double dv = P[0];
#pragma unroll
    for (int bi = 1; bi < NB; ++bi) dv = dv + P[bi];
    double x[NS], u[NCU];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = zx[k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    double dpk = 0.0, dmk = 0.0;
    if (!DS) {
        const FdRec& rc = rx[k * NS + i];
        rn_diag<NS, N>(Dtk, Xi, rc.xp, rc.xm, k, P, dpk, dmk);
    }
    // the node's state columns X(k,j)                                   [xcol_local_fd, row i]
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int rk = pb.xrank[j][i];
        if (rk < 0 || (DS && j == i)) continue;
        const FdRec& rc = rx[k * NS + j];
        if (j != i && !reads_state<M>(i, j)) {  // f_i does not read x_j: g+ == g- bit for bit
            RN_LOCAL_STORE(jac + (rc.cp + k + rk), 0.0);
            continue;
        }
        double xq[NS], xr[NS], fp[NS], fm[NS];
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            xq[a] = (a == j) ? rc.xp : x[a];
            xr[a] = (a == j) ? rc.xm : x[a];
        }
        Model<M>::f(xq, u, t, fp);
        Model<M>::f(xr, u, t, fm);
        double fpi = 0.0, fmi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) {
                fpi = fp[a];
                fmi = fm[a];
            }
        const double gp = sgr * (((i == j) ? dpk : dv) - h * fpi);
        const double gm = sgr * (((i == j) ? dmk : dv) - h * fmi);
        RN_LOCAL_STORE(jac + (rc.cp + k + rk), (gp - gm) * rc.ri);
    }
    // the node's control columns U(k,c)                                 [node_item, c < nc, row i]
    for (int c = 0; c < nc; ++c) {
        const int rk = pb.urank[c][i];
        if (rk < 0) continue;
        const FdRec& rc = m.rec[k * nc + c];
        if (c >= NCU || !reads_control<M>(i, c)) {  // unused or unread control: exactly +0.0
            RN_LOCAL_STORE(jac + (rc.cp + rk), 0.0);
            continue;
        }
        double up[NCU], um[NCU], fp[NS], fm[NS];
#pragma unroll
        for (int a = 0; a < NCU; ++a) {
            up[a] = (a == c) ? rc.xp : u[a];
            um[a] = (a == c) ? rc.xm : u[a];
        }
        Model<M>::f(x, up, t, fp);
        Model<M>::f(x, um, t, fm);
        double fpi = 0.0, fmi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) {
                fpi = fp[a];
                fmi = fm[a];
            }
        const double gp = sgr * (dv - h * fpi);
        const double gm = sgr * (dv - h * fmi);
        RN_LOCAL_STORE(jac + (rc.cp + rk), (gp - gm) * rc.ri);
    }
    // t0 / tf columns                                                    [node_item, time columns, row i]
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const FdRec& rc = m.rec[(NS + nc) * N + which];
        const double t0p = which == 0 ? rc.xp : t0, tfp = which == 1 ? rc.xp : tf;
        const double t0m = which == 0 ? rc.xm : t0, tfm = which == 1 ? rc.xm : tf;
        const double hp = 0.5 * (tfp - t0p), mp = 0.5 * (tfp + t0p);
        const double hm = 0.5 * (tfm - t0m), mm = 0.5 * (tfm + t0m);
        const double tp = hp * tau + mp, tm = hm * tau + mm;
        double fp[NS], fm[NS];
        Model<M>::f(x, u, tp, fp);
        Model<M>::f(x, u, tm, fm);
        double fpi = 0.0, fmi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) {
                fpi = fp[a];
                fmi = fm[a];
            }
        const double gp = sgr * (dv - hp * fpi);
        const double gm = sgr * (dv - hm * fmi);
        RN_LOCAL_STORE(jac + (rc.cp + k * NS + i), (gp - gm) * rc.ri);
    }
}

// ---- the same triplets as separate PIECES (ECUDA_RN_INTERLEAVE and the store-ring variant)
// One PIECE per column of the node: P < NS the state column X(k,P), NS <= P < NS + 8 the control column U(k,P-NS),
// then t0 and tf. A piece is straight-line code (loads, two evaluations of the dynamics, one store), independent of
// the D-coupled groups. They run after the last group; ECUDA_RN_INTERLEAVE runs piece P right after node group
// P % ngroups instead (tried to hide their dependent loads behind the groups' FP64 work: slower, see the macro).
// For models whose f_i reads x_i the row's own state column needs the perturbed diagonal dots and must be written
// after the group that stores the provisional diagonal value: rn_fd_end does it.
constexpr int kRnPieces = ECUDA_MAX_STATES + ECUDA_MAX_CONTROLS + 2;
template <int M, int N, int P>
ECUDA_HD void rn_fd_piece(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b, const RnRow<N>& st,
                          bool own_diag) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU, NB = (N + 7) / 8;
    constexpr bool DS = Model<M>::DIAG_FREE;
    constexpr bool IS_X = P < NS, IS_U = P >= ECUDA_MAX_STATES && P < ECUDA_MAX_STATES + ECUDA_MAX_CONTROLS,
                   IS_T = P >= ECUDA_MAX_STATES + ECUDA_MAX_CONTROLS;
    if (!(IS_X || IS_U || IS_T)) return;  // state slots beyond the model's states
    if (!st.row || !io.jac) return;
    const int nc = pb.nc, i = st.i, k = st.k;
    if (IS_U && P - ECUDA_MAX_STATES >= nc) return;  // uniform over the CTA
    if (IS_X && !own_diag && !DS && P == i) return;  // the row's own column: see rn_fd_end
    if (IS_X && own_diag && P != i) return;
    const double sgr = st.sgr;
    const double* zx = m.z + nc * N;
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double tau = m.tau[k];
    const double t = h * tau + mid;
    double dv = st.P[0];
#pragma unroll
    for (int bi = 1; bi < NB; ++bi) dv = dv + st.P[bi];
    double x[NS], u[NCU];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = zx[k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    if (IS_X) {  // the node's state column X(k,j)                          [xcol_local_fd, row i]
        constexpr int j = IS_X ? P : 0;
        const int rk = pb.xrank[j][i];
        if (rk < 0 || (DS && j == i)) return;
        const FdRec& rc = m.rec[nc * N + k * NS + j];
        if (j != i && !reads_state<M>(i, j)) {  // f_i does not read x_j: g+ == g- bit for bit
            RN_LOCAL_STORE(jac + (rc.cp + k + rk), 0.0);
            return;
        }
        double dpk = 0.0, dmk = 0.0;
        if (!DS && j == i) rn_diag<NS, N>(m.dt + k, zx + i, rc.xp, rc.xm, k, st.P, dpk, dmk);
        double xq[NS], xr[NS], fp[NS], fm[NS];
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            xq[a] = (a == j) ? rc.xp : x[a];
            xr[a] = (a == j) ? rc.xm : x[a];
        }
        Model<M>::f(xq, u, t, fp);
        Model<M>::f(xr, u, t, fm);
        double fpi = 0.0, fmi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) {
                fpi = fp[a];
                fmi = fm[a];
            }
        const double gp = sgr * (((i == j) ? dpk : dv) - h * fpi);
        const double gm = sgr * (((i == j) ? dmk : dv) - h * fmi);
        RN_LOCAL_STORE(jac + (rc.cp + k + rk), (gp - gm) * rc.ri);
    } else if (IS_U) {  // the node's control column U(k,c)                 [node_item, c < nc, row i]
        constexpr int c = IS_U ? P - ECUDA_MAX_STATES : 0;
        const int rk = pb.urank[c][i];
        if (rk < 0) return;
        const FdRec& rc = m.rec[k * nc + c];
        if constexpr (c >= NCU) {  // a control the model does not use: exactly +0.0
            RN_LOCAL_STORE(jac + (rc.cp + rk), 0.0);
        } else {
            if (!reads_control<M>(i, c)) {  // unread control: exactly +0.0
                RN_LOCAL_STORE(jac + (rc.cp + rk), 0.0);
                return;
            }
            double up[NCU], um[NCU], fp[NS], fm[NS];
#pragma unroll
            for (int a = 0; a < NCU; ++a) {
                up[a] = (a == c) ? rc.xp : u[a];
                um[a] = (a == c) ? rc.xm : u[a];
            }
            Model<M>::f(x, up, t, fp);
            Model<M>::f(x, um, t, fm);
            double fpi = 0.0, fmi = 0.0;
#pragma unroll
            for (int a = 0; a < NS; ++a)
                if (a == i) {
                    fpi = fp[a];
                    fmi = fm[a];
                }
            const double gp = sgr * (dv - h * fpi);
            const double gm = sgr * (dv - h * fmi);
            RN_LOCAL_STORE(jac + (rc.cp + rk), (gp - gm) * rc.ri);
        }
    } else {  // t0 / tf column                                            [node_item, time columns, row i]
        constexpr int which = IS_T ? P - ECUDA_MAX_STATES - ECUDA_MAX_CONTROLS : 0;
        const FdRec& rc = m.rec[(NS + nc) * N + which];
        const double t0p = which == 0 ? rc.xp : t0, tfp = which == 1 ? rc.xp : tf;
        const double t0m = which == 0 ? rc.xm : t0, tfm = which == 1 ? rc.xm : tf;
        const double hp = 0.5 * (tfp - t0p), mp = 0.5 * (tfp + t0p);
        const double hm = 0.5 * (tfm - t0m), mm = 0.5 * (tfm + t0m);
        const double tp = hp * tau + mp, tm = hm * tau + mm;
        double fp[NS], fm[NS];
        Model<M>::f(x, u, tp, fp);
        Model<M>::f(x, u, tm, fm);
        double fpi = 0.0, fmi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) {
                fpi = fp[a];
                fmi = fm[a];
            }
        const double gp = sgr * (dv - hp * fpi);
        const double gm = sgr * (dv - hm * fmi);
        RN_LOCAL_STORE(jac + (rc.cp + k * NS + i), (gp - gm) * rc.ri);
    }
}
// the pieces that run after node group G: P = G, G + ngroups, ...
template <int M, int N, int G, int P = G>
struct RnPiecesAfter {
    ECUDA_HD static void run(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b,
                             const RnRow<N>& st) {
        rn_fd_piece<M, N, P>(pb, ph, io, m, b, st, false);
        RnPiecesAfter<M, N, G, (P + (N + kRnGroup - 1) / kRnGroup < kRnPieces) ? P + (N + kRnGroup - 1) / kRnGroup : -1>::run(
            pb, ph, io, m, b, st);
    }
};
template <int M, int N, int G>
struct RnPiecesAfter<M, N, G, -1> {
    ECUDA_HD static void run(const ProbDev&, const PhaseDev&, const EvalIO&, const RnMem&, int, const RnRow<N>&) {}
};
// what is left after the last group: the row's own state column for models whose f_i reads x_i
template <int M, int N, int J = 0>
struct RnOwnDiag {
    ECUDA_HD static void run(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b,
                             const RnRow<N>& st) {
        rn_fd_piece<M, N, J>(pb, ph, io, m, b, st, true);
        RnOwnDiag<M, N, J + 1>::run(pb, ph, io, m, b, st);
    }
};
template <int M, int N>
struct RnOwnDiag<M, N, Model<M>::NS> {
    ECUDA_HD static void run(const ProbDev&, const PhaseDev&, const EvalIO&, const RnMem&, int, const RnRow<N>&) {}
};
template <int M, int N>
ECUDA_HD void rn_fd_end(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b, const RnRow<N>& st) {
    if (ECUDA_RN_INTERLEAVE) {
        if (!Model<M>::DIAG_FREE) RnOwnDiag<M, N>::run(pb, ph, io, m, b, st);
    } else {
        rn_fd_local_all<M, N>(pb, ph, io, m, b, st);  // one function: 0.159 ms on C2; as pieces run at the end 0.167 ms
    }
}

// path row q at (x, y, t); TRK = false: the problem has no moving zones, every path row is a static record
template <int M, bool TRK>
ECUDA_HD double rn_path_row(const ProbDev& pb, const PhaseDev& ph, const CtaMem& cm, int q, double x, double y, double t) {
    if (!TRK) return Model<M>::static_row(cm.inst + ph.inst_off + q * Model<M>::REC, x, y);
    return path_row<M>(pb, ph, cm, q, x, y, t);
}

// ---- the other rows, finite differences: one item each [other_item<FD = true>] -----------------------------------
// item numbering: 0 objective | path rows (k,q) | event rows | duration row | linkage state rows towards the next
// phase | linkage triplets of the previous phase's rows in this phase
template <int M, int N>
ECUDA_HD int rn_items(const ProbDev& pb, const PhaseDev& ph, int p) {
    return 1 + ph.npath * N + pb.ne + 1 + (p + 1 < pb.nphases ? pb.ns : 0) + (p > 0 ? pb.ns : 0);
}

template <int M, int N, bool TRK, bool SUM>
ECUDA_HD void rn_item_fd(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const RnMem& m, const CtaMem& cm,
                         int b, int it, double& viol, double& fval) {
    constexpr int NS = Model<M>::NS;
    const int nc = pb.nc, np = ph.npath, ntr = np - ph.nstat;
    const double* zx = m.z + nc * N;
    const FdRec* rx = m.rec + nc * N;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double* sg = pb.sg;
    double* g = io.g ? io.g + static_cast<size_t>(b) * pb.ncons : nullptr;
    double* jac = io.jac ? io.jac + static_cast<size_t>(b) * pb.nnz : nullptr;
    const int tcol = (NS + nc) * N;
    auto note = [&](int r, double val, int cls) {
        if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, cls));
    };

    if (it == 0) {  // ---- objective: running cost per node and quadrature
#if defined(__CUDA_ARCH__) && ECUDA_RN_OBJWARP
        return;  // the kernel's last warp has computed it cooperatively (rn_objective_warp)
#endif
        if (!io.f) return;
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
            const double t = h * m.tau[k] + mid;
            const double L = Model<M>::cost(zx + k * NS, m.z + k * nc, t);
            acc = fma(m.w[k], pb.maximize ? -1.0 * L : L, acc);
        }
        const double fp = h * acc;
        if (pb.nphases == 1)
            io.f[b] = pb.sf * fp;
        else
            io.fpart[static_cast<size_t>(b) * pb.nphases + p] = fp;
        fval = pb.sf * fp;
        return;
    }
    it -= 1;
    if (it < np * N) {  // ---- path row (k,q)
        int k, q;
        path_item(ph, it, k, q);
        const double tau = m.tau[k];
        const double t = h * tau + mid;
        const double x0 = zx[k * NS], x1 = zx[k * NS + 1];
        const int r = ph.goff + NS * N + pb.ne + k * np + q;
        const double s = ECUDA_LDG(sg + r);
        // a moving-zone row at the node time: the zone's centre (two divisions, a waypoint search) is computed once for
        // the value and the four position perturbations -- the same operations, hence the same bits, as track_row
        const bool trk_row = TRK && q >= ph.nstat && q - ph.nstat < pb.ntracks;
        const double* trk = cm.inst + pb.track_off + (trk_row ? q - ph.nstat : 0) * pb.track_size;
        double xc = 0.0, yc = 0.0;
        if (trk_row) track_center(trk, pb.nway, t, &xc, &yc);
        auto row_at = [&](double xa, double ya) {
            return trk_row ? track_row_at(trk, xc, yc, xa, ya) : rn_path_row<M, TRK>(pb, ph, cm, q, xa, ya, t);
        };
        if (g) {
            const double val = s * row_at(x0, x1);
            RN_LOCAL_STORE(g + r, val);
            note(r, val, 1);
        }
        if (!jac) return;
        const int ev = (k == 0 || k == N - 1) ? 1 : 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const FdRec& rc = rx[k * NS + j];
            const double vp = row_at(j == 0 ? rc.xp : x0, j == 1 ? rc.xp : x1);
            const double vm = row_at(j == 0 ? rc.xm : x0, j == 1 ? rc.xm : x1);
            RN_LOCAL_STORE(jac + (rc.cp + N - 1 + pb.xcnt[j] + ev + q), (s * vp - s * vm) * rc.ri);
        }
        if (TRK && q >= ph.nstat) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const FdRec& rc = m.rec[tcol + which];
                const double t0p = which == 0 ? rc.xp : t0, tfp = which == 1 ? rc.xp : tf;
                const double t0m = which == 0 ? rc.xm : t0, tfm = which == 1 ? rc.xm : tf;
                const double hp = 0.5 * (tfp - t0p), mp = 0.5 * (tfp + t0p);
                const double hm = 0.5 * (tfm - t0m), mm = 0.5 * (tfm + t0m);
                const double tp = hp * tau + mp, tm = hm * tau + mm;
                const double vp = rn_path_row<M, TRK>(pb, ph, cm, q, x0, x1, tp);
                const double vm = rn_path_row<M, TRK>(pb, ph, cm, q, x0, x1, tm);
                RN_LOCAL_STORE(jac + (rc.cp + NS * N + k * ntr + (q - ph.nstat)), (s * vp - s * vm) * rc.ri);
            }
        }
        return;
    }
    it -= np * N;
    if (it < pb.ne) {  // ---- event row: x(t0) or x(tf)
        const int e = it;
        const int node = (e < NS) ? 0 : N - 1, i = (e < NS) ? e : e - NS;
        const int r = ph.goff + NS * N + e;
        const int lx = node * NS + i;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * zx[lx];
            RN_LOCAL_STORE(g + r, val);
            note(r, val, 2);
        }
        if (jac) {
            const FdRec& rc = rx[lx];
            RN_LOCAL_STORE(jac + (rc.cp + N - 1 + pb.xcnt[i]), (s * rc.xp - s * rc.xm) * rc.ri);
        }
        return;
    }
    it -= pb.ne;
    if (it == 0) {  // ---- duration row tf - t0, and the time linkage
        const int r = ph.goff + NS * N + pb.ne + np * N;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * (tf - t0);
            RN_LOCAL_STORE(g + r, val);
            note(r, val, 3);
            if (p + 1 < pb.nphases) {  // time continuity with the next phase
                const PhaseDev& nx = pb.ph[p + 1];
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                const double other = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                RN_LOCAL_STORE(g + rl, ECUDA_LDG(sg + rl) * (tf - other));
            }
        }
        if (!jac) return;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const FdRec& rc = m.rec[tcol + which];
            const int at = rc.cp + NS * N + N * ntr;
            const double ri = rc.ri;
            const double t0p = which == 0 ? rc.xp : t0, tfp = which == 1 ? rc.xp : tf;
            const double t0m = which == 0 ? rc.xm : t0, tfm = which == 1 ? rc.xm : tf;
            RN_LOCAL_STORE(jac + at, (s * (tfp - t0p) - s * (tfm - t0m)) * ri);
            if (which == 0 && p > 0) {
                const PhaseDev& pv = pb.ph[p - 1];
                const int rl = pb.linkoff + (p - 1) * (NS + 1) + NS;
                const double sl = ECUDA_LDG(sg + rl);
                const double o = other_phase_value(pb, io, b, pv.zoff + (NS + nc) * pv.N + 1);
                RN_LOCAL_STORE(jac + at + 1, (sl * (o - t0p) - sl * (o - t0m)) * ri);
            }
            if (which == 1 && p + 1 < pb.nphases) {
                const PhaseDev& nx = pb.ph[p + 1];
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                const double sl = ECUDA_LDG(sg + rl);
                const double o = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                RN_LOCAL_STORE(jac + at + 1, (sl * (tfp - o) - sl * (tfm - o)) * ri);
            }
        }
        return;
    }
    it -= 1;
    // ---- state linkage with the next phase (row owned by this phase) / with the previous phase (triplet of its
    // row in this phase's first node)
    const bool to_next = (p + 1 < pb.nphases) && it < NS;
    const int i = to_next ? it : it - (p + 1 < pb.nphases ? NS : 0);
    const int k = to_next ? N - 1 : 0;
    const int lx = k * NS + i;
    int pos = N - 1 + pb.xcnt[i] + 1;  // after the event triplet of a boundary node
    if (i < 2) pos += np;
    if (to_next) {
        const PhaseDev& nx = pb.ph[p + 1];
        const int r = pb.linkoff + p * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        const double o = other_phase_value(pb, io, b, nx.zoff + nc * nx.N + i);
        if (g) RN_LOCAL_STORE(g + r, s * (zx[lx] - o));
        if (jac) {
            const FdRec& rc = rx[lx];
            RN_LOCAL_STORE(jac + (rc.cp + pos), (s * (rc.xp - o) - s * (rc.xm - o)) * rc.ri);
        }
    } else if (jac) {
        const PhaseDev& pv = pb.ph[p - 1];
        const int r = pb.linkoff + (p - 1) * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        const double o = other_phase_value(pb, io, b, pv.zoff + nc * pv.N + (pv.N - 1) * NS + i);
        const FdRec& rc = rx[lx];
        RN_LOCAL_STORE(jac + (rc.cp + pos), (s * (o - rc.xp) - s * (o - rc.xm)) * rc.ri);
    }
}

// ---- exact mode --------------------------------------------------------------------------------------------------
// D-coupled triplets of defect row (k,i): (sg_r * D[k][l]) / sz(X(l,i)) -- the expression of build_jac_template, so
// the values equal the round-1 template copy bit for bit. They are COMPUTED here (two multiplications per triplet, D
// from L1, the column's record with one 128-bit shared load) instead of being streamed from a per-problem template:
// no copy warp and no L2 -> SM template traffic; they leave through the same shared-memory store ring as in FD mode.
// part 1: value of defect row (k,i)
template <int M, int N, bool SUM>
ECUDA_HD void rn_ex_begin(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, const CtaMem& cm, int b,
                          int tid, RnRow<N>& st, double& viol) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    st.row = tid < NS * N;
    if (!st.row) return;
    const int nc = pb.nc;
    const int i = tid / N, k = tid - i * N;
    st.i = i;
    st.k = k;
    rn_row_offsets<N>(pb, st);
    const double* zx = m.z + nc * N;
    const int r = ph.goff + k * NS + i;
    st.sgr = ECUDA_LDG(pb.sg + r);
    if (!io.g) return;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double t = h * ECUDA_LDG(ph.tau + k) + mid;
    double x[NS], u[NCU], f[NS];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = zx[k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    Model<M>::f(x, u, t, f);
    double fi = 0.0;
#pragma unroll
    for (int a = 0; a < NS; ++a)
        if (a == i) fi = f[a];
    const double dv = rn_dot<NS, N>(ph.Dt + k, zx + i, st.P);
    const double val = st.sgr * (dv - h * fi);
    ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, val);
    if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, 0));
}

// D-coupled triplets of row (k,i) in the state columns of node group G (see rn_fd_group for `out`)
template <int NS, int N, int G, bool RING>
ECUDA_HD void rn_ex_group(const ProbDev& pb, const PhaseDev& ph, const RnMem& m, const RnRow<N>& st, double* __restrict__ out) {
    constexpr int l0 = G * kRnGroup;
    constexpr int nin = (N - l0) < kRnGroup ? (N - l0) : kRnGroup;
    if (!st.row) return;
    const double* Dtk = ph.Dt + st.k;
    const ExRec* Ri = m.erec + pb.nc * N + st.i;
    const unsigned w = G < st.kg ? st.all : (G == st.kg ? st.mix : 0u);
    double* ok = out + st.k;
#pragma unroll
    for (int a = 0; a < nin; ++a) {
        const ExVals rc = rn_load(Ri + (l0 + a) * NS);
        // the l == k value, (sg D_kk) / sz, lands in the row's diagonal slot; rn_ex_end overwrites it with the full entry
        const double v = (st.sgr * ECUDA_LDG(Dtk + (l0 + a) * N)) * rc.isz;
        if (RING)
            ok[rc.cp + rn_byte(w, a)] = v;
        else
            ECUDA_STREAM_STORE(ok + (rc.cp + rn_byte(w, a)), v);
    }
}

// node-local triplets of row (k,i), exact                                   [rows_jacobian<exact>]
template <int M, int N>
ECUDA_HD void rn_ex_end(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b, const RnRow<N>& st) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    if (!st.row || !io.jac) return;
    const int nc = pb.nc, i = st.i, k = st.k;
    const double sgr = st.sgr;
    const double* zx = m.z + nc * N;
    const double* Dtk = ph.Dt + k;
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    const ExRec* rx = m.erec + nc * N;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double t = h * ECUDA_LDG(ph.tau + k) + mid;
    double x[NS], u[NCU], f[NS];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = zx[k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    Model<M>::f(x, u, t, f);
    double fi = 0.0;
#pragma unroll
    for (int a = 0; a < NS; ++a)
        if (a == i) fi = f[a];
    double dfdx[NS][NS], dfdu[NS][NCU];
    Model<M>::jac(x, u, t, dfdx, dfdu);
    double fti = 0.0;
    if constexpr (Model<M>::TDEP) {  // dynamics that read t: - h (df_i/dt) (d t_k / d t0|tf) in the time columns
        double ft[NS], Lt;
        Model<M>::dtime(x, u, t, ft, &Lt);
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) fti = ft[a];
    }
    const double dkk = ECUDA_LDG(Dtk + k * N);
#pragma unroll
    for (int j = 0; j < NS; ++j) {  // [xcol_local_exact, row i]
        const int rk = pb.xrank[j][i];
        if (rk < 0) continue;
        const ExVals rc = rn_load(rx + k * NS + j);
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) d = dfdx[a][j];
        const double v = ((i == j) ? dkk : 0.0) - h * d;
        ECUDA_STREAM_STORE(jac + (rc.cp + k + rk), (sgr * v) * rc.isz);
    }
    for (int c = 0; c < nc; ++c) {  // [node_item exact, control columns, row i]
        const int rk = pb.urank[c][i];
        if (rk < 0) continue;
        const ExVals rc = rn_load(m.erec + k * nc + c);
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int c2 = 0; c2 < NCU; ++c2)
                if (a == i && c2 == c) d = dfdu[a][c2];
        const double v = -(h * d);
        ECUDA_STREAM_STORE(jac + (rc.cp + rk), (sgr * v) * rc.isz);
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // [node_item exact, time columns, row i]
        const ExVals rc = rn_load(m.erec + (NS + nc) * N + which);
        double v = which == 0 ? 0.5 * fi : -0.5 * fi;
        if constexpr (Model<M>::TDEP) {
            const double tau = ECUDA_LDG(ph.tau + k);
            v = v - h * (fti * (which == 0 ? 0.5 * (1.0 - tau) : 0.5 * (1.0 + tau)));
        }
        ECUDA_STREAM_STORE(jac + (rc.cp + k * NS + i), (sgr * v) * rc.isz);
    }
}

// the other rows, exact   [other_item<FD = false>]
template <int M, int N, bool TRK, bool SUM>
ECUDA_HD void rn_item_exact(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const RnMem& m, const CtaMem& cm,
                            int b, int it, double& viol, double& fval) {
    constexpr int NS = Model<M>::NS;
    const int nc = pb.nc, np = ph.npath, ntr = np - ph.nstat;
    const double* zx = m.z + nc * N;
    const ExRec* rx = m.erec + nc * N;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double* sg = pb.sg;
    double* g = io.g ? io.g + static_cast<size_t>(b) * pb.ncons : nullptr;
    double* jac = io.jac ? io.jac + static_cast<size_t>(b) * pb.nnz : nullptr;
    const int tcol = (NS + nc) * N;
    auto note = [&](int r, double val, int cls) {
        if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, cls));
    };

    if (it == 0) {  // ---- objective
#if defined(__CUDA_ARCH__) && ECUDA_RN_OBJWARP
        return;  // the kernel's last warp has computed it cooperatively (rn_objective_warp)
#endif
        if (!io.f) return;
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
            const double t = h * ECUDA_LDG(ph.tau + k) + mid;
            const double L = Model<M>::cost(zx + k * NS, m.z + k * nc, t);
            acc = fma(ECUDA_LDG(ph.w + k), pb.maximize ? -1.0 * L : L, acc);
        }
        const double fp = h * acc;
        if (pb.nphases == 1)
            io.f[b] = pb.sf * fp;
        else
            io.fpart[static_cast<size_t>(b) * pb.nphases + p] = fp;
        fval = pb.sf * fp;
        return;
    }
    it -= 1;
    if (it < np * N) {  // ---- path row (k,q)
        int k, q;
        path_item(ph, it, k, q);
        const double tau = ECUDA_LDG(ph.tau + k);
        const double t = h * tau + mid;
        const double x0 = zx[k * NS], x1 = zx[k * NS + 1];
        const int r = ph.goff + NS * N + pb.ne + k * np + q;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * rn_path_row<M, TRK>(pb, ph, cm, q, x0, x1, t);
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 1);
        }
        if (!jac) return;
        const int ev = (k == 0 || k == N - 1) ? 1 : 0;
        double ddx, ddy, ddt = 0.0;
        if (!TRK || q < ph.nstat)
            Model<M>::static_row_dxy(cm.inst + ph.inst_off + q * Model<M>::REC, x0, x1, &ddx, &ddy);
        else
            track_row_partials(cm.inst + pb.track_off + (q - ph.nstat) * pb.track_size, pb.nway, x0, x1, t, &ddx, &ddy, &ddt);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const ExVals rc = rn_load(rx + k * NS + j);
            const double v = (j == 0) ? ddx : ddy;
            ECUDA_STREAM_STORE(jac + (rc.cp + N - 1 + pb.xcnt[j] + ev + q), (s * v) * rc.isz);
        }
        if (TRK && q >= ph.nstat) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const ExVals rc = rn_load(m.erec + tcol + which);
                const double dtk = which == 0 ? 0.5 * (1.0 - tau) : 0.5 * (1.0 + tau);
                ECUDA_STREAM_STORE(jac + (rc.cp + NS * N + k * ntr + (q - ph.nstat)), (s * (ddt * dtk)) * rc.isz);
            }
        }
        return;
    }
    it -= np * N;
    if (it < pb.ne) {  // ---- event row
        const int e = it;
        const int node = (e < NS) ? 0 : N - 1, i = (e < NS) ? e : e - NS;
        const int r = ph.goff + NS * N + e;
        const int lx = node * NS + i;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * zx[lx];
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 2);
        }
        if (jac) {
            const ExVals rc = rn_load(rx + lx);
            ECUDA_STREAM_STORE(jac + (rc.cp + N - 1 + pb.xcnt[i]), (s * 1.0) * rc.isz);
        }
        return;
    }
    it -= pb.ne;
    if (it == 0) {  // ---- duration row tf - t0, and the time linkage
        const int r = ph.goff + NS * N + pb.ne + np * N;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * (tf - t0);
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 3);
            if (p + 1 < pb.nphases) {
                const PhaseDev& nx = pb.ph[p + 1];
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                const double other = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                ECUDA_STREAM_STORE(g + rl, ECUDA_LDG(sg + rl) * (tf - other));
            }
        }
        if (!jac) return;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const ExVals rc = rn_load(m.erec + tcol + which);
            const int at = rc.cp + NS * N + N * ntr;
            ECUDA_STREAM_STORE(jac + at, (s * (which == 0 ? -1.0 : 1.0)) * rc.isz);
            if (which == 0 && p > 0) {
                const int rl = pb.linkoff + (p - 1) * (NS + 1) + NS;
                ECUDA_STREAM_STORE(jac + at + 1, (ECUDA_LDG(sg + rl) * -1.0) * rc.isz);
            }
            if (which == 1 && p + 1 < pb.nphases) {
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                ECUDA_STREAM_STORE(jac + at + 1, (ECUDA_LDG(sg + rl) * 1.0) * rc.isz);
            }
        }
        return;
    }
    it -= 1;
    // ---- state linkage rows
    const bool to_next = (p + 1 < pb.nphases) && it < NS;
    const int i = to_next ? it : it - (p + 1 < pb.nphases ? NS : 0);
    const int k = to_next ? N - 1 : 0;
    const int lx = k * NS + i;
    int pos = N - 1 + pb.xcnt[i] + 1;
    if (i < 2) pos += np;
    if (to_next) {
        const PhaseDev& nx = pb.ph[p + 1];
        const int r = pb.linkoff + p * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        const double o = other_phase_value(pb, io, b, nx.zoff + nc * nx.N + i);
        if (g) ECUDA_STREAM_STORE(g + r, s * (zx[lx] - o));
        if (jac) {
            const ExVals rc = rn_load(rx + lx);
            ECUDA_STREAM_STORE(jac + (rc.cp + pos), (s * 1.0) * rc.isz);
        }
    } else if (jac) {
        const int r = pb.linkoff + (p - 1) * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        const ExVals rc = rn_load(rx + lx);
        ECUDA_STREAM_STORE(jac + (rc.cp + pos), (s * -1.0) * rc.isz);
    }
}

// Objective of the phase by ONE WARP: lane j evaluates the running cost at nodes j, j + 32, ...; the quadrature is
// the same serial ascending fma chain as everywhere else (objective_phase), its operands arriving by shuffle. As a
// single thread's item it was a 40-iteration loop of dependent loads on the CTA's critical path (the last warp
// finished ~4000 cycles after the others).
#if defined(__CUDA_ARCH__)
template <int M, int N>
__device__ __forceinline__ void rn_objective_warp(const ProbDev& pb, int p, const EvalIO& io, const RnMem& m, int b, int lane,
                                                  double& fval) {
    constexpr int NS = Model<M>::NS, NC32 = (N + 31) / 32;
    if (!io.f) return;
    const int nc = pb.nc;
    const double* zx = m.z + nc * N;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    double Lw[NC32], ww[NC32];
#pragma unroll
    for (int c = 0; c < NC32; ++c) {
        const int k = c * 32 + lane;
        Lw[c] = 0.0;
        ww[c] = 0.0;
        if (k < N) {
            const double L = Model<M>::cost(zx + k * NS, m.z + k * nc, h * m.tau[k] + mid);
            Lw[c] = pb.maximize ? -1.0 * L : L;
            ww[c] = m.w[k];
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k)
        acc = fma(__shfl_sync(0xffffffffu, ww[k >> 5], k & 31), __shfl_sync(0xffffffffu, Lw[k >> 5], k & 31), acc);
    const double fp = h * acc;
    if (lane == 31) {
        if (pb.nphases == 1)
            io.f[b] = pb.sf * fp;
        else
            io.fpart[static_cast<size_t>(b) * pb.nphases + p] = fp;
        fval = pb.sf * fp;
    }
}
#endif

// ---- the three parts of a thread's program after the staging barrier -------------------------------------------------
//   rn_begin   defect-row value (+ the per-row state kept in registers)
//   rn_group   D-coupled triplets of node group G into `out` (k_rows_n: a ring buffer that leaves by bulk store)
//   rn_end     node-local triplets of the row, then the other rows (objective, path, event, duration, linkage)
template <int M, int N, bool FD, bool SUM>
ECUDA_HD void rn_begin(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, const CtaMem& cm, int b,
                       int tid, RnRow<N>& st, double& viol, double& fval) {
    viol = 0.0;
    fval = 0.0;
    if (FD)
        rn_fd_begin<M, N, SUM>(pb, ph, io, m, cm, b, tid, st, viol);
    else
        rn_ex_begin<M, N, SUM>(pb, ph, io, m, cm, b, tid, st, viol);
}
template <int M, int N, bool FD, int G, bool RING>
ECUDA_HD void rn_group(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b, RnRow<N>& st,
                       double* out) {
    if (FD) {
        rn_fd_group<Model<M>::NS, N, G, RING>(pb, ph, m, st, out);
        // node-local pieces interleaved with the groups (direct stores; with the store ring they would have to wait
        // for the bulk stores, so the ring variant keeps them for the end)
        if (!RING && ECUDA_RN_INTERLEAVE) RnPiecesAfter<M, N, G>::run(pb, ph, io, m, b, st);
    } else {
        rn_ex_group<Model<M>::NS, N, G, RING>(pb, ph, m, st, out);
    }
}
// every piece at once (the store-ring variant, after its bulk stores are complete)
template <int M, int N, int G = 0>
struct RnAllPieces {
    ECUDA_HD static void run(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b,
                             const RnRow<N>& st) {
        RnPiecesAfter<M, N, G>::run(pb, ph, io, m, b, st);
        RnAllPieces<M, N, G + 1>::run(pb, ph, io, m, b, st);
    }
};
template <int M, int N>
struct RnAllPieces<M, N, (N + kRnGroup - 1) / kRnGroup> {
    ECUDA_HD static void run(const ProbDev&, const PhaseDev&, const EvalIO&, const RnMem&, int, const RnRow<N>&) {}
};
template <int N>
ECUDA_HD constexpr int rn_ngroups() { return (N + kRnGroup - 1) / kRnGroup; }
// triplet range [c0, c1) of the state columns of node group g: from the first state column of its first node to the
// first state column of the node after its last (the t0 column after the last node)
template <int M, int N, bool FD>
ECUDA_HD void rn_group_range(const ProbDev& pb, const RnMem& m, int g, int& c0, int& c1) {
    constexpr int NS = Model<M>::NS;
    const int l0 = g * kRnGroup, l1 = (l0 + kRnGroup) < N ? (l0 + kRnGroup) : N;
    const int a = pb.nc * N + l0 * NS, e = pb.nc * N + l1 * NS;
    c0 = FD ? m.rec[a].cp : m.erec[a].cp;
    c1 = FD ? m.rec[e].cp : m.erec[e].cp;
}
// RING: the FD node-local pieces have not run yet (see rn_group)
template <int M, int N, bool FD, bool TRK, bool SUM, bool RING = false>
ECUDA_HD void rn_end(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const RnMem& m, const CtaMem& cm, int b,
                     int tid, int nthr, const RnRow<N>& st, double& viol, double& fval) {
    if (FD) {
        if (RING && ECUDA_RN_INTERLEAVE) RnAllPieces<M, N>::run(pb, ph, io, m, b, st);
        rn_fd_end<M, N>(pb, ph, io, m, b, st);
    } else {
        rn_ex_end<M, N>(pb, ph, io, m, b, st);
    }
    const int nitems = rn_items<M, N>(pb, ph, p);
    for (int it = nthr - 1 - tid; it < nitems; it += nthr) {
        if (FD)
            rn_item_fd<M, N, TRK, SUM>(pb, ph, p, io, m, cm, b, it, viol, fval);
        else
            rn_item_exact<M, N, TRK, SUM>(pb, ph, p, io, m, cm, b, it, viol, fval);
    }
}
// node group by run-time index (the kernel-logic emulator of the test-suite; the kernel unrolls the groups)
template <int M, int N, bool FD, int G = 0>
struct RnGroupRt {
    ECUDA_HD static void run(int g, const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const RnMem& m, int b,
                             RnRow<N>& st, double* out) {
        if (g == G)
            rn_group<M, N, FD, G, false>(pb, ph, io, m, b, st, out);
        else
            RnGroupRt<M, N, FD, G + 1>::run(g, pb, ph, io, m, b, st, out);
    }
};
template <int M, int N, bool FD>
struct RnGroupRt<M, N, FD, (N + kRnGroup - 1) / kRnGroup> {
    ECUDA_HD static void run(int, const ProbDev&, const PhaseDev&, const EvalIO&, const RnMem&, int, RnRow<N>&, double*) {}
};

}  // namespace ecuda
#endif
