// ecuda_fast.cuh -- the specialised per-instance evaluation used when every phase has at most one
// defect row per thread (ns*N <= blockDim) and a compile-time number NB of summation blocks.
//
// Same arithmetic as ecuda_phases.cuh (every value is produced by the same operation sequence, so
// the results are bit-identical to the generic path and to the oracle); what changes is how the work
// is laid out on the CTA:
//   * thread (j,k) owns defect row (k,j): its NB block sums of (D X)[k][j] stay in REGISTERS from
//     phase B to phase C (no shared-memory round trip), the summation blocks are unrolled with the
//     block index known at compile time, so the serial prefix / tail of block sums around a
//     perturbed block are register operands and have exactly the length they need;
//   * the diagonal column's perturbed dots are computed once per row instead of being carried as
//     predicated moves through all N entries;
//   * 32-bit triplet index arithmetic, one 64-bit multiply-add per store;
//   * node-level work (path rows of g, path entries of the X columns, control and t0/tf columns) is
//     handed out one item per thread from the top of the thread range, so threads without a defect
//     row start on it and nobody evaluates eight obstacle rows in a row;
//   * exact mode: the D-coupled triplets sg*D/sz do not depend on the instance, so they are copied
//     from a per-problem template (L2 resident) with coalesced loads/stores instead of being
//     recomputed.
// Reference counterparts: as in ecuda_phases.cuh (PSOPT defect assembly + derivative drivers entered
// at src/ePSOPT/ePSOPT.cpp:84, callbacks src/ePSOPT/ePSOPT.cpp:186-306).
#ifndef ECUDA_FAST_CUH_
#define ECUDA_FAST_CUH_

#include "ecuda_phases.cuh"

namespace ecuda {

// per-thread state that lives in registers between phase B and phase C
template <int NB>
struct RowRegs {
    double P[NB];   // block sums of (D X)[k][j]
    double dp, dm;  // (D X)[k][j] with X[k][j] -> +-delta (finite differences only)
    double viol;    // fused summary: max bound violation over the g rows this thread has written
    double fval;    // fused summary: the objective (thread that ran the quadrature)
};

// bound violation of one constraint value (fused summary; io.nranks > 0)
// cls: 0 defect row, 1 path row, 2 event row, 3 duration row. With the compact form of the bounds (io.bev: every
// instance has defect bounds 0 and one pair of path-row bounds) only the event and duration rows read memory.
ECUDA_HD double row_violation(const EvalIO& io, const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int b, int r,
                              double val, int cls) {
    if (io.bev) {
        if (cls == 0) return fabs(val);  // max(0 - val, val - 0)
        if (cls == 1) return fmax(io.plo - val, val - io.phi);
        const int nev = pb.ne + 1, e = cls == 2 ? r - ph.goff - pb.ns * ph.N : pb.ne;
        const double* ev = io.bev + static_cast<size_t>(b) * 2 * nev;
        return fmax(ECUDA_LDG(ev + e) - val, val - ECUDA_LDG(ev + nev + e));
    }
    if (m.bl) return fmax(m.bl[r - ph.goff] - val, val - m.bu[r - ph.goff]);
    const size_t o = static_cast<size_t>(b) * pb.ncons + r;
    return fmax(ECUDA_LDG(io.bl + o) - val, val - ECUDA_LDG(io.bu + o));
}

// block sums + total of row k of D times state j of X (canonical blocked order, see dot_row)
// PARTIAL: the last block has fewer than ECUDA_DOT_BLOCK nodes (N < NB*BL); otherwise no bounds tests
template <int NS, int NB, bool PARTIAL>
ECUDA_HD double fast_dot(const double* __restrict__ Dtk, int N, const double* __restrict__ Xj, double (&P)[NB]) {
    constexpr int BL = ECUDA_DOT_BLOCK;
    double total = 0.0;
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
        const int l0 = bi * BL;
        double p = 0.0;
#pragma unroll
        for (int i = 0; i < BL; ++i)
            if (!PARTIAL || bi < NB - 1 || l0 + i < N) p = fma(ECUDA_LDG(Dtk + (l0 + i) * N), Xj[(l0 + i) * NS], p);
        P[bi] = p;
        total = (bi == 0) ? p : total + p;
    }
    return total;
}

// perturbed dots of the diagonal column: same operation sequence as a D-coupled entry with l == k
template <int NS, int NB>
ECUDA_HD void fast_diag(const double* __restrict__ Dtk, int N, const double* __restrict__ Xj, double xpk, double xmk,
                        int k, const double (&P)[NB], double& dp, double& dm) {
    constexpr int BL = ECUDA_DOT_BLOCK;
    const int bk = k / BL, l0 = bk * BL, krel = k - l0;
    double q = 0.0, sp = 0.0, sm = 0.0;
#pragma unroll
    for (int i = 0; i < BL; ++i) {
        if (l0 + i < N) {
            const double di = ECUDA_LDG(Dtk + (l0 + i) * N);
            const double xi = Xj[(l0 + i) * NS];
            if (i < krel) {
                q = fma(di, xi, q);
            } else if (i == krel) {
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
        if (t < bk) pre = pre + P[t];  // 0.0 + P[0] == P[0] bit for bit (a block sum is never -0.0)
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

// D-coupled triplets of row (k,j) in the state columns of summation block BI, by index-set central
// differences (row-restricted: see ecuda_phases.cuh). jac = triplet array of the instance;
// kpc = k + (number of node-local defect rows of column X(.,j)) - 1.
// DIAGSTORE: the l == k value is the finished diagonal triplet (models whose f_j does not read x_j:
// h*f_j is the same at the perturbed points, so the row's own column needs nothing else) and goes to
// its real slot kd = k + rank of row (k,j) among the node-local rows of column X(k,j).
template <int NS, int NB, int BI, bool PARTIAL, bool DIAGSTORE>
ECUDA_HD void fast_fd_block(const double* __restrict__ Dtk, int N, const double* __restrict__ Xj,
                            const double* __restrict__ XPj, const double* __restrict__ XMj,
                            const double* __restrict__ RIj, const int* __restrict__ CPj, const double (&P)[NB],
                            double sgr, double hfv, int k, int kpc, int kd, double* __restrict__ jac) {
    constexpr int BL = ECUDA_DOT_BLOCK;
    constexpr int l0 = BI * BL;
    constexpr bool LAST = PARTIAL && BI == NB - 1;  // only the last block can be partial
    double d[BL], xv[BL];
#pragma unroll
    for (int i = 0; i < BL; ++i) {
        if (!LAST || l0 + i < N) {
            d[i] = ECUDA_LDG(Dtk + (l0 + i) * N);
            xv[i] = Xj[(l0 + i) * NS];
        } else {
            d[i] = 0.0;
            xv[i] = 0.0;
        }
    }
    double pre = 0.0;
#pragma unroll
    for (int t = 0; t < BI; ++t) pre = (t == 0) ? P[0] : pre + P[t];
    double q = 0.0;  // unperturbed in-block prefix
#pragma unroll
    for (int a = 0; a < BL; ++a) {
        if (!LAST || l0 + a < N) {
            double sp = fma(d[a], XPj[(l0 + a) * NS], q);
            double sm = fma(d[a], XMj[(l0 + a) * NS], q);
#pragma unroll
            for (int i = a + 1; i < BL; ++i)
                if (!LAST || l0 + i < N) {
                    sp = fma(d[i], xv[i], sp);
                    sm = fma(d[i], xv[i], sm);
                }
            double tp = (BI > 0) ? pre + sp : sp;
            double tm = (BI > 0) ? pre + sm : sm;
#pragma unroll
            for (int t = BI + 1; t < NB; ++t) {
                tp = tp + P[t];
                tm = tm + P[t];
            }
            const double gp = sgr * (tp - hfv);
            const double gm = sgr * (tm - hfv);
            const double v = (gp - gm) * RIj[(l0 + a) * NS];
            // rows k < l sit at position k of column X(l,j); rows k > l come after its node-local block.
            // For l == k: DIAGSTORE -> the diagonal triplet's own slot; otherwise the first node-local
            // defect triplet of the thread's own column, which xcol_local_fd overwrites later in this
            // thread's program order (an unconditional store is cheaper than a branch around it).
            const int kge = (DIAGSTORE && (l0 + a) == k) ? kd : k;
            const unsigned idx = static_cast<unsigned>(CPj[(l0 + a) * NS] + ((l0 + a) < k ? kpc : kge));
            ECUDA_STREAM_STORE(jac + idx, v);
            q = fma(d[a], xv[a], q);
        }
    }
}

template <int NS, int NB, int BI, bool DS>
struct FastFdBlocks {
    ECUDA_HD static void run(const double* Dtk, int N, const double* Xj, const double* XPj, const double* XMj,
                             const double* RIj, const int* CPj, const double (&P)[NB], double sgr, double hfv, int k,
                             int kpc, int kd, double* jac) {
        if (BI < NB - 1 || N == NB * ECUDA_DOT_BLOCK)  // uniform over the CTA
            fast_fd_block<NS, NB, BI, false, DS>(Dtk, N, Xj, XPj, XMj, RIj, CPj, P, sgr, hfv, k, kpc, kd, jac);
        else
            fast_fd_block<NS, NB, BI, true, DS>(Dtk, N, Xj, XPj, XMj, RIj, CPj, P, sgr, hfv, k, kpc, kd, jac);
        FastFdBlocks<NS, NB, BI + 1, DS>::run(Dtk, N, Xj, XPj, XMj, RIj, CPj, P, sgr, hfv, k, kpc, kd, jac);
    }
};
template <int NS, int NB, bool DS>
struct FastFdBlocks<NS, NB, NB, DS> {
    ECUDA_HD static void run(const double*, int, const double*, const double*, const double*, const double*,
                             const int*, const double (&)[NB], double, double, int, int, int, double*) {}
};

// exact mode: copy this phase's slice of the per-problem template into the instance's triplet array
// (plain coalesced copies; the node-local triplets are overwritten after the next barrier)
ECUDA_HD void fast_copy_template(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, int b, int tid, int nthr) {
    const int e0 = ECUDA_LDG(pb.colptr + ph.zoff + pb.nc * ph.N);        // first state column of the phase
    const int e1 = ECUDA_LDG(pb.colptr + ph.zoff + (pb.nc + pb.ns) * ph.N);  // its t0 column
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    // batches of UNR independent loads, then UNR stores: one L2 round trip per batch instead of one
    // per element (the template is L2 resident, ~100 KB shared by every CTA)
    constexpr int UNR = 12;
    for (int e = e0 + tid; e < e1; e += UNR * nthr) {
        double v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (e + u * nthr < e1) v[u] = ECUDA_LDG(pb.jtmpl + e + u * nthr);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (e + u * nthr < e1) ECUDA_STREAM_STORE(jac + e + u * nthr, v[u]);
    }
}

// ---- phase B ------------------------------------------------------------------------------------------
template <int M, int NB, bool FD>
ECUDA_HD void fast_phase_b(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int tid,
                           int nthr, RowRegs<NB>& rr) {
    constexpr int NS = Model<M>::NS;
    const int N = ph.N, nc = pb.nc, np = ph.npath;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double* g = io.g ? io.g + static_cast<size_t>(b) * pb.ncons : nullptr;
    const double* sg = pb.sg;
    rr.viol = 0.0;
    rr.fval = 0.0;
    // defect rows: block sums into registers, totals into shared memory
    if (tid < NS * N) {
        const int j = fast_div(tid, ph.mN), k = tid - j * N;
        const double* Dtk = ph.Dt + k;
        const double* Xj = m.z + nc * N + j;
        m.dotv[k * NS + j] = (N == NB * ECUDA_DOT_BLOCK) ? fast_dot<NS, NB, false>(Dtk, N, Xj, rr.P)
                                                         : fast_dot<NS, NB, true>(Dtk, N, Xj, rr.P);
        if (FD && io.jac && !Model<M>::DIAG_FREE) {
            const int lcol = nc * N + k * NS + j;
            fast_diag<NS, NB>(Dtk, N, Xj, m.xp[lcol], m.xm[lcol], k, rr.P, rr.dp, rr.dm);
        }
    }
    // node values, handed out from the top of the thread range (threads without a defect row first)
    for (int k = nthr - 1 - tid; k < N; k += nthr) {
        const double* x = m.z + nc * N + k * NS;
        const double* u = m.z + k * nc;
        const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
        double f[NS];
        Model<M>::f(x, u, t, f);
#pragma unroll
        for (int i = 0; i < NS; ++i) m.hf[k * NS + i] = pt.h * f[i];
        const double L = Model<M>::cost(x, u, t);
        m.Lk[k] = pb.maximize ? -1.0 * L : L;
    }
    if (g) {
        // path rows: one row per thread
        for (int it = nthr - 1 - tid; it < np * N; it += nthr) {
            int k, q;
            path_item(ph, it, k, q);
            const double* x = m.z + nc * N + k * NS;
            const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
            const int r = ph.goff + NS * N + pb.ne + k * np + q;
            const double val = ECUDA_LDG(sg + r) * path_row<M>(pb, ph, m, q, x[0], x[1], t);
            ECUDA_STREAM_STORE(g + r, val);
            if (io.nranks > 0) rr.viol = fmax(rr.viol, row_violation(io, pb, ph, m, b, r, val, 1));
        }
        for (int e = tid; e < pb.ne; e += nthr) {
            const int r = ph.goff + NS * N + e;
            const int node = (e < NS) ? 0 : N - 1;
            const int i = (e < NS) ? e : e - NS;
            const double val = ECUDA_LDG(sg + r) * m.z[nc * N + node * NS + i];
            ECUDA_STREAM_STORE(g + r, val);
            if (io.nranks > 0) rr.viol = fmax(rr.viol, row_violation(io, pb, ph, m, b, r, val, 2));
        }
        if (tid == 0) {
            const int r = ph.goff + NS * N + pb.ne + np * N;
            const double val = ECUDA_LDG(sg + r) * (pt.tf - pt.t0);
            ECUDA_STREAM_STORE(g + r, val);
            if (io.nranks > 0) rr.viol = fmax(rr.viol, row_violation(io, pb, ph, m, b, r, val, 3));
        }
        if (p + 1 < pb.nphases) {
            const PhaseDev& nx = pb.ph[p + 1];
            for (int i = tid; i <= NS; i += nthr) {
                const int r = pb.linkoff + p * (NS + 1) + i;
                const double mine = (i < NS) ? m.z[nc * N + (N - 1) * NS + i] : pt.tf;
                const int ocol = (i < NS) ? nx.zoff + nc * nx.N + i : nx.zoff + (NS + nc) * nx.N;
                const double other = other_phase_value(pb, io, b, ocol);
                ECUDA_STREAM_STORE(g + r, ECUDA_LDG(sg + r) * (mine - other));
            }
        }
    }
}

// ---- phase C ------------------------------------------------------------------------------------------
// SM (exact mode only): triplets go to the shared-memory image whose virtual base is `image`
// (image[e] = slot of triplet e of the instance) instead of the caller's global array.
template <int M, int NB, bool FD, bool SM = false>
ECUDA_HD void fast_phase_c(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int tid,
                           int nthr, RowRegs<NB>& rr, double* image = nullptr) {
    static_assert(!(FD && SM), "the shared-memory image is an exact-mode path");
    constexpr int NS = Model<M>::NS;
    const int N = ph.N, nc = pb.nc, np = ph.npath;
    if (tid == nthr - 1) rr.fval = objective_phase(pb, ph, p, io, m, b);
    double* jac = SM ? image : (io.jac ? io.jac + static_cast<size_t>(b) * pb.nnz : nullptr);
    if (tid < NS * N && (io.g || jac)) {
        const int j = fast_div(tid, ph.mN), k = tid - j * N;
        const int r = ph.goff + k * NS + j;
        const double sgr = ECUDA_LDG(pb.sg + r);
        const double hfv = m.hf[k * NS + j];
        if (io.g) {
            const double val = sgr * (m.dotv[k * NS + j] - hfv);
            ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, val);
            if (io.nranks > 0) rr.viol = fmax(rr.viol, row_violation(io, pb, ph, m, b, r, val, 0));
        }
        if (jac) {
            if (FD) {
                const int xoff = nc * N + j;
                constexpr bool DS = Model<M>::DIAG_FREE;
                FastFdBlocks<NS, NB, 0, DS>::run(ph.Dt + k, N, m.z + xoff, m.xp + xoff, m.xm + xoff, m.rinv + xoff,
                                                 m.colp + xoff, rr.P, sgr, hfv, k, k + pb.xcnt[j] - 1,
                                                 k + pb.xrank[j][j], jac);
                xcol_local_fd<M, DS>(pb, ph, p, io, m, b, j, k, DS ? 0.0 : rr.dp, DS ? 0.0 : rr.dm, jac);
            } else {
                xcol_local_exact<M, SM>(pb, ph, p, m, j, k, jac);
            }
        }
    }
    if (jac) {
        // one item per thread from the top of the thread range: path triplets of the state columns
        // (states 0 and 1 of every node), then the control and t0/tf columns
        const int nP = (NS >= 2) ? N * 2 * np : 0, nU = (nc + 2) * N;
        for (int it = nthr - 1 - tid; it < nP + nU; it += nthr) {
            if (it < nP) {
                const int j = it & 1;  // neighbouring lanes: the same path row, state columns 0 and 1
                int k, q;
                path_item(ph, it >> 1, k, q);
                if (FD)
                    xcol_path_fd<M>(pb, ph, m, j, k, q, jac);
                else
                    xcol_path_exact<M, SM>(pb, ph, m, j, k, q, jac);
            } else {
                const int i2 = it - nP;
                const int c = fast_div(i2, ph.mN), k = i2 - c * N;
                node_item<M, SM>(pb, ph, p, io, m, b, k, c, SM ? jac : nullptr);
            }
        }
    }
}

}  // namespace ecuda
#endif
