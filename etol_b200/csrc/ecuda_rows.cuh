// ecuda_rows.cuh -- row-owner formulation of the specialised evaluation (ns*N <= blockDim rows, NB
// summation blocks): after the decision vector is staged there is NO further CTA-wide dependency.
//
//   * thread (i,k) owns constraint row (k,i) -- defect of state i at node k -- completely: it evaluates
//     the dynamics at its node, keeps the block sums of (D X)[k][i] in registers, writes g[row] and then
//     EVERY triplet of that row: the N-1 D-coupled ones (fast_fd_block), the node's state columns,
//     control columns and the t0/tf columns. Nothing goes through shared memory between threads (no
//     hf / dot arrays, no second barrier), the node's variables are loaded once per row.
//   * the remaining rows -- path rows, event rows, duration row, linkage rows -- are independent items
//     handed out one per thread from the top of the thread range; each item writes its g value and
//     all triplets of its row; the objective is one such item.
// Every value is produced by the operation sequence of the corresponding function of ecuda_phases.cuh
// (cited per block), so results are bit-identical to the generic kernels and to the oracle.
#ifndef ECUDA_ROWS_CUH_
#define ECUDA_ROWS_CUH_

#include "ecuda_fast.cuh"

namespace ecuda {

// 1/sz of a variable of the phase in exact mode: the shared-memory copy the exact row-owner kernel staged in the
// (otherwise unused) rinv slot, else global memory
ECUDA_HD double isz_of(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int lcol) {
    return m.rinv ? m.rinv[lcol] : ECUDA_LDG(pb.isz + ph.zoff + lcol);
}

template <int M, int NB>
struct RowState {
    double P[NB];  // block sums of (D X)[k][i]
    double dv;     // (D X)[k][i]
    double hfv;    // h * f_i(node k)
    double sgr;    // scale of the row
    double viol;   // fused summary: max bound violation over the rows this thread wrote
    double fval;   // fused summary: objective (the thread that ran the quadrature)
};

// ---- part 1: values --------------------------------------------------------------------------------------
// defect row (k,i): dynamics at the node, block sums, g value       [phase_b + state_item's g store]
template <int M, int NB, bool FD>
ECUDA_HD void rows_values(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const CtaMem& m, int b, int tid,
                          RowState<M, NB>& rs) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, nc = pb.nc;
    rs.viol = 0.0;
    rs.fval = 0.0;
    if (tid >= NS * N) return;
    const int i = fast_div(tid, ph.mN), k = tid - i * N;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double x[NS], u[NCU], f[NS];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = m.z[nc * N + k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
    Model<M>::f(x, u, t, f);
    double fi = 0.0;
#pragma unroll
    for (int a = 0; a < NS; ++a)
        if (a == i) fi = f[a];
    rs.hfv = pt.h * fi;
    const double* Dtk = ph.Dt + k;
    const double* Xi = m.z + nc * N + i;
    rs.dv = (N == NB * ECUDA_DOT_BLOCK) ? fast_dot<NS, NB, false>(Dtk, N, Xi, rs.P) : fast_dot<NS, NB, true>(Dtk, N, Xi, rs.P);
    const int r = ph.goff + k * NS + i;
    rs.sgr = ECUDA_LDG(pb.sg + r);
    if (io.g) {
        const double val = rs.sgr * (rs.dv - rs.hfv);
        ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, val);
        if (io.nranks > 0) rs.viol = fmax(rs.viol, row_violation(io, pb, ph, m, b, r, val, 0));
    }
}

// ---- part 2: the triplets of defect row (k,i) ---------------------------------------------------------------
template <int M, int NB, bool FD>
ECUDA_HD void rows_jacobian(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const CtaMem& m, int b, int tid,
                            const RowState<M, NB>& rs) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, nc = pb.nc;
    if (tid >= NS * N || !io.jac) return;
    const int i = fast_div(tid, ph.mN), k = tid - i * N;
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double tau = ECUDA_LDG(ph.tau + k);
    const double t = pt.h * tau + pt.m;
    const double sgr = rs.sgr, dv = rs.dv;
    double x[NS], u[NCU];
#pragma unroll
    for (int a = 0; a < NS; ++a) x[a] = m.z[nc * N + k * NS + a];
#pragma unroll
    for (int a = 0; a < NCU; ++a) u[a] = m.z[k * nc + a];
    const int tcol = (NS + nc) * N;  // phase-local index of t0; tf follows

    if (FD) {
        // D-coupled triplets (and, for DIAG_FREE models, the row's own diagonal triplet)
        constexpr bool DS = Model<M>::DIAG_FREE;
        const int xoff = nc * N + i;
        // the l == k value always goes to the row's own diagonal slot: final for DIAG_FREE models, otherwise
        // overwritten below by this same thread (j == i) -- never a slot another row's thread owns
        FastFdBlocks<NS, NB, 0, true>::run(ph.Dt + k, N, m.z + xoff, m.xp + xoff, m.xm + xoff, m.rinv + xoff, m.colp + xoff,
                                           rs.P, sgr, rs.hfv, k, k + pb.xcnt[i] - 1, k + pb.xrank[i][i], jac);
        double dpk = 0.0, dmk = 0.0;
        if (!DS) {
            const int lc = nc * N + k * NS + i;
            fast_diag<NS, NB>(ph.Dt + k, N, m.z + xoff, m.xp[lc], m.xm[lc], k, rs.P, dpk, dmk);
        }
        // the node's state columns X(k,j)                                   [xcol_local_fd, row i]
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int rk = pb.xrank[j][i];
            if (rk < 0 || (DS && j == i)) continue;
            const int lcol = nc * N + k * NS + j;
            if (j != i && !reads_state<M>(i, j)) {  // f_i does not read x_j: g+ == g- bit for bit
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + k + rk, 0.0);
                continue;
            }
            double xq[NS], xr[NS], fp[NS], fm[NS];
#pragma unroll
            for (int a = 0; a < NS; ++a) {
                xq[a] = (a == j) ? m.xp[lcol] : x[a];
                xr[a] = (a == j) ? m.xm[lcol] : x[a];
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
            const double gp = sgr * (((i == j) ? dpk : dv) - pt.h * fpi);
            const double gm = sgr * (((i == j) ? dmk : dv) - pt.h * fmi);
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + k + rk, (gp - gm) * m.rinv[lcol]);
        }
        // the node's control columns U(k,c)                                 [node_item, c < nc, row i]
        for (int c = 0; c < nc; ++c) {
            const int rk = pb.urank[c][i];
            if (rk < 0) continue;
            const int lcol = k * nc + c;
            if (c >= NCU || !reads_control<M>(i, c)) {  // unused or unread control: exactly +0.0
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + rk, 0.0);
                continue;
            }
            double up[NCU], um[NCU], fp[NS], fm[NS];
#pragma unroll
            for (int a = 0; a < NCU; ++a) {
                up[a] = (a == c) ? m.xp[lcol] : u[a];
                um[a] = (a == c) ? m.xm[lcol] : u[a];
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
            const double gp = sgr * (dv - pt.h * fpi);
            const double gm = sgr * (dv - pt.h * fmi);
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + rk, (gp - gm) * m.rinv[lcol]);
        }
        // t0 / tf columns                                                    [node_item, time columns, row i]
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const int lcol = tcol + which;
            const double t0p = which == 0 ? m.xp[lcol] : pt.t0, tfp = which == 1 ? m.xp[lcol] : pt.tf;
            const double t0m = which == 0 ? m.xm[lcol] : pt.t0, tfm = which == 1 ? m.xm[lcol] : pt.tf;
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
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + k * NS + i, (gp - gm) * m.rinv[lcol]);
        }
    } else {
        // exact: the D-coupled triplets come from the template (copy warp); node-local ones here
        double dfdx[NS][NS], dfdu[NS][NCU], f[NS];
        Model<M>::jac(x, u, t, dfdx, dfdu);
        Model<M>::f(x, u, t, f);
        const double dkk = ECUDA_LDG(ph.Dt + static_cast<size_t>(k) * N + k);
        const double sgi = rs.sgr;
#pragma unroll
        for (int j = 0; j < NS; ++j) {  // [xcol_local_exact, row i]
            const int rk = pb.xrank[j][i];
            if (rk < 0) continue;
            const int lcol = nc * N + k * NS + j;
            double d = 0.0;
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int c2 = 0; c2 < NS; ++c2)
                    if (a == i && c2 == j) d = dfdx[a][c2];
            const double v = ((i == j) ? dkk : 0.0) - pt.h * d;
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + k + rk, (sgi * v) * isz_of(pb, ph, m, lcol));
        }
        for (int c = 0; c < nc; ++c) {  // [node_item exact, control columns, row i]
            const int rk = pb.urank[c][i];
            if (rk < 0) continue;
            const int lcol = k * nc + c;
            double d = 0.0;
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int c2 = 0; c2 < NCU; ++c2)
                    if (a == i && c2 == c) d = dfdu[a][c2];
            const double v = -(pt.h * d);
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + rk, (sgi * v) * isz_of(pb, ph, m, lcol));
        }
        double fi = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) fi = f[a];
        double fti = 0.0;
        if constexpr (Model<M>::TDEP) {  // dynamics that read t: - h (df_i/dt) (d t_k / d t0|tf)
            double ft[NS], Lt;
            Model<M>::dtime(x, u, t, ft, &Lt);
#pragma unroll
            for (int a = 0; a < NS; ++a)
                if (a == i) fti = ft[a];
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {  // [node_item exact, time columns, row i]
            const int lcol = tcol + which;
            double v = which == 0 ? 0.5 * fi : -0.5 * fi;
            if constexpr (Model<M>::TDEP) {
                const double tau = ECUDA_LDG(ph.tau + k);
                v = v - pt.h * (fti * (which == 0 ? 0.5 * (1.0 - tau) : 0.5 * (1.0 + tau)));
            }
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + k * NS + i, (sgi * v) * isz_of(pb, ph, m, lcol));
        }
    }
}

// ---- the other rows: one item each ----------------------------------------------------------------------------
// item numbering (top of the thread range first): 0 objective | path rows (k,q) | event rows | duration row |
// linkage state rows towards the next phase | linkage triplets of the previous phase's rows in this phase
template <int M, bool FD>
ECUDA_HD int other_items(const ProbDev& pb, const PhaseDev& ph, int p) {
    return 1 + ph.npath * ph.N + pb.ne + 1 + (p + 1 < pb.nphases ? pb.ns : 0) + (p > 0 ? pb.ns : 0);
}

template <int M, int NB, bool FD>
ECUDA_HD void other_item(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m, int b, int it,
                         RowState<M, NB>& rs) {
    constexpr int NS = Model<M>::NS;
    const int N = ph.N, nc = pb.nc, np = ph.npath, ntr = np - ph.nstat;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double* sg = pb.sg;
    double* g = io.g ? io.g + static_cast<size_t>(b) * pb.ncons : nullptr;
    double* jac = io.jac ? io.jac + static_cast<size_t>(b) * pb.nnz : nullptr;
    const int tcol = (NS + nc) * N;
    auto note = [&](int r, double val, int cls) {
        if (io.nranks > 0) rs.viol = fmax(rs.viol, row_violation(io, pb, ph, m, b, r, val, cls));
    };

    if (it == 0) {  // ---- objective: running cost per node and quadrature            [phase_b + objective_phase]
        if (!io.f) return;
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
            const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
            const double L = Model<M>::cost(m.z + nc * N + k * NS, m.z + k * nc, t);
            acc = fma(ECUDA_LDG(ph.w + k), pb.maximize ? -1.0 * L : L, acc);
        }
        const double fp = pt.h * acc;
        if (pb.nphases == 1)
            io.f[b] = pb.sf * fp;
        else
            io.fpart[static_cast<size_t>(b) * pb.nphases + p] = fp;
        rs.fval = pb.sf * fp;
        return;
    }
    it -= 1;
    if (it < np * N) {  // ---- path row (k,q)                       [phase_b path rows, xcol_path_*, node_item tracks]
        int k, q;
        path_item(ph, it, k, q);
        const double tau = ECUDA_LDG(ph.tau + k);
        const double t = pt.h * tau + pt.m;
        const int lcol0 = nc * N + k * NS;
        const double x0 = m.z[lcol0], x1 = m.z[lcol0 + 1];
        const int r = ph.goff + NS * N + pb.ne + k * np + q;
        const double s = ECUDA_LDG(sg + r);
        const PathRowAt<M> row(pb, ph, m, q, t);  // value and position perturbations share a moving zone's centre
        if (g) {
            const double val = s * row(x0, x1);
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 1);
        }
        if (!jac) return;
        const int ev = (k == 0 || k == N - 1) ? 1 : 0;
        if (FD) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int lcol = lcol0 + j;
                const double xpv = m.xp[lcol], xmv = m.xm[lcol];
                const double vp = row(j == 0 ? xpv : x0, j == 1 ? xpv : x1);
                const double vm = row(j == 0 ? xmv : x0, j == 1 ? xmv : x1);
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + N - 1 + pb.xcnt[j] + ev + q, (s * vp - s * vm) * m.rinv[lcol]);
            }
            if (q >= ph.nstat) {
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int lcol = tcol + which;
                    const double t0p = which == 0 ? m.xp[lcol] : pt.t0, tfp = which == 1 ? m.xp[lcol] : pt.tf;
                    const double t0m = which == 0 ? m.xm[lcol] : pt.t0, tfm = which == 1 ? m.xm[lcol] : pt.tf;
                    const double hp = 0.5 * (tfp - t0p), mp = 0.5 * (tfp + t0p);
                    const double hm = 0.5 * (tfm - t0m), mm = 0.5 * (tfm + t0m);
                    const double tp = hp * tau + mp, tm = hm * tau + mm;
                    const double vp = path_row<M>(pb, ph, m, q, x0, x1, tp);
                    const double vm = path_row<M>(pb, ph, m, q, x0, x1, tm);
                    ECUDA_STREAM_STORE(jac + m.colp[lcol] + NS * N + k * ntr + (q - ph.nstat), (s * vp - s * vm) * m.rinv[lcol]);
                }
            }
        } else {
            double ddx, ddy, ddt = 0.0;
            if (q < ph.nstat)
                Model<M>::static_row_dxy(m.inst + ph.inst_off + q * Model<M>::REC, x0, x1, &ddx, &ddy);
            else
                moving_row_partials<M>(pb, ph, m, q, x0, x1, t, &ddx, &ddy, &ddt);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int lcol = lcol0 + j;
                const double v = (j == 0) ? ddx : ddy;
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + N - 1 + pb.xcnt[j] + ev + q, (s * v) * isz_of(pb, ph, m, lcol));
            }
            if (q >= ph.nstat) {
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int lcol = tcol + which;
                    const double dtk = which == 0 ? 0.5 * (1.0 - tau) : 0.5 * (1.0 + tau);
                    ECUDA_STREAM_STORE(jac + m.colp[lcol] + NS * N + k * ntr + (q - ph.nstat),
                                       (s * (ddt * dtk)) * isz_of(pb, ph, m, lcol));
                }
            }
        }
        return;
    }
    it -= np * N;
    if (it < pb.ne) {  // ---- event row: x(t0) or x(tf)                         [phase_b events, xcol_local_* event]
        const int e = it;
        const int node = (e < NS) ? 0 : N - 1, i = (e < NS) ? e : e - NS;
        const int r = ph.goff + NS * N + e;
        const int lcol = nc * N + node * NS + i;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * m.z[lcol];
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 2);
        }
        if (jac) {
            const int pos = N - 1 + pb.xcnt[i];
            if (FD)
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * m.xp[lcol] - s * m.xm[lcol]) * m.rinv[lcol]);
            else
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * 1.0) * isz_of(pb, ph, m, lcol));
        }
        return;
    }
    it -= pb.ne;
    if (it == 0) {  // ---- duration row tf - t0, and the time linkage             [phase_b, node_item k == 0 parts]
        const int r = ph.goff + NS * N + pb.ne + np * N;
        const double s = ECUDA_LDG(sg + r);
        if (g) {
            const double val = s * (pt.tf - pt.t0);
            ECUDA_STREAM_STORE(g + r, val);
            note(r, val, 3);
            if (p + 1 < pb.nphases) {  // time continuity with the next phase
                const PhaseDev& nx = pb.ph[p + 1];
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                const double other = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                ECUDA_STREAM_STORE(g + rl, ECUDA_LDG(sg + rl) * (pt.tf - other));
            }
        }
        if (!jac) return;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const int lcol = tcol + which;
            const int at = m.colp[lcol] + NS * N + N * ntr;
            if (FD) {
                const double ri = m.rinv[lcol];
                const double t0p = which == 0 ? m.xp[lcol] : pt.t0, tfp = which == 1 ? m.xp[lcol] : pt.tf;
                const double t0m = which == 0 ? m.xm[lcol] : pt.t0, tfm = which == 1 ? m.xm[lcol] : pt.tf;
                ECUDA_STREAM_STORE(jac + at, (s * (tfp - t0p) - s * (tfm - t0m)) * ri);
                if (which == 0 && p > 0) {
                    const PhaseDev& pv = pb.ph[p - 1];
                    const int rl = pb.linkoff + (p - 1) * (NS + 1) + NS;
                    const double sl = ECUDA_LDG(sg + rl);
                    const double o = other_phase_value(pb, io, b, pv.zoff + (NS + nc) * pv.N + 1);
                    ECUDA_STREAM_STORE(jac + at + 1, (sl * (o - t0p) - sl * (o - t0m)) * ri);
                }
                if (which == 1 && p + 1 < pb.nphases) {
                    const PhaseDev& nx = pb.ph[p + 1];
                    const int rl = pb.linkoff + p * (NS + 1) + NS;
                    const double sl = ECUDA_LDG(sg + rl);
                    const double o = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                    ECUDA_STREAM_STORE(jac + at + 1, (sl * (tfp - o) - sl * (tfm - o)) * ri);
                }
            } else {
                const double is = isz_of(pb, ph, m, lcol);
                ECUDA_STREAM_STORE(jac + at, (s * (which == 0 ? -1.0 : 1.0)) * is);
                if (which == 0 && p > 0) {
                    const int rl = pb.linkoff + (p - 1) * (NS + 1) + NS;
                    ECUDA_STREAM_STORE(jac + at + 1, (ECUDA_LDG(sg + rl) * -1.0) * is);
                }
                if (which == 1 && p + 1 < pb.nphases) {
                    const int rl = pb.linkoff + p * (NS + 1) + NS;
                    ECUDA_STREAM_STORE(jac + at + 1, (ECUDA_LDG(sg + rl) * 1.0) * is);
                }
            }
        }
        return;
    }
    it -= 1;
    // ---- state linkage with the next phase (row owned by this phase) / with the previous phase (triplet
    // of its row in this phase's first node)                                      [phase_b, xcol_local_* linkage]
    const bool to_next = (p + 1 < pb.nphases) && it < NS;
    const int i = to_next ? it : it - (p + 1 < pb.nphases ? NS : 0);
    const int k = to_next ? N - 1 : 0;
    const int lcol = nc * N + k * NS + i;
    int pos = N - 1 + pb.xcnt[i] + 1;  // after the event triplet of a boundary node
    if (i < 2) pos += np;
    if (to_next) {
        const PhaseDev& nx = pb.ph[p + 1];
        const int r = pb.linkoff + p * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        const double o = other_phase_value(pb, io, b, nx.zoff + nc * nx.N + i);
        if (g) ECUDA_STREAM_STORE(g + r, s * (m.z[lcol] - o));
        if (jac) {
            if (FD)
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * (m.xp[lcol] - o) - s * (m.xm[lcol] - o)) * m.rinv[lcol]);
            else
                ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * 1.0) * isz_of(pb, ph, m, lcol));
        }
    } else if (jac) {
        const PhaseDev& pv = pb.ph[p - 1];
        const int r = pb.linkoff + (p - 1) * (NS + 1) + i;
        const double s = ECUDA_LDG(sg + r);
        if (FD) {
            const double o = other_phase_value(pb, io, b, pv.zoff + nc * pv.N + (pv.N - 1) * NS + i);
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * (o - m.xp[lcol]) - s * (o - m.xm[lcol])) * m.rinv[lcol]);
        } else {
            ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * -1.0) * isz_of(pb, ph, m, lcol));
        }
    }
}

template <int M, int NB, bool FD>
ECUDA_HD void rows_other(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m, int b, int tid,
                         int nthr, RowState<M, NB>& rs, bool values, bool triplets) {
    // `values` / `triplets` select what an item writes, so that exact mode can put the CTA-wide barrier
    // that waits for the template copy between the g values and the triplets
    EvalIO part = io;
    if (!values) part.g = nullptr, part.f = nullptr;
    if (!triplets) part.jac = nullptr;
    const int nitems = other_items<M, FD>(pb, ph, p);
    for (int it = nthr - 1 - tid; it < nitems; it += nthr) other_item<M, NB, FD>(pb, ph, p, part, m, b, it, rs);
}

}  // namespace ecuda
#endif
