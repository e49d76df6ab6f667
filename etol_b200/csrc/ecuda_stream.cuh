// ecuda_stream.cuh -- exact-Jacobian evaluation as a STREAM: values first, then every triplet in address order.
//
// derivatives = "automatic" is the reference's default (src/ePSOPT/ePSOPT.cpp:64). Every exact triplet of the
// collocation NLP has the same form,   J[e] = (sg[row(e)] * T(e)) / sz[col(e)],   where T(e) is either an entry of
// the differentiation matrix D (the 74 % "D-coupled" triplets: defect row (k,i) x state column X(l,i), l != k) or
// one of a few hundred per-instance numbers (node-local partials of the defects, partials of the path rows, +-1 for
// event / duration / linkage rows). The host builds, with the sparsity pattern, one 64-bit descriptor per triplet
// {which T, row, column} (ecuda_host.cpp: build_structure; desc_pack in ecuda_internal.hpp). The kernel then is
//   phase 1  one thread per defect row / per other row: dynamics, their partials, D X, path rows -> the instance's
//            table T and its scaled constraint values, both in shared memory;
//   barrier
//   phase 2  all 256 threads walk the instance's triplet range IN ADDRESS ORDER: descriptor (coalesced 8-byte loads,
//            L2 resident, shared by the whole batch), two multiplications, one coalesced streaming store -- every warp
//            instruction writes 256 contiguous bytes, the CTA writes its 101 KB front to back. g leaves the same way.
// No per-row scatter (a store-only kernel with the row-owner pattern reaches 2.2 TB/s on a B200, an address-ordered
// one 6 TB/s: scripts/wroof.cu), one barrier after staging, no second pass over the triplets.
// Same arithmetic as xcol_local_exact / node_item / xcol_path_exact of ecuda_phases.cuh: bit-identical results.
#ifndef ECUDA_STREAM_CUH_
#define ECUDA_STREAM_CUH_

#include "ecuda_rowsn.cuh"

namespace ecuda {

struct StMem {
    double* inst;  // [inst_stride] obstacle / track records (bulk-copied)
    double* z;     // [nv]   unscaled variables of the phase
    double* isz;   // [nv]   1 / sz
    double* tab;   // [desc_table_size] per-instance table T
    double* gbuf;  // [phase_ncons]     scaled constraint values of the phase
    double* sg;    // [ncons]           row scales (all rows: a phase's triplets also sit in linkage rows)
};

template <int M>
ECUDA_HD size_t st_doubles(const ProbDev& pb, const PhaseDev& ph, int N) {
    const size_t nv = static_cast<size_t>(rn_nv<M>(pb, N)), nve = nv + (nv & 1);
    const size_t nt = static_cast<size_t>(desc_table_size(Model<M>::NS, pb.nc, N, ph.npath));
    const size_t ng = static_cast<size_t>(phase_ncons(pb, ph));
    const size_t nr = static_cast<size_t>(pb.ncons);
    return static_cast<size_t>(pb.inst_stride) + 2 * nve + (nt + (nt & 1)) + (ng + (ng & 1)) + (nr + (nr & 1));
}
template <int M>
ECUDA_HD void st_carve(StMem& m, double* base, const ProbDev& pb, const PhaseDev& ph, int N) {
    const size_t nv = static_cast<size_t>(rn_nv<M>(pb, N)), nve = nv + (nv & 1);
    const size_t nt = static_cast<size_t>(desc_table_size(Model<M>::NS, pb.nc, N, ph.npath));
    m.inst = base;
    base += pb.inst_stride;
    m.z = base;
    base += nve;
    m.isz = base;
    base += nve;
    m.tab = base;
    base += nt + (nt & 1);
    m.gbuf = base;
    const size_t ng = static_cast<size_t>(phase_ncons(pb, ph));
    base += ng + (ng & 1);
    m.sg = base;
}

// stage: z = z~ / sz and 1 / sz
template <int M, int N>
ECUDA_HD void st_stage(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const StMem& m, int b, int tid, int nthr) {
    const int nv = rn_nv<M>(pb, N);
    const double* xs = io.x + static_cast<size_t>(b) * pb.nvars + ph.zoff;
    const double* is = pb.isz + ph.zoff;
    for (int c0 = tid; c0 < nv; c0 += 2 * nthr) {
        double zt[2], s[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (c0 + u * nthr < nv) {
                zt[u] = ECUDA_LDG(xs + c0 + u * nthr);
                s[u] = ECUDA_LDG(is + c0 + u * nthr);
            }
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (c0 + u * nthr < nv) {
                m.z[c0 + u * nthr] = zt[u] * s[u];
                m.isz[c0 + u * nthr] = s[u];
            }
    }
    if (tid == 0) {
        m.tab[0] = 1.0;
        m.tab[1] = -1.0;
    }
    // D^T of the phase behind the per-instance entries: phase 2 reads every T from one shared-memory table
    double* td = m.tab + desc_d_off(Model<M>::NS, pb.nc, N, ph.npath);
    if (io.jac) {
        for (int e = tid; e < N * N; e += nthr) td[e] = ECUDA_LDG(ph.Dt + e);
        for (int r = tid; r < pb.ncons; r += nthr) m.sg[r] = ECUDA_LDG(pb.sg + r);
    }
}

// phase 1, defect row (k,i): value and the row of the table           [rows_values + rows_jacobian<exact>]
template <int M, int N, bool SUM>
ECUDA_HD void st_row(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const StMem& m, const CtaMem& cm, int b, int tid,
                     double& viol) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU, NB = (N + 7) / 8;
    if (tid >= NS * N) return;
    const int nc = pb.nc;
    const int k = tid / NS, i = tid - k * NS;  // node-major: the table rows and g values are written contiguously
    const double* zx = m.z + nc * N;
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
    const double* Dtk = ph.Dt + k;
    if (io.g) {
        const int r = ph.goff + k * NS + i;
        double P[NB];
        const double dv = rn_dot<NS, N>(Dtk, zx + i, P);
        const double val = ECUDA_LDG(pb.sg + r) * (dv - h * fi);
        m.gbuf[k * NS + i] = val;
        if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, 0));
    }
    if (!io.jac) return;
    double* row = m.tab + 2 + (k * NS + i) * desc_rowtab_width(NS, nc);
    double dfdx[NS][NS], dfdu[NS][NCU];
    static_assert(!Model<M>::TDEP, "k_stream_exact is instantiated for the built-in (autonomous) models only");
    Model<M>::jac(x, u, t, dfdx, dfdu);
    const double dkk = ECUDA_LDG(Dtk + k * N);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
            if (a == i) d = dfdx[a][j];
        row[j] = ((i == j) ? dkk : 0.0) - h * d;
    }
    for (int c = 0; c < nc; ++c) {
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int c2 = 0; c2 < NCU; ++c2)
                if (a == i && c2 == c) d = dfdu[a][c2];
        row[NS + c] = -(h * d);
    }
    row[NS + nc] = 0.5 * fi;
    row[NS + nc + 1] = -0.5 * fi;
}

// phase 1, the other rows: one item each (numbering of rn_items)              [other_item<exact>, values and partials]
template <int M, int N, bool TRK, bool SUM>
ECUDA_HD void st_item(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const StMem& m, const CtaMem& cm, int b,
                      int it, double& viol, double& fval) {
    constexpr int NS = Model<M>::NS;
    const int nc = pb.nc, np = ph.npath;
    const double* zx = m.z + nc * N;
    const double t0 = m.z[(NS + nc) * N], tf = m.z[(NS + nc) * N + 1];
    const double h = 0.5 * (tf - t0), mid = 0.5 * (tf + t0);
    const double* sg = pb.sg;
    auto note = [&](int r, double val, int cls) {
        if (SUM) viol = fmax(viol, row_violation(io, pb, ph, cm, b, r, val, cls));
    };
    if (it == 0) {  // ---- objective
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
    if (it < np * N) {  // ---- path row (k,q): value, partials
        int k, q;
        path_item(ph, it, k, q);
        const double tau = ECUDA_LDG(ph.tau + k);
        const double t = h * tau + mid;
        const double x0 = zx[k * NS], x1 = zx[k * NS + 1];
        const int lr = NS * N + pb.ne + k * np + q, r = ph.goff + lr;
        if (io.g) {
            const double val = ECUDA_LDG(sg + r) * rn_path_row<M, TRK>(pb, ph, cm, q, x0, x1, t);
            m.gbuf[lr] = val;
            note(r, val, 1);
        }
        if (!io.jac) return;
        double ddx, ddy, ddt = 0.0;
        if (!TRK || q < ph.nstat)
            Model<M>::static_row_dxy(cm.inst + ph.inst_off + q * Model<M>::REC, x0, x1, &ddx, &ddy);
        else
            track_row_partials(cm.inst + pb.track_off + (q - ph.nstat) * pb.track_size, pb.nway, x0, x1, t, &ddx, &ddy, &ddt);
        double* e = m.tab + desc_path_off(NS, nc, N) + (k * np + q) * 4;
        e[0] = ddx;
        e[1] = ddy;
        if (TRK) {
            e[2] = ddt * (0.5 * (1.0 - tau));
            e[3] = ddt * (0.5 * (1.0 + tau));
        }
        return;
    }
    it -= np * N;
    if (it < pb.ne) {  // ---- event row
        const int e = it;
        const int node = (e < NS) ? 0 : N - 1, i = (e < NS) ? e : e - NS;
        const int lr = NS * N + e, r = ph.goff + lr;
        if (io.g) {
            const double val = ECUDA_LDG(sg + r) * zx[node * NS + i];
            m.gbuf[lr] = val;
            note(r, val, 2);
        }
        return;
    }
    it -= pb.ne;
    if (it == 0) {  // ---- duration row tf - t0, and the time linkage
        const int lr = NS * N + pb.ne + np * N, r = ph.goff + lr;
        if (io.g) {
            const double val = ECUDA_LDG(sg + r) * (tf - t0);
            m.gbuf[lr] = val;
            note(r, val, 3);
            if (p + 1 < pb.nphases) {
                const PhaseDev& nx = pb.ph[p + 1];
                const int rl = pb.linkoff + p * (NS + 1) + NS;
                const double other = other_phase_value(pb, io, b, nx.zoff + (NS + nc) * nx.N);
                ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + rl, ECUDA_LDG(sg + rl) * (tf - other));
            }
        }
        return;
    }
    it -= 1;
    // ---- state linkage with the next phase: the row is owned by this phase (its triplets are +-1 entries)
    if ((p + 1 < pb.nphases) && it < NS && io.g) {
        const int i = it;
        const PhaseDev& nx = pb.ph[p + 1];
        const int r = pb.linkoff + p * (NS + 1) + i;
        const double o = other_phase_value(pb, io, b, nx.zoff + nc * nx.N + i);
        ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, ECUDA_LDG(sg + r) * (zx[(N - 1) * NS + i] - o));
    }
}

template <int M, int N, bool TRK, bool SUM>
ECUDA_HD void st_phase1(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const StMem& m, const CtaMem& cm, int b,
                        int tid, int nthr, double& viol, double& fval) {
    viol = 0.0;
    fval = 0.0;
    st_row<M, N, SUM>(pb, ph, io, m, cm, b, tid, viol);
    const int nitems = rn_items<M, N>(pb, ph, p);
    for (int it = nthr - 1 - tid; it < nitems; it += nthr) st_item<M, N, TRK, SUM>(pb, ph, p, io, m, cm, b, it, viol, fval);
}

// phase 2: the triplets of the phase in address order, then its constraint values. Per triplet: descriptor (one
// coalesced 64-bit load), T and 1/sz from shared memory, sg[row] from L1, two multiplications, one streaming store.
template <int M, int N>
ECUDA_HD void st_phase2(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const StMem& m, int b, int tid, int nthr) {
    constexpr int UNR = 8;
    const int nv = rn_nv<M>(pb, N);
    if (io.jac) {
        const int e0 = ECUDA_LDG(pb.colptr + ph.zoff), e1 = ECUDA_LDG(pb.colptr + ph.zoff + nv);
        double* __restrict__ jac = io.jac + static_cast<size_t>(b) * pb.nnz + e0;
        const unsigned long long* __restrict__ desc = pb.desc + e0;
        const int n = e1 - e0, nfull = n - n % (UNR * nthr);
        // Software pipeline: the descriptors of the next batch are in flight while this one is computed and stored --
        // the descriptor read is the only global load of the loop, and under the kernel's own write traffic a load
        // takes thousands of cycles.
        unsigned long long d[UNR], dn[UNR];
        int e = tid;
        if (e < nfull) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) d[u] = ECUDA_LDG(desc + e + u * nthr);
        }
        for (; e < nfull; e += UNR * nthr) {
            const bool more = e + UNR * nthr < nfull;
            if (more) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) dn[u] = ECUDA_LDG(desc + e + (UNR + u) * nthr);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const unsigned hi = static_cast<unsigned>(d[u] >> 32);
                const double v = (m.sg[hi & 0xffffu] * m.tab[static_cast<unsigned>(d[u])]) * m.isz[hi >> 16];
                ECUDA_STREAM_STORE(jac + e + u * nthr, v);
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) d[u] = dn[u];
            }
        }
        for (; e < n; e += nthr) {
            const unsigned long long dd = ECUDA_LDG(desc + e);
            const unsigned hi = static_cast<unsigned>(dd >> 32);
            ECUDA_STREAM_STORE(jac + e, (m.sg[hi & 0xffffu] * m.tab[static_cast<unsigned>(dd)]) * m.isz[hi >> 16]);
        }
    }
    if (io.g) {
        const int ncp = phase_ncons(pb, ph);
        double* g = io.g + static_cast<size_t>(b) * pb.ncons + ph.goff;
        for (int r = tid; r < ncp; r += nthr) ECUDA_STREAM_STORE(g + r, m.gbuf[r]);
    }
}

}  // namespace ecuda
#endif
