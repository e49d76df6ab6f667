// ecuda_phases.cuh -- the per-instance evaluation of the collocation NLP, written as barrier-free
// "phases" over the threads of one CTA. One CTA evaluates one (instance, phase) pair:
//
//   stage     z~ -> unscaled z in shared memory; per-column FD data (z+-delta, 1/(2 delta));
//             the instance's obstacle/track records (bulk-copied by the kernel)
//   phase B   per node: state derivatives f, h*f, running cost, path rows (K1 + K3);
//             per (row k, state j): the blocked dot product  sum_l D[k][l] X[l][j]  (K2);
//             event / duration / linkage rows
//   phase C   objective quadrature (K0); defect rows; Jacobian values written straight into the
//             (col,row)-sorted triplet array (K4), either exact or by index-set central differences
//
// Reference counterparts: PSOPT's defect assembly and quadrature entered through ePSOPT::solve
// (src/ePSOPT/ePSOPT.cpp:84) with the callbacks of src/ePSOPT/ePSOPT.cpp:186-306, and its
// derivatives="automatic"/"numerical" Jacobian drivers (mode chosen at ePSOPT.cpp:64).
//
// Finite differences are evaluated *row-restricted*: perturbing column c only changes the rows in
// c's sparsity pattern, and each of those rows is recomputed with exactly the operation sequence a
// full re-evaluation of g(z +- delta e_c) would use, so the values equal the column-grouped
// (Curtis-Powell-Reid) scheme bit for bit while doing O(nnz) instead of O(groups * ncons) work.
// The blocked summation of D*X (ECUDA_DOT_BLOCK) is what makes the re-evaluation of a defect row
// after a one-element change cost one block instead of the whole dot product.
//
// All functions are __host__ __device__: tests/emu steps the same source on the CPU so that index
// logic can be checked where there is no GPU. The product only ever runs the __global__ kernels.
#ifndef ECUDA_PHASES_CUH_
#define ECUDA_PHASES_CUH_

#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#include "ecuda_models.cuh"

namespace ecuda {

#if defined(__CUDA_ARCH__)
#define ECUDA_LDG(p) __ldg(p)
// result stores. ECUDA_STORE_MODE 0: st.global.cs (streaming, evict first); 1: plain write-back stores (lines stay in
// L2 until they are complete, so the node-local triplets written later merge with the D-coupled ones of the same
// sectors before anything goes to HBM); 2: st.global.cg
#ifndef ECUDA_STORE_MODE
#define ECUDA_STORE_MODE 0
#endif
#if ECUDA_STORE_MODE == 1
#define ECUDA_STREAM_STORE(p, v) (*(p) = (v))
#elif ECUDA_STORE_MODE == 2
#define ECUDA_STREAM_STORE(p, v) __stcg((p), (v))
#else
#define ECUDA_STREAM_STORE(p, v) __stcs((p), (v))
#endif
#else
#define ECUDA_LDG(p) (*(p))
#define ECUDA_STREAM_STORE(p, v) (*(p) = (v))
#endif

#define ECUDA_SQRT_EPS 1.4901161193847656e-08 /* 2^-26 */

// Triplet sink. SM = false: streaming store to the caller's global array. SM = true: plain store into
// the CTA's shared-memory image of the instance's triplet range (k_eval_image), which leaves through
// one bulk copy.
template <bool SM>
ECUDA_HD void jstore(double* p, double v) {
    if (SM)
        *p = v;
    else
        ECUDA_STREAM_STORE(p, v);
}

struct CtaMem {
    double* z;     // [nvars_p] unscaled variables of this phase
    double* xp;    // [nvars_p] (z~ + delta) * isz
    double* xm;    // [nvars_p] (z~ - delta) * isz
    double* rinv;  // [nvars_p] 1 / (2 delta)
    double* hf;    // [N*ns]    h * f_i(node k)
    double* dotv;  // [N*ns]    (D X)[k][i]
    double* Lk;    // [N]       running cost per node
    double* inst;  // [inst_stride] obstacle + track records of the instance
    double* P;     // [nb][nthr] thread-private block sums (generic block count only)
    int* colp;     // [nvars_p + 1] first triplet index of each column of this phase (+ end of the phase)
    const double* bl;  // fused summary only: this phase's block of the instance's constraint bounds,
    const double* bu;  //   staged in shared memory by the kernel (null: read them from global memory)
};

// what a CTA keeps in shared memory: the generic kernels need everything; the fast kernels keep the
// block sums in registers (no P) and, in exact mode, have no finite-difference data (no xp/xm/rinv)
enum { CARVE_P = 1, CARVE_FD = 2, CARVE_ALL = 3, CARVE_ISZ = 4 };
#define ECUDA_STAGE_BATCH 2  // strides of the staging loops whose loads are in flight together

// shared-memory footprint in doubles for one CTA working on phase `ph`
ECUDA_HD size_t cta_doubles(const ProbDev& pb, const PhaseDev& ph, int nthr, int what = CARVE_ALL) {
    size_t n = 0;
    n += ((what & CARVE_FD) ? 4 : 1) * static_cast<size_t>(ph.nvars + (ph.nvars & 1));
    n += 2 * static_cast<size_t>(ph.N) * pb.ns;
    n += static_cast<size_t>(ph.N + (ph.N & 1));
    n += static_cast<size_t>(pb.inst_stride);
    if (what & CARVE_P) n += static_cast<size_t>(ph.nb) * nthr;
    if (what & CARVE_ISZ) n += static_cast<size_t>(ph.nvars + (ph.nvars & 1));
    n += static_cast<size_t>((ph.nvars + 3) / 2);  // colp (ints), nvars_p + 1 of them
    return n;
}

ECUDA_HD void carve(CtaMem& m, double* base, const ProbDev& pb, const PhaseDev& ph, int nthr, int what = CARVE_ALL) {
    size_t nv = static_cast<size_t>(ph.nvars + (ph.nvars & 1));
    m.inst = base;  // first: 16-byte aligned destination of the bulk copy
    base += pb.inst_stride;
    m.z = base;     base += nv;
    m.xp = m.xm = m.rinv = m.P = nullptr;
    m.bl = m.bu = nullptr;
    if (what & CARVE_FD) {
        m.xp = base;    base += nv;
        m.xm = base;    base += nv;
        m.rinv = base;  base += nv;
    }
    if ((what & CARVE_ISZ) && !(what & CARVE_FD)) {  // exact row-owner kernel: the rinv slot holds 1/sz instead
        m.rinv = base;
        base += nv;
    }
    m.hf = base;    base += static_cast<size_t>(ph.N) * pb.ns;
    m.dotv = base;  base += static_cast<size_t>(ph.N) * pb.ns;
    m.Lk = base;    base += static_cast<size_t>(ph.N + (ph.N & 1));
    if (what & CARVE_P) {
        m.P = base;
        base += static_cast<size_t>(ph.nb) * nthr;
    }
    m.colp = reinterpret_cast<int*>(base);
}

struct PhaseTimes {
    double t0, tf, h, m;
};
ECUDA_HD PhaseTimes phase_times(const ProbDev& pb, const PhaseDev& ph, const double* z) {
    PhaseTimes pt;
    pt.t0 = z[(pb.ns + pb.nc) * ph.N];
    pt.tf = z[(pb.ns + pb.nc) * ph.N + 1];
    pt.h = 0.5 * (pt.tf - pt.t0);
    pt.m = 0.5 * (pt.tf + pt.t0);
    return pt;
}

// ---- stage ------------------------------------------------------------------------------------------
// ISZ: also keep 1/sz in shared memory (in m.rinv, which exact mode does not otherwise use; CARVE_ISZ) -- compile
// time, so that kernels without it pay nothing
template <bool ISZ = false>
ECUDA_HD void stage_vars(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, CtaMem& m, int b, int tid,
                         int nthr, bool fd) {
    const double* xs = io.x + static_cast<size_t>(b) * pb.nvars + ph.zoff;
    const double* is = pb.isz + ph.zoff;
    if (tid == 0) m.colp[ph.nvars] = ECUDA_LDG(pb.colptr + ph.zoff + ph.nvars);
    // all global loads of a batch of ECUDA_STAGE_BATCH strides are issued before the first dependent
    // store, so that a thread pays the DRAM latency once per batch instead of once per element
    for (int c0 = tid; c0 < ph.nvars; c0 += ECUDA_STAGE_BATCH * nthr) {
        double zt[ECUDA_STAGE_BATCH], s[ECUDA_STAGE_BATCH];
        int cp[ECUDA_STAGE_BATCH];
#pragma unroll
        for (int u = 0; u < ECUDA_STAGE_BATCH; ++u) {
            const int c = c0 + u * nthr;
            if (c < ph.nvars) {
                zt[u] = ECUDA_LDG(xs + c);
                s[u] = ECUDA_LDG(is + c);
                cp[u] = ECUDA_LDG(pb.colptr + ph.zoff + c);
            }
        }
#pragma unroll
        for (int u = 0; u < ECUDA_STAGE_BATCH; ++u) {
            const int c = c0 + u * nthr;
            if (c < ph.nvars) {
                m.z[c] = zt[u] * s[u];
                m.colp[c] = cp[u];
                if (ISZ) m.rinv[c] = s[u];
                if (fd) {
                    double delta = ECUDA_SQRT_EPS * (1.0 + fabs(zt[u]));
                    m.xp[c] = (zt[u] + delta) * s[u];
                    m.xm[c] = (zt[u] - delta) * s[u];
                    m.rinv[c] = 1.0 / (2.0 * delta);
                }
            }
        }
    }
}

// ---- small helpers ----------------------------------------------------------------------------------
template <int M>
ECUDA_HD double path_row(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int q, double x, double y,
                         double t) {
    if (q < ph.nstat) return Model<M>::static_row(m.inst + ph.inst_off + q * Model<M>::REC, x, y);
    if constexpr (Model<M>::NUSER > 0) {  // traced path rows of a user model follow the moving-zone rows
        if (q - ph.nstat >= pb.ntracks) return Model<M>::user_row(q - ph.nstat - pb.ntracks, x, y, t);
    }
    return track_row(m.inst + pb.track_off + (q - ph.nstat) * pb.track_size, pb.nway, x, y, t);
}
// path row q at ONE time and several positions (value and the x / y finite differences of a row): a moving-zone row
// computes its centre -- two divisions and a waypoint search -- once; same operations as path_row, hence the same bits
template <int M>
struct PathRowAt {
    const ProbDev& pb;
    const PhaseDev& ph;
    const CtaMem& m;
    int q;
    double t, xc, yc;
    bool trk;
    const double* rec;
    ECUDA_HD PathRowAt(const ProbDev& pb_, const PhaseDev& ph_, const CtaMem& m_, int q_, double t_)
        : pb(pb_), ph(ph_), m(m_), q(q_), t(t_), xc(0.0), yc(0.0) {
        trk = q >= ph.nstat && q - ph.nstat < pb.ntracks;
        rec = m.inst + pb.track_off + (trk ? q - ph.nstat : 0) * pb.track_size;
        if (trk) track_center(rec, pb.nway, t, &xc, &yc);
    }
    ECUDA_HD double operator()(double x, double y) const {
        return trk ? track_row_at(rec, xc, yc, x, y) : path_row<M>(pb, ph, m, q, x, y, t);
    }
};
// partials of path row q >= nstat (a moving zone, or a traced row of a user model) in x_0, x_1 and t
template <int M>
ECUDA_HD void moving_row_partials(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int q, double x, double y,
                                  double t, double* ddx, double* ddy, double* ddt) {
    if constexpr (Model<M>::NUSER > 0) {
        if (q - ph.nstat >= pb.ntracks) {
            Model<M>::user_row_partials(q - ph.nstat - pb.ntracks, x, y, t, ddx, ddy, ddt);
            return;
        }
    }
    track_row_partials(m.inst + pb.track_off + (q - ph.nstat) * pb.track_size, pb.nway, x, y, t, ddx, ddy, ddt);
}

// canonical blocked dot of row k of D with state j of X: blocks of ECUDA_DOT_BLOCK nodes, each a
// serial ascending fma chain from 0; block sums added serially in ascending order. The block sums
// are also left in P[bi*pstride] (thread-private shared memory) for the finite-difference pass.
ECUDA_HD double dot_row(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int k, int j, double* P,
                        int pstride) {
    const double* X = m.z + pb.nc * ph.N + j;
    const double* Dt = ph.Dt + k;
    const int N = ph.N, ns = pb.ns;
    double total = 0.0;
    for (int l0 = 0, bi = 0; l0 < N; l0 += ECUDA_DOT_BLOCK, ++bi) {
        double p = 0.0;
        if (l0 + ECUDA_DOT_BLOCK <= N) {
#pragma unroll
            for (int i = 0; i < ECUDA_DOT_BLOCK; ++i)
                p = fma(ECUDA_LDG(Dt + static_cast<size_t>(l0 + i) * N), X[(l0 + i) * ns], p);
        } else {
            for (int l = l0; l < N; ++l) p = fma(ECUDA_LDG(Dt + static_cast<size_t>(l) * N), X[l * ns], p);
        }
        P[bi * pstride] = p;
        total = (l0 == 0) ? p : total + p;
    }
    return total;
}

// unscaled value of a variable of another phase of the same instance (linkage rows)
ECUDA_HD double other_phase_value(const ProbDev& pb, const EvalIO& io, int b, int gcol) {
    return ECUDA_LDG(io.x + static_cast<size_t>(b) * pb.nvars + gcol) * ECUDA_LDG(pb.isz + gcol);
}

// Generic kernels: a phase may be split over `nslices` CTAs (one instance whose phase has more defect
// rows than a CTA has threads, e.g. the 200-node fixed-wing problem). Slice s owns the nodes
// k = s, s + nslices, ...: their defect rows, path rows and node-local triplets; slice 0 also owns the
// objective and the event, duration and linkage rows. Every slice stages the whole decision vector.
ECUDA_HD int slice_nodes(int N, int slice, int nslices) { return (N - slice + nslices - 1) / nslices; }
ECUDA_HD int generic_slices(int ns, int N, int nthr) {
    const int s = (ns * N + nthr - 1) / nthr;
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}

// ---- phase B: node evaluations, dots, event/duration/linkage rows --------------------------------------
template <int M>
ECUDA_HD void phase_b(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int tid,
                      int nthr, int slice = 0, int nslices = 1) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, ns = pb.ns, nc = pb.nc;
    const int nown = slice_nodes(N, slice, nslices);
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double* g = io.g ? io.g + static_cast<size_t>(b) * pb.ncons : nullptr;
    const double* sg = pb.sg;
    for (int k = tid; k < N; k += nthr) {
        const double* x = m.z + nc * N + k * ns;
        const double* u = m.z + k * nc;
        double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
        double f[NS];
        Model<M>::f(x, u, t, f);
#pragma unroll
        for (int i = 0; i < NS; ++i) m.hf[k * ns + i] = pt.h * f[i];
        double L = Model<M>::cost(x, u, t);
        m.Lk[k] = pb.maximize ? -1.0 * L : L;
        if (g && k % nslices == slice) {  // path rows of the slice's own nodes
            const int r0 = ph.goff + ns * N + pb.ne + k * ph.npath;
            for (int q = 0; q < ph.npath; ++q)
                ECUDA_STREAM_STORE(g + r0 + q, ECUDA_LDG(sg + r0 + q) * path_row<M>(pb, ph, m, q, x[0], x[1], t));
        }
        (void)NCU;
    }
    for (int it = tid; it < ns * nown; it += nthr) {  // defect rows of the slice's own nodes
        int j = it / nown, k = slice + (it - j * nown) * nslices;
        m.dotv[k * ns + j] = dot_row(pb, ph, m, k, j, m.P + tid, nthr);
    }
    if (g && slice == 0) {
        for (int e = tid; e < pb.ne; e += nthr) {
            int r = ph.goff + ns * N + e;
            int node = (e < ns) ? 0 : N - 1;
            int i = (e < ns) ? e : e - ns;
            ECUDA_STREAM_STORE(g + r, ECUDA_LDG(sg + r) * m.z[nc * N + node * ns + i]);
        }
        if (tid == 0) {
            int r = ph.goff + ns * N + pb.ne + ph.npath * N;
            ECUDA_STREAM_STORE(g + r, ECUDA_LDG(sg + r) * (pt.tf - pt.t0));
        }
        if (p + 1 < pb.nphases) {
            const PhaseDev& nx = pb.ph[p + 1];
            for (int i = tid; i <= ns; i += nthr) {
                int r = pb.linkoff + p * (ns + 1) + i;
                double mine = (i < ns) ? m.z[nc * N + (N - 1) * ns + i] : pt.tf;
                int ocol = (i < ns) ? nx.zoff + nc * nx.N + i : nx.zoff + (ns + nc) * nx.N;
                double other = other_phase_value(pb, io, b, ocol);
                ECUDA_STREAM_STORE(g + r, ECUDA_LDG(sg + r) * (mine - other));
            }
        }
    }
}

// ---- phase C ------------------------------------------------------------------------------------------
// objective of this phase: h * sum_k w_k L_k, serial ascending fma chain
ECUDA_HD double objective_phase(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m,
                                int b) {
    if (!io.f) return 0.0;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double acc = 0.0;
    for (int k = 0; k < ph.N; ++k) acc = fma(ECUDA_LDG(ph.w + k), m.Lk[k], acc);
    double fp = pt.h * acc;
    if (pb.nphases == 1)
        io.f[b] = pb.sf * fp;
    else
        io.fpart[static_cast<size_t>(b) * pb.nphases + p] = fp;
    return pb.sf * fp;  // the instance's scaled objective when there is one phase
}

// ---- Jacobian work items ------------------------------------------------------------------------------
// A "state item" (j,k) owns defect row (k,j): it writes that row of g, the D-coupled entries of the
// row (columns X(l,j), l != k) and the whole node-local part of column X(k,j) (defect rows of node
// k, event row, obstacle rows, linkage row). A "node item" (k,c) owns the node-local part of a
// control column or of the t0/tf columns at node k.

// position of defect row (k, state j) inside column X(l, j), k != l
ECUDA_HD int dot_entry_pos(const ProbDev& pb, int j, int k, int l) { return k < l ? k : k - 1 + pb.xcnt[j]; }

// ---- node-local part of column X(k,j) -------------------------------------------------------------------
// Column X(k,j) holds, in row order: the D-coupled defect rows (k',j), k' < k | the defect rows of
// node k | the D-coupled rows k' > k | the event row (k = 0 or N-1) | the path rows of node k
// (j < 2: every obstacle row reads the two horizontal positions) | the linkage row.
// `jac` is the base such that jac[e] is the slot of triplet e of this instance.

// defect rows of node k, event row, linkage row -- by central differences. dpk/dmk = (D X)[k][j]
// with X[k][j] replaced by its +/- perturbed value.
template <int M, bool SKIPDIAG = false>
ECUDA_HD void xcol_local_fd(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m, int b,
                            int j, int k, double dpk, double dmk, double* jac) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, ns = pb.ns, nc = pb.nc, np = ph.npath;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
    const double* sg = pb.sg;
    const int lcol = nc * N + k * ns + j;
    const int base = m.colp[lcol];
    const double xpv = m.xp[lcol], xmv = m.xm[lcol], ri = m.rinv[lcol];
    double xq[NS], xr[NS], u[NCU], fp[NS], fm[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double xi = m.z[nc * N + k * ns + i];
        xq[i] = (i == j) ? xpv : xi;
        xr[i] = (i == j) ? xmv : xi;
    }
#pragma unroll
    for (int i = 0; i < NCU; ++i) u[i] = m.z[k * nc + i];
    Model<M>::f(xq, u, t, fp);
    Model<M>::f(xr, u, t, fm);
    const int rdef0 = ph.goff + k * ns;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        int rk = pb.xrank[j][i];
        if (rk >= 0 && !(SKIPDIAG && i == j)) {  // SKIPDIAG: the caller has written the diagonal triplet
            double s = ECUDA_LDG(sg + rdef0 + i);
            double dv = m.dotv[k * ns + i];
            double gp = s * (((i == j) ? dpk : dv) - pt.h * fp[i]);
            double gm = s * (((i == j) ? dmk : dv) - pt.h * fm[i]);
            ECUDA_STREAM_STORE(jac + base + k + rk, (gp - gm) * ri);
        }
    }
    int pos = N - 1 + pb.xcnt[j];
    if (k == 0 || k == N - 1) {
        int r = ph.goff + ns * N + (k == 0 ? j : ns + j);
        double s = ECUDA_LDG(sg + r);
        ECUDA_STREAM_STORE(jac + base + pos, (s * xpv - s * xmv) * ri);
        ++pos;
    }
    if (j < 2) pos += np;
    if (k == N - 1 && p + 1 < pb.nphases) {
        const PhaseDev& nx = pb.ph[p + 1];
        int r = pb.linkoff + p * (ns + 1) + j;
        double s = ECUDA_LDG(sg + r);
        double o = other_phase_value(pb, io, b, nx.zoff + nc * nx.N + j);
        ECUDA_STREAM_STORE(jac + base + pos, (s * (xpv - o) - s * (xmv - o)) * ri);
    }
    if (k == 0 && p > 0) {
        const PhaseDev& pv = pb.ph[p - 1];
        int r = pb.linkoff + (p - 1) * (ns + 1) + j;
        double s = ECUDA_LDG(sg + r);
        double o = other_phase_value(pb, io, b, pv.zoff + nc * pv.N + (pv.N - 1) * ns + j);
        ECUDA_STREAM_STORE(jac + base + pos, (s * (o - xpv) - s * (o - xmv)) * ri);
    }
}

// path row q of node k in column X(k,j), j < 2 -- by central differences
template <int M>
ECUDA_HD void xcol_path_fd(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int j, int k, int q, double* jac) {
    const int N = ph.N, ns = pb.ns, nc = pb.nc, np = ph.npath;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
    const int lcol = nc * N + k * ns + j;
    const double xpv = m.xp[lcol], xmv = m.xm[lcol], ri = m.rinv[lcol];
    const double x0 = m.z[nc * N + k * ns], x1 = m.z[nc * N + k * ns + 1];
    const int r = ph.goff + ns * N + pb.ne + k * np + q;
    const double s = ECUDA_LDG(pb.sg + r);
    const PathRowAt<M> row(pb, ph, m, q, t);
    double vp = row(j == 0 ? xpv : x0, j == 1 ? xpv : x1);
    double vm = row(j == 0 ? xmv : x0, j == 1 ? xmv : x1);
    const int pos = N - 1 + pb.xcnt[j] + ((k == 0 || k == N - 1) ? 1 : 0) + q;
    ECUDA_STREAM_STORE(jac + m.colp[lcol] + pos, (s * vp - s * vm) * ri);
}

// defect rows of node k, event row, linkage row -- analytic
template <int M, bool SM = false>
ECUDA_HD void xcol_local_exact(const ProbDev& pb, const PhaseDev& ph, int p, const CtaMem& m, int j, int k,
                               double* jac) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, ns = pb.ns, nc = pb.nc, np = ph.npath;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double* sg = pb.sg;
    const int lcol = nc * N + k * ns + j, col = ph.zoff + lcol;
    const int base = m.colp[lcol];
    double x[NS], u[NCU];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = m.z[nc * N + k * ns + i];
#pragma unroll
    for (int i = 0; i < NCU; ++i) u[i] = m.z[k * nc + i];
    double dfdx[NS][NS], dfdu[NS][NCU];
    Model<M>::jac(x, u, pt.h * ECUDA_LDG(ph.tau + k) + pt.m, dfdx, dfdu);
    const double is = ECUDA_LDG(pb.isz + col);
    const double dkk = ECUDA_LDG(ph.Dt + static_cast<size_t>(k) * N + k);
    const int rdef0 = ph.goff + k * ns;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        int rk = pb.xrank[j][i];
        if (rk >= 0) {
            double d = 0.0;
#pragma unroll
            for (int jj = 0; jj < NS; ++jj)
                if (jj == j) d = dfdx[i][jj];
            double v = ((i == j) ? dkk : 0.0) - pt.h * d;
            jstore<SM>(jac + base + k + rk, (ECUDA_LDG(sg + rdef0 + i) * v) * is);
        }
    }
    int pos = N - 1 + pb.xcnt[j];
    if (k == 0 || k == N - 1) {
        int r = ph.goff + ns * N + (k == 0 ? j : ns + j);
        jstore<SM>(jac + base + pos, (ECUDA_LDG(sg + r) * 1.0) * is);
        ++pos;
    }
    if (j < 2) pos += np;
    if (k == N - 1 && p + 1 < pb.nphases) {
        int r = pb.linkoff + p * (ns + 1) + j;
        jstore<SM>(jac + base + pos, (ECUDA_LDG(sg + r) * 1.0) * is);
    }
    if (k == 0 && p > 0) {
        int r = pb.linkoff + (p - 1) * (ns + 1) + j;
        jstore<SM>(jac + base + pos, (ECUDA_LDG(sg + r) * -1.0) * is);
    }
}

// path row q of node k in column X(k,j), j < 2 -- analytic
template <int M, bool SM = false>
ECUDA_HD void xcol_path_exact(const ProbDev& pb, const PhaseDev& ph, const CtaMem& m, int j, int k, int q,
                              double* jac) {
    const int N = ph.N, ns = pb.ns, nc = pb.nc, np = ph.npath;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
    const int lcol = nc * N + k * ns + j, col = ph.zoff + lcol;
    const double x0 = m.z[nc * N + k * ns], x1 = m.z[nc * N + k * ns + 1];
    double ddx, ddy, ddt = 0.0;
    if (q < ph.nstat)
        Model<M>::static_row_dxy(m.inst + ph.inst_off + q * Model<M>::REC, x0, x1, &ddx, &ddy);
    else
        moving_row_partials<M>(pb, ph, m, q, x0, x1, t, &ddx, &ddy, &ddt);
    const double v = (j == 0) ? ddx : ddy;
    const int r = ph.goff + ns * N + pb.ne + k * np + q;
    const int pos = N - 1 + pb.xcnt[j] + ((k == 0 || k == N - 1) ? 1 : 0) + q;
    jstore<SM>(jac + m.colp[lcol] + pos, (ECUDA_LDG(pb.sg + r) * v) * ECUDA_LDG(pb.isz + col));
}

// whole node-local part of column X(k,j), written to the caller's global array (generic path)
template <int M>
ECUDA_HD void state_column_fd(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m,
                              int b, int j, int k, double dpk, double dmk) {
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    xcol_local_fd<M>(pb, ph, p, io, m, b, j, k, dpk, dmk, jac);
    if (j < 2)
        for (int q = 0; q < ph.npath; ++q) xcol_path_fd<M>(pb, ph, m, j, k, q, jac);
}
template <int M>
ECUDA_HD void state_column_exact(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m,
                                 int b, int j, int k) {
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    xcol_local_exact<M>(pb, ph, p, m, j, k, jac);
    if (j < 2)
        for (int q = 0; q < ph.npath; ++q) xcol_path_exact<M>(pb, ph, m, j, k, q, jac);
}

// Same as fd_block for any tail length and any number of valid nodes (runtime loops).
template <int NS>
ECUDA_HD double fd_block_generic(const double* __restrict__ Dtb, int N, const double* __restrict__ Xb,
                                 const double* __restrict__ XPb, const double* __restrict__ XMb,
                                 const double* __restrict__ RIb, const int* __restrict__ CPb, int nin, double pre,
                                 const double* __restrict__ Ptail, int ntail, int nthr, double sgr, double hfv,
                                 int krel, int cntm1, double* __restrict__ jac, double& dpk, double& dmk) {
    // Same operation sequence per triplet as the serial form; organised so that the block's D entries and
    // node values are loaded once, and every following block sum is read from shared memory once for all
    // eight nodes of the block (16 independent accumulators) instead of once per node.
    constexpr int BL = ECUDA_DOT_BLOCK;
    double d[BL], xv[BL], tp[BL], tm[BL];
#pragma unroll
    for (int i = 0; i < BL; ++i) {
        const bool in = i < nin;
        d[i] = in ? ECUDA_LDG(Dtb + static_cast<size_t>(i) * N) : 0.0;
        xv[i] = in ? Xb[i * NS] : 0.0;
    }
    double own = 0.0;
#pragma unroll
    for (int i = 0; i < BL; ++i)
        if (i < nin) own = fma(d[i], xv[i], own);
    double qa = 0.0;  // unperturbed prefix q[a]
#pragma unroll
    for (int a = 0; a < BL; ++a) {
        if (a < nin) {
            double sp = fma(d[a], XPb[a * NS], qa);
            double sm = fma(d[a], XMb[a * NS], qa);
#pragma unroll
            for (int i = a + 1; i < BL; ++i)
                if (i < nin) {
                    sp = fma(d[i], xv[i], sp);
                    sm = fma(d[i], xv[i], sm);
                }
            tp[a] = pre + sp;
            tm[a] = pre + sm;
            qa = fma(d[a], xv[a], qa);
        } else {
            tp[a] = tm[a] = 0.0;
        }
    }
    for (int t = 0; t < ntail; ++t) {
        const double pv = Ptail[t * nthr];
#pragma unroll
        for (int a = 0; a < BL; ++a) {
            tp[a] = tp[a] + pv;
            tm[a] = tm[a] + pv;
        }
    }
#pragma unroll
    for (int a = 0; a < BL; ++a) {
        if (a >= nin) continue;
        if (a != krel) {
            const double gp = sgr * (tp[a] - hfv);
            const double gm = sgr * (tm[a] - hfv);
            ECUDA_STREAM_STORE(jac + CPb[a * NS] + (krel > a ? cntm1 : 0), (gp - gm) * RIb[a * NS]);
        } else {
            dpk = tp[a];
            dmk = tm[a];
        }
    }
    return own;
}

// One summation block of the row-restricted FD of defect row (k,j): for each of the block's nodes
// ls = l0+a the two perturbed dots (X[ls][j] -> +-), rebuilt from the unperturbed in-block prefix
// q[a], the serial prefix `pre` over the earlier blocks and the NT block sums that follow.
//   NS   state stride (compile time), NT  number of following blocks, FULL  block has BL valid nodes
template <int NS, int NT, bool FULL>
ECUDA_HD double fd_block(const double* __restrict__ Dtb, int N, const double* __restrict__ Xb,
                         const double* __restrict__ XPb, const double* __restrict__ XMb,
                         const double* __restrict__ RIb, const int* __restrict__ CPb, int nin, double pre,
                         const double* __restrict__ Ptail, int ntail, int nthr, double sgr, double hfv, int krel,
                         int cntm1, double* __restrict__ jac, double& dpk, double& dmk) {
    constexpr int BL = ECUDA_DOT_BLOCK;
    double d[BL], xv[BL], q[BL + 1], T[NT > 0 ? NT : 1];
    q[0] = 0.0;
#pragma unroll
    for (int i = 0; i < BL; ++i) {
        if (FULL || i < nin) {
            d[i] = ECUDA_LDG(Dtb + static_cast<size_t>(i) * N);
            xv[i] = Xb[i * NS];
            q[i + 1] = fma(d[i], xv[i], q[i]);  // unperturbed in-block prefix
        } else {
            d[i] = 0.0;
            xv[i] = 0.0;
            q[i + 1] = q[i];
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) T[t] = (t < ntail) ? Ptail[t * nthr] : 0.0;
#pragma unroll
    for (int a = 0; a < BL; ++a) {
        if (FULL || a < nin) {
            double sp = fma(d[a], XPb[a * NS], q[a]);
            double sm = fma(d[a], XMb[a * NS], q[a]);
            if (FULL) {
#pragma unroll
                for (int i = a + 1; i < BL; ++i) {
                    sp = fma(d[i], xv[i], sp);
                    sm = fma(d[i], xv[i], sm);
                }
            } else {
                for (int i = a + 1; i < nin; ++i) {
                    double di = ECUDA_LDG(Dtb + static_cast<size_t>(i) * N), xi = Xb[i * NS];
                    sp = fma(di, xi, sp);
                    sm = fma(di, xi, sm);
                }
            }
            double tp = pre + sp;  // a block sum is never -0.0, so 0.0 + s == s bit for bit when pre == 0
            double tm = pre + sm;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                tp = tp + T[t];
                tm = tm + T[t];
            }
            if (a != krel) {
                double gp = sgr * (tp - hfv);
                double gm = sgr * (tm - hfv);
                // rows k < ls sit at position k of column X(ls,j); rows k > ls come after its node-local block
                ECUDA_STREAM_STORE(jac + CPb[a * NS] + (krel > a ? cntm1 : 0), (gp - gm) * RIb[a * NS]);
            } else {
                dpk = tp;
                dmk = tm;
            }
        }
    }
    return q[BL];  // this block's unperturbed sum
}

// State item (j,k). NB > 0: the phase has exactly NB summation blocks, so the block sums that
// follow the current block fit a register window of NB-1 values; NB == 0: any block count.
template <int M, int NB>
ECUDA_HD void state_item(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int j,
                         int k, int tid, int nthr, bool recompute) {
    constexpr int BL = ECUDA_DOT_BLOCK;
    const int N = ph.N, ns = pb.ns, nc = pb.nc;
    const int r = ph.goff + k * ns + j;
    const double sgr = ECUDA_LDG(pb.sg + r);
    const double hfv = m.hf[k * ns + j];
    if (io.g) ECUDA_STREAM_STORE(io.g + static_cast<size_t>(b) * pb.ncons + r, sgr * (m.dotv[k * ns + j] - hfv));
    if (!io.jac) return;
    double* jac = io.jac + static_cast<size_t>(b) * pb.nnz;
    const int xoff = nc * N + j;  // phase-local index of X(0,j); X(l,j) = xoff + l*ns
    const double* Dt = ph.Dt + k;
    if (io.jac_mode == ECUDA_JAC_EXACT) {
        const int cntm1 = pb.xcnt[j] - 1;
        // batches of 8 columns: all loads of a batch are in flight before the first store (a serial loop pays
        // one L2 round trip per triplet, which was 36 % of the stall samples of a 200-node instance)
        for (int l0 = 0; l0 < N; l0 += 8) {
            double dv[8], sv[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int l = l0 + a;
                if (l < N && l != k) {
                    dv[a] = ECUDA_LDG(Dt + static_cast<size_t>(l) * N);
                    sv[a] = ECUDA_LDG(pb.isz + ph.zoff + xoff + l * ns);
                }
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int l = l0 + a;
                if (l < N && l != k)
                    ECUDA_STREAM_STORE(jac + m.colp[xoff + l * ns] + (k < l ? k : k + cntm1), (sgr * dv[a]) * sv[a]);
            }
        }
        state_column_exact<M>(pb, ph, p, io, m, b, j, k);
        return;
    }
    // ---- index-set central differences, row-restricted ----------------------------------------------
    const double* X = m.z + xoff;
    const double* XP = m.xp + xoff;
    const double* XM = m.xm + xoff;
    const double* RI = m.rinv + xoff;
    const int* CP = m.colp + xoff;
    const int cntm1 = pb.xcnt[j] - 1;
    double dpk = 0.0, dmk = 0.0;  // (D X)[k][j] with X[k][j] perturbed (column X(k,j) itself)
    // block sums P[bi] of the unperturbed row in thread-private shared memory (P[bi*nthr]). When
    // every thread owns at most one state item, phase B left them there; otherwise rebuild them.
    double* P = m.P + tid;
    if (recompute) dot_row(pb, ph, m, k, j, P, nthr);
    constexpr int NSC = Model<M>::NS;  // == pb.ns for every model
    const int nb = NB > 0 ? NB : ph.nb;
    double* jk = jac + k;  // entry of row k in column X(ls,j): colptr + k (+ cntm1 when k > ls)
    double pre = 0.0;      // P[0] + ... + P[bi-1], serial
#pragma unroll 1
    for (int bi = 0; bi < nb; ++bi) {
        const int l0 = bi * BL;
        const int nin = (N - l0) < BL ? (N - l0) : BL;
        const int ntail = nb - 1 - bi;
        const double* Dtb = Dt + static_cast<size_t>(l0) * N;
        const double* Ptail = P + (bi + 1) * nthr;
        const int o = l0 * NSC;
        const int krel = k - l0;  // position of the diagonal inside this block, or outside [0,BL)
        double own;
        constexpr int NTW = NB > 0 ? NB - 1 : 0;
        if (NB > 0 && nin == BL) {
            // one code path for every block: the register window always holds NB-1 following sums,
            // padded with +0.0 (x + 0.0 == x bit for bit: no partial sum here is ever -0.0)
            own = fd_block<NSC, NTW, true>(Dtb, N, X + o, XP + o, XM + o, RI + o, CP + o, nin, pre, Ptail, ntail, nthr,
                                           sgr, hfv, krel, cntm1, jk, dpk, dmk);
        } else {
            // any block count / partial last block: runtime loops, tail sums read from shared memory
            own = fd_block_generic<NSC>(Dtb, N, X + o, XP + o, XM + o, RI + o, CP + o, nin, pre, Ptail, ntail, nthr,
                                        sgr, hfv, krel, cntm1, jk, dpk, dmk);
        }
        pre = pre + own;
    }
    state_column_fd<M>(pb, ph, p, io, m, b, j, k, dpk, dmk);
}

// node-local Jacobian entries of column c at node k: c in [0,nc) control, nc -> t0, nc+1 -> tf
template <int M, bool SM = false>
ECUDA_HD void node_item(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, const CtaMem& m, int b,
                        int k, int c, double* jacbase = nullptr) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, ns = pb.ns, nc = pb.nc, np = ph.npath;
    const bool fd = io.jac_mode == ECUDA_JAC_FD_INDEXSET;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double tau = ECUDA_LDG(ph.tau + k);
    const double t = pt.h * tau + pt.m;
    const double* sg = pb.sg;
    double* jac = jacbase ? jacbase : io.jac + static_cast<size_t>(b) * pb.nnz;
    double x[NS], u[NCU];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = m.z[nc * N + k * ns + i];
#pragma unroll
    for (int i = 0; i < NCU; ++i) u[i] = m.z[k * nc + i];
    const int rdef0 = ph.goff + k * ns;
    const int rpath0 = ph.goff + ns * N + pb.ne + k * np;

    if (c < nc) {  // ---- control column U(k, j)
        const int j = c, lcol = k * nc + j, col = ph.zoff + lcol;
        const int base = m.colp[lcol];
        if (fd) {
            double up[NCU], um[NCU], fp[NS], fm[NS];
#pragma unroll
            for (int i = 0; i < NCU; ++i) {
                up[i] = (i == j) ? m.xp[lcol] : u[i];
                um[i] = (i == j) ? m.xm[lcol] : u[i];
            }
            Model<M>::f(x, up, t, fp);
            Model<M>::f(x, um, t, fm);
            const double ri = m.rinv[lcol];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                int rk = pb.urank[j][i];
                if (rk >= 0) {
                    double s = ECUDA_LDG(sg + rdef0 + i), dv = m.dotv[k * ns + i];
                    double gp = s * (dv - pt.h * fp[i]);
                    double gm = s * (dv - pt.h * fm[i]);
                    jstore<SM>(jac + base + rk, (gp - gm) * ri);
                }
            }
        } else {
            double dfdx[NS][NS], dfdu[NS][NCU];
            Model<M>::jac(x, u, t, dfdx, dfdu);
            const double is = ECUDA_LDG(pb.isz + col);
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                int rk = pb.urank[j][i];
                if (rk >= 0) {
                    double d = 0.0;
#pragma unroll
                    for (int jj = 0; jj < NCU; ++jj)
                        if (jj == j) d = dfdu[i][jj];
                    double v = -(pt.h * d);
                    jstore<SM>(jac + base + rk, (ECUDA_LDG(sg + rdef0 + i) * v) * is);
                }
            }
        }
        return;
    }
    // ---- time columns t0 / tf
    const int which = c - nc;  // 0: t0, 1: tf
    const int lcol = (ns + nc) * N + which, col = ph.zoff + lcol;
    const int base = m.colp[lcol];
    const int ntr = np - ph.nstat;
    if (fd) {
        const double ri = m.rinv[lcol];
        const double t0p = which == 0 ? m.xp[lcol] : pt.t0, tfp = which == 1 ? m.xp[lcol] : pt.tf;
        const double t0m = which == 0 ? m.xm[lcol] : pt.t0, tfm = which == 1 ? m.xm[lcol] : pt.tf;
        const double hp = 0.5 * (tfp - t0p), mp = 0.5 * (tfp + t0p);
        const double hm = 0.5 * (tfm - t0m), mm = 0.5 * (tfm + t0m);
        const double tp = hp * tau + mp, tm = hm * tau + mm;
        double fp[NS], fm[NS];
        Model<M>::f(x, u, tp, fp);
        Model<M>::f(x, u, tm, fm);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double s = ECUDA_LDG(sg + rdef0 + i), dv = m.dotv[k * ns + i];
            double gp = s * (dv - hp * fp[i]);
            double gm = s * (dv - hm * fm[i]);
            jstore<SM>(jac + base + k * ns + i, (gp - gm) * ri);
        }
        for (int q = ph.nstat; q < np; ++q) {
            double s = ECUDA_LDG(sg + rpath0 + q);
            double vp = path_row<M>(pb, ph, m, q, x[0], x[1], tp);
            double vm = path_row<M>(pb, ph, m, q, x[0], x[1], tm);
            jstore<SM>(jac + base + ns * N + k * ntr + (q - ph.nstat), (s * vp - s * vm) * ri);
        }
        if (k == 0) {
            int r = ph.goff + ns * N + pb.ne + np * N;
            double s = ECUDA_LDG(sg + r);
            jstore<SM>(jac + base + ns * N + N * ntr, (s * (tfp - t0p) - s * (tfm - t0m)) * ri);
            if (which == 0 && p > 0) {
                const PhaseDev& pv = pb.ph[p - 1];
                int rl = pb.linkoff + (p - 1) * (ns + 1) + ns;
                double sl = ECUDA_LDG(sg + rl);
                double o = other_phase_value(pb, io, b, pv.zoff + (ns + nc) * pv.N + 1);
                jstore<SM>(jac + base + ns * N + N * ntr + 1, (sl * (o - t0p) - sl * (o - t0m)) * ri);
            }
            if (which == 1 && p + 1 < pb.nphases) {
                const PhaseDev& nx = pb.ph[p + 1];
                int rl = pb.linkoff + p * (ns + 1) + ns;
                double sl = ECUDA_LDG(sg + rl);
                double o = other_phase_value(pb, io, b, nx.zoff + (ns + nc) * nx.N);
                jstore<SM>(jac + base + ns * N + N * ntr + 1, (sl * (tfp - o) - sl * (tfm - o)) * ri);
            }
        }
    } else {
        const double is = ECUDA_LDG(pb.isz + col);
        const double dtk = which == 0 ? 0.5 * (1.0 - tau) : 0.5 * (1.0 + tau);  // d t_k / d t0|tf
        double f[NS], ft[NS], Lt;
        Model<M>::f(x, u, t, f);
        if constexpr (Model<M>::TDEP) Model<M>::dtime(x, u, t, ft, &Lt);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            // d zeta / d t0 = +f/2 ; d zeta / d tf = -f/2 ; time-dependent dynamics add -h (df/dt) (d t_k / d t0|tf)
            double v = which == 0 ? 0.5 * f[i] : -0.5 * f[i];
            if constexpr (Model<M>::TDEP) v = v - pt.h * (ft[i] * dtk);
            jstore<SM>(jac + base + k * ns + i, (ECUDA_LDG(sg + rdef0 + i) * v) * is);
        }
        for (int q = ph.nstat; q < np; ++q) {
            double ddx, ddy, ddt;
            moving_row_partials<M>(pb, ph, m, q, x[0], x[1], t, &ddx, &ddy, &ddt);
            jstore<SM>(jac + base + ns * N + k * ntr + (q - ph.nstat),
                               (ECUDA_LDG(sg + rpath0 + q) * (ddt * dtk)) * is);
        }
        if (k == 0) {
            int r = ph.goff + ns * N + pb.ne + np * N;
            jstore<SM>(jac + base + ns * N + N * ntr, (ECUDA_LDG(sg + r) * (which == 0 ? -1.0 : 1.0)) * is);
            if (which == 0 && p > 0) {
                int rl = pb.linkoff + (p - 1) * (ns + 1) + ns;
                jstore<SM>(jac + base + ns * N + N * ntr + 1, (ECUDA_LDG(sg + rl) * -1.0) * is);
            }
            if (which == 1 && p + 1 < pb.nphases) {
                int rl = pb.linkoff + p * (ns + 1) + ns;
                jstore<SM>(jac + base + ns * N + N * ntr + 1, (ECUDA_LDG(sg + rl) * 1.0) * is);
            }
        }
    }
}

template <int M, int NB>
ECUDA_HD void phase_c(const ProbDev& pb, const PhaseDev& ph, int p, const EvalIO& io, CtaMem& m, int b, int tid,
                      int nthr, int slice = 0, int nslices = 1) {
    const int N = ph.N, ns = pb.ns, nc = pb.nc;
    const int nown = slice_nodes(N, slice, nslices);
    if (slice == 0 && tid == nthr - 1) objective_phase(pb, ph, p, io, m, b);
    if (io.g || io.jac) {
        // when a thread owns more than one defect row, phase B left only the last row's block sums
        const bool recompute = ns * nown > nthr;
        for (int it = tid; it < ns * nown; it += nthr) {
            int j = it / nown, k = slice + (it - j * nown) * nslices;
            state_item<M, NB>(pb, ph, p, io, m, b, j, k, tid, nthr, recompute);
        }
    }
    if (io.jac) {
        // the lighter node items are handed out from the top of the thread range downwards, so the
        // threads that had no state item above start on these first
        const int nitems = (nc + 2) * nown;
        for (int it = nthr - 1 - tid; it < nitems; it += nthr) {
            int c = it / nown, k = slice + (it - c * nown) * nslices;
            node_item<M>(pb, ph, p, io, m, b, k, c);
        }
    }
}

// running cost per node only (the gradient kernel needs Lk but none of the rest of phase B)
template <int M>
ECUDA_HD void cost_nodes(const ProbDev& pb, const PhaseDev& ph, CtaMem& m, int tid, int nthr) {
    const int N = ph.N, ns = pb.ns, nc = pb.nc;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    for (int k = tid; k < N; k += nthr) {
        double t = pt.h * ECUDA_LDG(ph.tau + k) + pt.m;
        double L = Model<M>::cost(m.z + nc * N + k * ns, m.z + k * nc, t);
        m.Lk[k] = pb.maximize ? -1.0 * L : L;
    }
}

// gradient of the objective (exact), one thread per node
template <int M>
ECUDA_HD void gradient_phase(const ProbDev& pb, const PhaseDev& ph, const EvalIO& io, const CtaMem& m, int b,
                             int tid, int nthr) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, ns = pb.ns, nc = pb.nc;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double* grad = io.grad + static_cast<size_t>(b) * pb.nvars + ph.zoff;
    const double* is = pb.isz + ph.zoff;
    for (int k = tid; k < N; k += nthr) {
        const double* x = m.z + nc * N + k * ns;
        const double* u = m.z + k * nc;
        double dx[NS], du[NCU];
        Model<M>::dcost(x, u, pt.h * ECUDA_LDG(ph.tau + k) + pt.m, dx, du);
        const double w = ECUDA_LDG(ph.w + k);
        const double sgn = pb.maximize ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            int c = nc * N + k * ns + i;
            grad[c] = (pb.sf * (pt.h * (w * (sgn * dx[i])))) * ECUDA_LDG(is + c);
        }
        for (int j = 0; j < nc; ++j) {
            int c = k * nc + j;
            double d = 0.0;
#pragma unroll
            for (int jj = 0; jj < NCU; ++jj)
                if (jj == j) d = du[jj];
            grad[c] = (pb.sf * (pt.h * (w * (sgn * d)))) * ECUDA_LDG(is + c);
        }
    }
    if (tid == 0) {
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc = fma(ECUDA_LDG(ph.w + k), m.Lk[k], acc);
        int c0 = (ns + nc) * N;
        double g0 = -0.5 * acc, g1 = 0.5 * acc;
        if constexpr (Model<M>::TDEP) {
            // a running cost that reads t: + h sum_k w_k (dL/dt)(t_k) (d t_k / d t0|tf)
            double a0 = 0.0, a1 = 0.0;
            const double sgn = pb.maximize ? -1.0 : 1.0;
            for (int k = 0; k < N; ++k) {
                const double tau = ECUDA_LDG(ph.tau + k);
                double ft[NS], Lt;
                Model<M>::dtime(m.z + nc * N + k * ns, m.z + k * nc, pt.h * tau + pt.m, ft, &Lt);
                const double wl = ECUDA_LDG(ph.w + k) * (sgn * Lt);
                a0 = fma(wl, 0.5 * (1.0 - tau), a0);
                a1 = fma(wl, 0.5 * (1.0 + tau), a1);
            }
            g0 = g0 + pt.h * a0;
            g1 = g1 + pt.h * a1;
        }
        grad[c0] = (pb.sf * g0) * ECUDA_LDG(is + c0);
        grad[c0 + 1] = (pb.sf * g1) * ECUDA_LDG(is + c0 + 1);
    }
}

// ---- discretisation error per mesh interval (ecuda_ode_error; the estimate is stated in ecuda_mesh.cpp) ----
// step 1, all threads: (D X)[k][i] into m.dotv. Barrier. step 2, threads i < ns: the scale weights
// w_i = max_k max(|x_ik|, |(D X)_ki / h|) into m.hf[i]. Barrier. step 3, thread k < N-1: Gauss-Legendre
// quadrature of |dx~/dtau - h f(x~, u~)| over interval k, states and controls interpolated with the
// Lagrange basis rows E / dE (serial ascending fma chains).
template <int M>
ECUDA_HD void ode_error_dots(const ProbDev& pb, const PhaseDev& ph, CtaMem& m, int tid, int nthr) {
    const int ns = pb.ns;
    for (int it = tid; it < ph.N * ns; it += nthr) {
        const int k = it / ns, j = it - k * ns;
        m.dotv[k * ns + j] = dot_row(pb, ph, m, k, j, m.P + tid, nthr);
    }
}
template <int M>
ECUDA_HD void ode_error_weights(const ProbDev& pb, const PhaseDev& ph, CtaMem& m, int tid) {
    const int ns = pb.ns, N = ph.N;
    if (tid >= ns) return;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    double w = 0.0;
    for (int k = 0; k < N; ++k) {
        w = fmax(w, fabs(m.z[pb.nc * N + k * ns + tid]));
        w = fmax(w, fabs(m.dotv[k * ns + tid] / pt.h));
    }
    m.hf[tid] = w;
}
// one thread per quadrature point: |dx~/dtau - h f(x~, u~)| of every state into m.P[(k*Q + q)*NS + i]
template <int M>
ECUDA_HD void ode_error_points(const ProbDev& pb, const PhaseDev& ph, int p, const MeshDev& mesh, CtaMem& m, int tid,
                               int nthr) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU;
    const int N = ph.N, nc = pb.nc;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    for (int r = tid; r < (N - 1) * ECUDA_MESH_Q; r += nthr) {
        const double* E = mesh.E[p] + static_cast<size_t>(r) * N;
        const double* dE = mesh.dE[p] + static_cast<size_t>(r) * N;
        double xq[NS], dxq[NS], uq[NCU], f[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) xq[i] = dxq[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NCU; ++j) uq[j] = 0.0;
        for (int l = 0; l < N; ++l) {
            const double e = ECUDA_LDG(E + l), de = ECUDA_LDG(dE + l);
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const double xv = m.z[nc * N + l * NS + i];
                xq[i] = fma(e, xv, xq[i]);
                dxq[i] = fma(de, xv, dxq[i]);
            }
#pragma unroll
            for (int j = 0; j < NCU; ++j) uq[j] = fma(e, m.z[l * nc + j], uq[j]);
        }
        Model<M>::f(xq, uq, pt.h * ECUDA_LDG(mesh.tq[p] + r) + pt.m, f);
#pragma unroll
        for (int i = 0; i < NS; ++i) m.P[static_cast<size_t>(r) * NS + i] = fabs(dxq[i] - pt.h * f[i]);
    }
}
// after a barrier, thread k: the quadrature sums of interval k in point order, then the scaled maximum
template <int M>
ECUDA_HD void ode_error_intervals(const ProbDev& pb, const PhaseDev& ph, int p, const MeshDev& mesh, const CtaMem& m,
                                  int b, int tid, int nthr) {
    constexpr int NS = Model<M>::NS;
    const int N = ph.N;
    for (int k = tid; k < N - 1; k += nthr) {
        double err = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double eta = 0.0;
            for (int q = 0; q < ECUDA_MESH_Q; ++q) {
                const size_t r = static_cast<size_t>(k) * ECUDA_MESH_Q + q;
                eta = fma(ECUDA_LDG(mesh.wq[p] + r), m.P[r * NS + i], eta);
            }
            err = fmax(err, eta / (m.hf[i] + 1.0));
        }
        mesh.out[static_cast<size_t>(b) * mesh.nint + mesh.eoff[p] + k] = err;
    }
}

// ---- Hessian of the Lagrangian (ecuda_eval_hess) ---------------------------------------------------------
// Node k contributes, with h = (tf - t0)/2, t = h tau_k + m, a = dt/dt0 = (1 - tau_k)/2, b = dt/dtf = (1 + tau_k)/2:
//   c_i = -lambda~_r sg_r for its defect rows, c_L = sigma sf w_k (negated when maximising), c_q = lambda~_r sg_r
//   for its path rows;  node block  h (sum_i c_i d2 f_i + c_L d2 L) + sum_q c_q d2 p_q;
//   (v, t0 | tf) = (-+1/2) (sum_i c_i df_i/dv + c_L dL/dv) + c_q d2p_q/dvdt (a | b)   [moving circles only];
//   (t0 | tf, t0 | tf) = sum over nodes and moving circles of c_q d2p_q/dt2 (a a | a b | b b).
// Built-in dynamics and cost are autonomous and add nothing to the time-time block; user models that read t add, with
// G = sum_i c_i f_i + c_L L:  (v, t0 | tf) += h G_vt (a | b);  (t0,t0) += -G_t a + h G_tt a a;  (t0,tf) += G_t (a - b)/2 +
// h G_tt a b;  (tf,tf) += G_t b + h G_tt b b  (Model::tdir). Traced path rows bring all six second derivatives in
// (x_0, x_1, t) (Model::user_row_hess). Entries are multiplied by
// 1/sz of both variables (the solver's variables are z sz). Thread k owns node k; the 3 time-time partial
// sums per node go through shared memory (m.hf, 3 per node) and thread 0 adds them in node order.
template <int M>
ECUDA_HD void hess_nodes(const ProbDev& pb, const PhaseDev& ph, int p, const HessIO& hio, CtaMem& m, int b, int tid,
                         int nthr) {
    constexpr int NS = Model<M>::NS, NCU = Model<M>::NCU, NV = NS + NCU;
    const int N = ph.N, nc = pb.nc, np = ph.npath, nstat = ph.nstat;
    const PhaseTimes pt = phase_times(pb, ph, m.z);
    const double* lam = hio.lambda + static_cast<size_t>(b) * pb.ncons;
    const double* sg = pb.sg;
    const double* isz = pb.isz + ph.zoff;
    const double sig = (hio.sigma ? ECUDA_LDG(hio.sigma + b) : hio.sigma0) * pb.sf * (pb.maximize ? -1.0 : 1.0);
    double* out = hio.vals + static_cast<size_t>(b) * hio.nnz_h + hio.hoff[p];
    const int Cu = hess_cu(NS, nc), Cx = hess_cx(NS);
    const int it0 = (NS + nc) * N;
    const double iszt0 = ECUDA_LDG(isz + it0), iszt1 = ECUDA_LDG(isz + it0 + 1);
    for (int k = tid; k < N; k += nthr) {
        const double tau = ECUDA_LDG(ph.tau + k);
        const double t = pt.h * tau + pt.m, ta = 0.5 * (1.0 - tau), tb = 0.5 * (1.0 + tau);
        double x[NS], u[NCU], lamf[NS], cdef[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            x[i] = m.z[nc * N + k * NS + i];
            const int r = ph.goff + k * NS + i;
            cdef[i] = -(ECUDA_LDG(lam + r) * ECUDA_LDG(sg + r));
            lamf[i] = cdef[i] * pt.h;
        }
#pragma unroll
        for (int j = 0; j < NCU; ++j) u[j] = m.z[k * nc + j];
        const double cL = sig * ECUDA_LDG(ph.w + k);
        double H[NV][NV];
        Model<M>::hess(x, u, t, lamf, cL * pt.h, H);
        // first derivatives for the couplings with t0 / tf
        double dfdx[NS][NS], dfdu[NS][NCU], dLx[NS], dLu[NCU], gv[NV];
        Model<M>::jac(x, u, t, dfdx, dfdu);
        Model<M>::dcost(x, u, t, dLx, dLu);
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            double s = cL * dLx[a];
#pragma unroll
            for (int i = 0; i < NS; ++i) s = fma(cdef[i], dfdx[i][a], s);
            gv[a] = s;
        }
#pragma unroll
        for (int a = 0; a < NCU; ++a) {
            double s = cL * dLu[a];
#pragma unroll
            for (int i = 0; i < NS; ++i) s = fma(cdef[i], dfdu[i][a], s);
            gv[NS + a] = s;
        }
        // dynamics / cost that read t (user models): d2/dv dt, d/dt and d2/dt2 of h (sum_i c_i f_i + c_L L)
        double gt[NV], tt0 = 0.0, tt1 = 0.0, tt2 = 0.0;
#pragma unroll
        for (int a = 0; a < NV; ++a) gt[a] = 0.0;
        if constexpr (Model<M>::TDEP) {
            double sth, stth, st1, stt1, g1[NV];
            Model<M>::tdir(x, u, t, lamf, cL * pt.h, gt, &sth, &stth);   // weights carry h: h G_vt, h G_t, h G_tt
            Model<M>::tdir(x, u, t, cdef, cL, g1, &st1, &stt1);          // G_t without h (from d h / d t0|tf = -+ 1/2)
            (void)sth; (void)stt1; (void)g1;
            tt0 = -(st1 * ta) + (stth * ta) * ta;
            tt1 = 0.5 * (st1 * (ta - tb)) + (stth * ta) * tb;
            tt2 = st1 * tb + (stth * tb) * tb;
        }
        // path rows: static obstacles (position block), moving circles (position block, time couplings)
        double xt[2] = {0.0, 0.0}, tt = 0.0;
        const int rp0 = ph.goff + NS * N + pb.ne + k * np;
        for (int q = 0; q < np; ++q) {
            const double c = ECUDA_LDG(lam + rp0 + q) * ECUDA_LDG(sg + rp0 + q);
            if (q < nstat) {
                double hxx, hxy, hyy;
                Model<M>::static_row_hess(m.inst + ph.inst_off + q * Model<M>::REC, &hxx, &hxy, &hyy);
                H[0][0] = fma(c, hxx, H[0][0]);
                H[0][1] = fma(c, hxy, H[0][1]);
                H[1][0] = fma(c, hxy, H[1][0]);
                H[1][1] = fma(c, hyy, H[1][1]);
            } else if (Model<M>::NUSER > 0 && q - nstat >= pb.ntracks) {  // traced path row of a user model
                double h6[6];
                if constexpr (Model<M>::NUSER > 0) Model<M>::user_row_hess(q - nstat - pb.ntracks, x[0], x[1], t, h6);
                H[0][0] = fma(c, h6[0], H[0][0]);
                H[0][1] = fma(c, h6[1], H[0][1]);
                H[1][0] = fma(c, h6[1], H[1][0]);
                H[1][1] = fma(c, h6[2], H[1][1]);
                xt[0] = fma(c, h6[3], xt[0]);
                xt[1] = fma(c, h6[4], xt[1]);
                tt = fma(c, h6[5], tt);
            } else {
                double hxt, hyt, htt;
                track_row_hess(m.inst + pb.track_off + (q - nstat) * pb.track_size, pb.nway, t, &hxt, &hyt, &htt);
                H[0][0] = fma(c, -2.0, H[0][0]);
                H[1][1] = fma(c, -2.0, H[1][1]);
                xt[0] = fma(c, hxt, xt[0]);
                xt[1] = fma(c, hyt, xt[1]);
                tt = fma(c, htt, tt);
            }
        }
        m.hf[3 * k] = (tt * ta) * ta + tt0;
        m.hf[3 * k + 1] = (tt * ta) * tb + tt1;
        m.hf[3 * k + 2] = (tt * tb) * tb + tt2;
        // control columns of the node
        for (int j = 0; j < nc; ++j) {
            double* col = out + k * Cu + j * (nc + NS + 2) - j * (j - 1) / 2;
            const double sj = ECUDA_LDG(isz + k * nc + j);
            for (int j2 = j; j2 < nc; ++j2) {
                double v = 0.0;
#pragma unroll
                for (int a = 0; a < NCU; ++a)
#pragma unroll
                    for (int c2 = 0; c2 < NCU; ++c2)
                        if (a == j && c2 == j2) v = H[NS + a][NS + c2];
                col[j2 - j] = (v * sj) * ECUDA_LDG(isz + k * nc + j2);
            }
            double gj = 0.0;
#pragma unroll
            for (int a = 0; a < NCU; ++a)
                if (a == j) gj = gv[NS + a];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                double v = 0.0;
#pragma unroll
                for (int a = 0; a < NCU; ++a)
                    if (a == j) v = H[i][NS + a];
                col[(nc - j) + i] = (v * sj) * ECUDA_LDG(isz + nc * N + k * NS + i);
            }
            double gtj = 0.0;
#pragma unroll
            for (int a = 0; a < NCU; ++a)
                if (a == j) gtj = gt[NS + a];
            col[(nc - j) + NS] = ((-0.5 * gj + gtj * ta) * sj) * iszt0;
            col[(nc - j) + NS + 1] = ((0.5 * gj + gtj * tb) * sj) * iszt1;
        }
        // state columns of the node
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double* col = out + N * Cu + k * Cx + i * (NS + 2) - i * (i - 1) / 2;
            const double si = ECUDA_LDG(isz + nc * N + k * NS + i);
#pragma unroll
            for (int i2 = i; i2 < NS; ++i2) col[i2 - i] = (H[i][i2] * si) * ECUDA_LDG(isz + nc * N + k * NS + i2);
            const double xti = (i < 2 ? xt[i] : 0.0) + gt[i];
            col[NS - i] = ((-0.5 * gv[i] + xti * ta) * si) * iszt0;
            col[NS - i + 1] = ((0.5 * gv[i] + xti * tb) * si) * iszt1;
        }
    }
}
// after a barrier: the time-time block
ECUDA_HD void hess_time_block(const ProbDev& pb, const PhaseDev& ph, int p, const HessIO& hio, const CtaMem& m, int b,
                              int tid) {
    if (tid != 0) return;
    const int N = ph.N, it0 = (pb.ns + pb.nc) * N;
    const double* isz = pb.isz + ph.zoff;
    double s00 = 0.0, s01 = 0.0, s11 = 0.0;
    for (int k = 0; k < N; ++k) {
        s00 = s00 + m.hf[3 * k];
        s01 = s01 + m.hf[3 * k + 1];
        s11 = s11 + m.hf[3 * k + 2];
    }
    double* out = hio.vals + static_cast<size_t>(b) * hio.nnz_h + hio.hoff[p] + N * (hess_cu(pb.ns, pb.nc) + hess_cx(pb.ns));
    const double z0 = ECUDA_LDG(isz + it0), z1 = ECUDA_LDG(isz + it0 + 1);
    out[0] = (s00 * z0) * z0;
    out[1] = (s01 * z0) * z1;
    out[2] = (s11 * z1) * z1;
}

}  // namespace ecuda
#endif
