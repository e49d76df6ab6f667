// ecuda_host.cpp -- host half of libecuda.so: NLP dimensions, Gauss-Lobatto collocation data,
// Jacobian sparsity pattern in CSC form, Curtis-Powell-Reid column groups, obstacle-edge geometry.
// Runs once per problem structure; nothing here needs a GPU.
//
// What it stands in for on the reference side:
//   dimensions   ePSOPT::setup            src/ePSOPT/ePSOPT.cpp:41-45,58
//   collocation  PSOPT "Legendre"/"Chebyshev" selected at src/ePSOPT/ePSOPT.cpp:68
//   pattern      the (iRow,jCol) structure PSOPT/ADOL-C hand to IPOPT (mode at ePSOPT.cpp:64)
//   edge records static geometry of src/Examples/PSOPT/etol_psopt_example1.cpp:164-172,178-179
// Compiled with -ffp-contract=off so that every expression below is a sequence of single IEEE
// operations.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ecuda_internal.hpp"

namespace ecuda {

bool model_info(int model, ModelInfo* m) {
    std::memset(m, 0, sizeof(*m));
    m->path_x = 0x3u;  // every obstacle row reads the two horizontal position states
    switch (model) {
        case ECUDA_MODEL_SI2D:
            m->ns = 2; m->nc_default = 2; m->nc_used = 2; m->rec_size = 6;
            m->fu[0] = 0x1u; m->fu[1] = 0x2u;
            return true;
        case ECUDA_MODEL_PM3D:
            m->ns = 6; m->nc_default = 3; m->nc_used = 3; m->rec_size = 4;
            m->fx[0] = 1u << 3; m->fx[1] = 1u << 4; m->fx[2] = 1u << 5;
            m->fu[3] = 0x1u; m->fu[4] = 0x2u; m->fu[5] = 0x4u;
            return true;
        case ECUDA_MODEL_FW6:
            m->ns = 6; m->nc_default = 3; m->nc_used = 3; m->rec_size = 4;
            m->fx[0] = m->fx[1] = (1u << 3) | (1u << 4) | (1u << 5);
            m->fx[2] = (1u << 3) | (1u << 4);
            m->fx[3] = (1u << 4);
            m->fu[3] = 0x1u; m->fu[4] = 0x2u; m->fu[5] = 0x4u;
            return true;
        default: {
            const UserModel* um = user_model(model);
            if (!um) return false;
            m->ns = um->ns;
            m->nc_default = m->nc_used = um->nc;
            m->rec_size = um->static_kind == ECUDA_STATIC_EDGE ? 6 : 4;
            for (int i = 0; i < um->ns; ++i) {
                m->fx[i] = um->fx[i];
                m->fu[i] = um->fu[i];
            }
            return true;
        }
    }
}

// ---- collocation ------------------------------------------------------------------------------------
static const double PI = 3.14159265358979323846;

// value of the Legendre polynomials of degree n-1 and n at x
static inline void legendre2(int n, double x, double& lower, double& upper) {
    double pa = 1.0;  // P_0
    double pb = x;    // P_1
    int m = 1;
    while (m < n) {
        double pc = ((2.0 * m + 1.0) * x * pb - m * pa) / (m + 1.0);
        pa = pb;
        pb = pc;
        ++m;
    }
    lower = pa;
    upper = pb;
}

static void lobatto_symmetric_nodes(int N, bool newton, std::vector<double>& tau) {
    const int No = N - 1;
    tau.assign(N, 0.0);
    tau[0] = -1.0;
    tau[No] = 1.0;
    for (int k = 1; 2 * k < No; ++k) {
        double x = -std::cos(PI * k / No);  // Chebyshev-Gauss-Lobatto point (also the Newton seed)
        if (newton) {
            // roots of (1-x^2) P'_No(x)
            int it = 0;
            while (it < 100) {
                double lo, up;
                legendre2(No, x, lo, up);
                double dx = (x * up - lo) / (N * up);
                x = x - dx;
                ++it;
                if (std::fabs(dx) <= 1e-16) break;
            }
        }
        tau[k] = x;
        tau[No - k] = -x;
    }
    if ((No & 1) == 0) tau[No / 2] = 0.0;
}

bool build_collocation(int kind, int N, Collocation* c, std::string* err) {
    if (N < 2) {
        if (err) *err = "a phase needs at least 2 collocation nodes";
        return false;
    }
    if (kind != ECUDA_LEGENDRE && kind != ECUDA_CHEBYSHEV) {
        if (err) *err = "unknown collocation method";
        return false;
    }
    const int No = N - 1;
    c->N = N;
    c->w.assign(N, 0.0);
    c->D.assign(static_cast<size_t>(N) * N, 0.0);
    double* D = c->D.data();
    if (kind == ECUDA_LEGENDRE) {
        lobatto_symmetric_nodes(N, true, c->tau);
        std::vector<double> PN(N);
        for (int k = 0; k < N; ++k) {
            double lo, up;
            legendre2(No, c->tau[k], lo, up);
            PN[k] = up;
            c->w[k] = 2.0 / (No * (No + 1.0) * (up * up));
        }
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < N; ++j) {
                if (j == k) continue;
                D[static_cast<size_t>(k) * N + j] = (PN[k] / PN[j]) / (c->tau[k] - c->tau[j]);
            }
        }
        D[0] = -(No * (No + 1.0)) / 4.0;
        D[static_cast<size_t>(No) * N + No] = (No * (No + 1.0)) / 4.0;
        return true;
    }
    lobatto_symmetric_nodes(N, false, c->tau);
    const double* t = c->tau.data();
    for (int k = 0; k < N; ++k) {
        const double ck = (k == 0 || k == No) ? 2.0 : 1.0;
        for (int j = 0; j < N; ++j) {
            if (j == k) continue;
            const double cj = (j == 0 || j == No) ? 2.0 : 1.0;
            const double sgn = ((k + j) & 1) ? -1.0 : 1.0;
            D[static_cast<size_t>(k) * N + j] = ((ck / cj) * sgn) / (t[k] - t[j]);
        }
        if (k != 0 && k != No) D[static_cast<size_t>(k) * N + k] = -t[k] / (2.0 * (1.0 - t[k] * t[k]));
    }
    D[0] = -(2.0 * No * No + 1.0) / 6.0;
    D[static_cast<size_t>(No) * N + No] = (2.0 * No * No + 1.0) / 6.0;
    // Clenshaw-Curtis quadrature weights
    const bool even = (No % 2 == 0);
    const double edge = even ? 1.0 / (static_cast<double>(No) * No - 1.0) : 1.0 / (static_cast<double>(No) * No);
    c->w[0] = edge;
    c->w[No] = edge;
    for (int i = 1; i < No; ++i) {
        const double th = PI * i / No;
        double v = 1.0;
        const int kmax = even ? No / 2 - 1 : (No - 1) / 2;
        for (int k = 1; k <= kmax; ++k) v = v - 2.0 * std::cos(2.0 * k * th) / (4.0 * k * k - 1.0);
        if (even) v = v - std::cos(No * th) / (static_cast<double>(No) * No - 1.0);
        c->w[i] = 2.0 * v / No;
    }
    return true;
}

// ---- layout -----------------------------------------------------------------------------------------
bool build_layout(const ecuda_problem_desc& d, HostProblem* hp, std::string* err) {
    auto fail = [&](const char* m) {
        if (err) *err = m;
        return false;
    };
    if (!model_info(d.model, &hp->mi)) return fail("unknown model id");
    if (d.nphases < 1 || d.nphases > ECUDA_MAX_PHASES) return fail("nphases out of range");
    if (d.batch < 1) return fail("batch must be >= 1");
    if (d.index_base != 0 && d.index_base != 1) return fail("index_base must be 0 or 1");
    if (d.pattern_mode != ECUDA_PATTERN_DENSE_NODE && d.pattern_mode != ECUDA_PATTERN_MODEL_DEPS)
        return fail("unknown pattern mode");
    if (d.collocation != ECUDA_LEGENDRE && d.collocation != ECUDA_CHEBYSHEV)
        return fail("unknown collocation method");
    hp->desc = d;
    hp->ns = hp->mi.ns;
    hp->nc = d.ncontrols > 0 ? d.ncontrols : hp->mi.nc_default;
    if (hp->nc < hp->mi.nc_default || hp->nc > ECUDA_MAX_CONTROLS) return fail("ncontrols out of range for model");
    if (d.model != ECUDA_MODEL_SI2D && hp->nc != hp->mi.nc_default) return fail("ncontrols is fixed for this model");
    if (d.ntracks < 0 || (d.model != ECUDA_MODEL_SI2D && d.model < ECUDA_MODEL_USER_BASE && d.ntracks != 0))
        return fail("moving-obstacle tracks are only defined for the si2d model and user models");
    if (d.ntracks > 0 && d.nwaypoints < 2) return fail("tracks need at least 2 waypoints");
    hp->ne = 2 * hp->ns;  // ePSOPT.cpp:43
    hp->nphases = d.nphases;
    hp->N.clear(); hp->npath.clear(); hp->nstat.clear(); hp->zoff.clear(); hp->goff.clear();
    hp->nvars_p.clear(); hp->inst_off.clear();
    int z = 0, g = 0, io = 0;
    const UserModel* um = d.model >= ECUDA_MODEL_USER_BASE ? user_model(d.model) : nullptr;
    const int nuser = um ? static_cast<int>(um->row_out.size()) : 0;  // traced path rows of a user model
    for (int p = 0; p < d.nphases; ++p) {
        const int N = d.nnodes[p];
        if (N < 2) return fail("a phase needs at least 2 collocation nodes");
        if (d.nstatic[p] < 0) return fail("negative obstacle count");
        const int np = d.nstatic[p] + d.ntracks + nuser;
        hp->N.push_back(N);
        hp->nstat.push_back(d.nstatic[p]);
        hp->npath.push_back(np);
        hp->zoff.push_back(z);
        hp->goff.push_back(g);
        hp->inst_off.push_back(io);
        const int nv = (hp->ns + hp->nc) * N + 2;
        hp->nvars_p.push_back(nv);
        z += nv;
        g += hp->ns * N + hp->ne + np * N + 1;
        io += d.nstatic[p] * hp->mi.rec_size;
    }
    hp->track_off = io;
    const int track_size = d.ntracks > 0 ? 1 + 3 * d.nwaypoints : 0;
    io += d.ntracks * track_size;
    const int stride = ((io + 3) / 4) * 4;  // 32-byte multiple: bulk-copy friendly
    hp->linkoff = g;
    ecuda_dims& dm = hp->dims;
    dm.nvars = z;
    dm.nlinkages = (d.nphases - 1) * (hp->ns + 1);
    dm.ncons = g + dm.nlinkages;
    dm.nstates = hp->ns;
    dm.ncontrols = hp->nc;
    dm.inst_stride = stride > 0 ? stride : 4;
    dm.rec_size = hp->mi.rec_size;
    dm.track_size = track_size;
    dm.nnz = 0;
    dm.ngroups = 0;
    // node-local rank tables
    const bool dense = d.pattern_mode == ECUDA_PATTERN_DENSE_NODE;
    for (int j = 0; j < ECUDA_MAX_STATES; ++j) {
        hp->xcnt[j] = 0;
        for (int i = 0; i < ECUDA_MAX_STATES; ++i) hp->xrank[j][i] = -1;
    }
    for (int j = 0; j < ECUDA_MAX_CONTROLS; ++j) {
        hp->ucnt[j] = 0;
        for (int i = 0; i < ECUDA_MAX_STATES; ++i) hp->urank[j][i] = -1;
    }
    for (int j = 0; j < hp->ns; ++j)
        for (int i = 0; i < hp->ns; ++i)
            if (dense || i == j || (hp->mi.fx[i] >> j & 1u)) hp->xrank[j][i] = static_cast<signed char>(hp->xcnt[j]++);
    for (int j = 0; j < hp->nc; ++j)
        for (int i = 0; i < hp->ns; ++i)
            if (dense || (hp->mi.fu[i] >> j & 1u)) hp->urank[j][i] = static_cast<signed char>(hp->ucnt[j]++);
    return true;
}

// ---- sparsity pattern + column groups ---------------------------------------------------------------
// Rows are emitted in ascending order column by column, so the triplets come out sorted by
// (col,row) without a sort pass.
void build_structure(HostProblem* hp) {
    const int ns = hp->ns, nc = hp->nc, ne = hp->ne, P = hp->nphases;
    const int nvars = hp->dims.nvars, ncons = hp->dims.ncons;
    hp->irow.clear();
    hp->jcol.clear();
    hp->colptr.assign(1, 0);
    auto& R = hp->irow;
    // T[e]: where the exact value of triplet e comes from (desc_pack: D entry or per-instance table entry)
    std::vector<unsigned> T;
    auto add = [&](int row, unsigned tab) {
        R.push_back(row);
        T.push_back(tab);
    };
    auto close_col = [&]() {
        int c = static_cast<int>(hp->colptr.size()) - 1;
        for (size_t e = hp->colptr.back(); e < R.size(); ++e) hp->jcol.push_back(c);
        hp->colptr.push_back(static_cast<int32_t>(R.size()));
    };
    auto link_row = [&](int a, int i) { return hp->linkoff + a * (ns + 1) + i; };
    for (int p = 0; p < P; ++p) {
        const int N = hp->N[p], np = hp->npath[p], nstat = hp->nstat[p], g0 = hp->goff[p];
        const int r_ev = g0 + ns * N, r_path = r_ev + ne, r_last = r_path + np * N;
        const unsigned W = desc_rowtab_width(ns, nc), TP = desc_path_off(ns, nc, N), TD = desc_d_off(ns, nc, N, np);
        const unsigned ONE = 0u, MINUS = 1u;
        auto rowtab = [&](int k, int i, int j) { return 2u + static_cast<unsigned>(k * ns + i) * W + static_cast<unsigned>(j); };
        for (int k = 0; k < N; ++k)  // control columns, node-major
            for (int j = 0; j < nc; ++j) {
                for (int i = 0; i < ns; ++i)
                    if (hp->urank[j][i] >= 0) add(g0 + k * ns + i, rowtab(k, i, ns + j));
                close_col();
            }
        for (int l = 0; l < N; ++l)  // state columns, node-major
            for (int j = 0; j < ns; ++j) {
                for (int k = 0; k < N; ++k) {
                    if (k != l) {
                        add(g0 + k * ns + j, TD + static_cast<unsigned>(l * N + k));
                    } else {
                        for (int i = 0; i < ns; ++i)
                            if (hp->xrank[j][i] >= 0) add(g0 + k * ns + i, rowtab(k, i, j));
                    }
                }
                if (l == 0) add(r_ev + j, ONE);
                if (l == N - 1) add(r_ev + ns + j, ONE);
                if (hp->mi.path_x >> j & 1u)
                    for (int q = 0; q < np; ++q) add(r_path + l * np + q, TP + static_cast<unsigned>((l * np + q) * 4 + j));
                if (l == N - 1 && p + 1 < P) add(link_row(p, j), ONE);
                if (l == 0 && p > 0) add(link_row(p - 1, j), MINUS);
                close_col();
            }
        for (int which = 0; which < 2; ++which) {  // t0 then tf
            for (int r = 0; r < ns * N; ++r) add(g0 + r, 2u + static_cast<unsigned>(r) * W + static_cast<unsigned>(ns + nc + which));
            for (int k = 0; k < N; ++k)
                for (int q = nstat; q < np; ++q)  // track rows read t
                    add(r_path + k * np + q, TP + static_cast<unsigned>((k * np + q) * 4 + 2 + which));
            add(r_last, which == 0 ? MINUS : ONE);
            if (which == 0 && p > 0) add(link_row(p - 1, ns), MINUS);
            if (which == 1 && p + 1 < P) add(link_row(p, ns), ONE);
            close_col();
        }
    }
    hp->dims.nnz = static_cast<int32_t>(R.size());
    // 16-bit rows / phase-local columns (ecuda_set_problem does not use the streaming kernel beyond that)
    hp->tdesc.resize(R.size());
    for (int p = 0; p < P; ++p)
        for (int c = hp->zoff[p]; c < hp->zoff[p] + hp->nvars_p[p]; ++c)
            for (int e = hp->colptr[c]; e < hp->colptr[c + 1]; ++e)
                hp->tdesc[e] = desc_pack(T[e], static_cast<unsigned>(R[e]) & 0xffffu, static_cast<unsigned>(c - hp->zoff[p]) & 0xffffu);
    // Curtis-Powell-Reid first-fit in natural column order; group row sets kept as 64-bit masks
    const int words = (ncons + 63) / 64;
    std::vector<std::vector<uint64_t>> cover;
    hp->group_of_col.assign(nvars, -1);
    std::vector<uint64_t> mine(words);
    for (int c = 0; c < nvars; ++c) {
        std::fill(mine.begin(), mine.end(), 0ull);
        for (int e = hp->colptr[c]; e < hp->colptr[c + 1]; ++e) mine[R[e] >> 6] |= 1ull << (R[e] & 63);
        size_t g = 0;
        for (; g < cover.size(); ++g) {
            bool hit = false;
            const uint64_t* cv = cover[g].data();
            for (int wd = 0; wd < words; ++wd)
                if (cv[wd] & mine[wd]) {
                    hit = true;
                    break;
                }
            if (!hit) break;
        }
        if (g == cover.size()) cover.emplace_back(words, 0ull);
        for (int wd = 0; wd < words; ++wd) cover[g][wd] |= mine[wd];
        hp->group_of_col[c] = static_cast<int32_t>(g);
    }
    hp->dims.ngroups = static_cast<int32_t>(cover.size());
}

void fill_probdev(const HostProblem& hp, ProbDev* out) {
    ProbDev& pd = *out;
    std::memset(&pd, 0, sizeof(pd));
    const ecuda_problem_desc& d = hp.desc;
    pd.model = d.model;
    pd.ns = hp.ns;
    pd.nc = hp.nc;
    pd.ne = hp.ne;
    pd.nphases = hp.nphases;
    pd.nvars = hp.dims.nvars;
    pd.ncons = hp.dims.ncons;
    pd.nnz = hp.dims.nnz;
    pd.nlink = hp.dims.nlinkages;
    pd.linkoff = hp.linkoff;
    pd.ntracks = d.ntracks;
    pd.nway = d.nwaypoints;
    pd.track_off = hp.track_off;
    pd.track_size = hp.dims.track_size;
    pd.rec_size = hp.dims.rec_size;
    pd.inst_stride = hp.dims.inst_stride;
    pd.maximize = d.maximize ? 1 : 0;
    pd.dense = d.pattern_mode == ECUDA_PATTERN_DENSE_NODE;
    pd.sf = 1.0;
    std::memcpy(pd.xrank, hp.xrank, sizeof(pd.xrank));
    std::memcpy(pd.urank, hp.urank, sizeof(pd.urank));
    std::memcpy(pd.xcnt, hp.xcnt, sizeof(pd.xcnt));
    std::memcpy(pd.ucnt, hp.ucnt, sizeof(pd.ucnt));
    for (int p = 0; p < hp.nphases; ++p) {
        PhaseDev& ph = pd.ph[p];
        ph.N = hp.N[p];
        ph.npath = hp.npath[p];
        ph.nstat = hp.nstat[p];
        ph.nb = (hp.N[p] + ECUDA_DOT_BLOCK - 1) / ECUDA_DOT_BLOCK;
        ph.zoff = hp.zoff[p];
        ph.goff = hp.goff[p];
        ph.nvars = hp.nvars_p[p];
        ph.inst_off = hp.inst_off[p];
        ph.mN = fast_div_magic(ph.N);
        ph.mnp = fast_div_magic(ph.npath);
        ph.m2np = fast_div_magic(2 * ph.npath);
        ph.mns = fast_div_magic(ph.nstat);
        ph.mnt = fast_div_magic(ph.npath - ph.nstat);
    }
}

void build_jac_template(const HostProblem& hp, const double* isz, const double* sg, std::vector<double>* tmpl) {
    tmpl->assign(static_cast<size_t>(hp.dims.nnz), 0.0);
    const int ns = hp.ns, nc = hp.nc;
    for (int p = 0; p < hp.nphases; ++p) {
        const int N = hp.N[p];
        const double* D = hp.col[p].D.data();
        for (int l = 0; l < N; ++l)
            for (int j = 0; j < ns; ++j) {
                const int col = hp.zoff[p] + nc * N + l * ns + j;
                const int base = hp.colptr[col];
                for (int k = 0; k < N; ++k) {
                    if (k == l) continue;
                    const int row = hp.goff[p] + k * ns + j;
                    // same two roundings as the kernels: (sg * D) * (1/sz)
                    (*tmpl)[base + (k < l ? k : k + hp.xcnt[j] - 1)] = (sg[row] * D[static_cast<size_t>(k) * N + l]) * isz[col];
                }
            }
    }
}

// Triplets that are NOT covered by the exact-mode template (everything but the D-coupled off-diagonal entries):
// the per-instance part of an exact Jacobian, in ascending triplet order. See ecuda_eval_compact.
void build_local_index(const HostProblem& hp, std::vector<int32_t>* local) {
    std::vector<char> shared(static_cast<size_t>(hp.dims.nnz), 0);
    const int ns = hp.ns, nc = hp.nc;
    for (int p = 0; p < hp.nphases; ++p) {
        const int N = hp.N[p];
        for (int l = 0; l < N; ++l)
            for (int j = 0; j < ns; ++j) {
                const int base = hp.colptr[hp.zoff[p] + nc * N + l * ns + j];
                for (int k = 0; k < N; ++k)
                    if (k != l) shared[base + (k < l ? k : k + hp.xcnt[j] - 1)] = 1;
            }
    }
    local->clear();
    for (int e = 0; e < hp.dims.nnz; ++e)
        if (!shared[e]) local->push_back(e);
}

}  // namespace ecuda

// ---- extern "C" host helpers ---------------------------------------------------------------------------
using namespace ecuda;

// lower triangle of the Lagrangian Hessian, sorted by (column, row); see HessIO in ecuda_internal.hpp
void ecuda::build_hess_structure(const HostProblem& hp, std::vector<int32_t>* irow, std::vector<int32_t>* jcol) {
    irow->clear();
    jcol->clear();
    const int ns = hp.ns, nc = hp.nc;
    for (int p = 0; p < hp.nphases; ++p) {
        const int N = hp.N[p], z0 = hp.zoff[p], t0 = z0 + (ns + nc) * N, tf = t0 + 1;
        auto put = [&](int r, int c) {
            irow->push_back(r);
            jcol->push_back(c);
        };
        for (int k = 0; k < N; ++k)
            for (int j = 0; j < nc; ++j) {
                const int c = z0 + k * nc + j;
                for (int j2 = j; j2 < nc; ++j2) put(z0 + k * nc + j2, c);
                for (int i = 0; i < ns; ++i) put(z0 + nc * N + k * ns + i, c);
                put(t0, c);
                put(tf, c);
            }
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < ns; ++i) {
                const int c = z0 + nc * N + k * ns + i;
                for (int i2 = i; i2 < ns; ++i2) put(z0 + nc * N + k * ns + i2, c);
                put(t0, c);
                put(tf, c);
            }
        put(t0, t0);
        put(tf, t0);
        put(tf, tf);
    }
}

extern "C" {

int ecuda_host_hess_structure(const ecuda_problem_desc* desc, int32_t* nnz_h, int32_t* iRow, int32_t* jCol) {
    if (!desc) return ECUDA_ERR_ARG;
    ecuda::HostProblem hp;
    std::string err;
    if (!ecuda::build_layout(*desc, &hp, &err)) return ECUDA_ERR_ARG;
    std::vector<int32_t> ir, jc;
    ecuda::build_hess_structure(hp, &ir, &jc);
    if (nnz_h) *nnz_h = static_cast<int32_t>(ir.size());
    for (size_t e = 0; e < ir.size(); ++e) {
        if (iRow) iRow[e] = ir[e] + desc->index_base;
        if (jCol) jCol[e] = jc[e] + desc->index_base;
    }
    return ECUDA_OK;
}

int ecuda_abi_version(void) { return ECUDA_ABI_VERSION; }

int ecuda_si2d_edge_records(const double* corners_xy, int ncorners, double* rec6) {
    if (!corners_xy || !rec6 || ncorners < 2) return ECUDA_ERR_ARG;
    for (int e = 0; e < ncorners; ++e) {
        const int n = (e + 1 == ncorners) ? 0 : e + 1;  // last edge closes the polygon
        const double xa = corners_xy[2 * e], ya = corners_xy[2 * e + 1];
        const double xb = corners_xy[2 * n], yb = corners_xy[2 * n + 1];
        const double xc = (xb + xa) / 2.;
        const double slope = (yb - ya) / (xb - xa);
        const double yc = ya + slope * (xc - xa);
        const double radsq = std::pow(xc - xa, 2.0) + std::pow(yc - ya, 2.0);
        const double tt = -1.0 * std::atan2(yc - ya, xc - xa);
        double* r = rec6 + 6 * e;
        r[0] = xc;
        r[1] = yc;
        r[2] = std::cos(tt);
        r[3] = std::sin(tt);
        r[4] = radsq;       // asq
        r[5] = .2 * radsq;  // bsq
    }
    return ECUDA_OK;
}

int ecuda_host_dims(const ecuda_problem_desc* desc, ecuda_dims* out) {
    if (!desc || !out) return ECUDA_ERR_ARG;
    HostProblem hp;
    std::string err;
    if (!build_layout(*desc, &hp, &err)) return ECUDA_ERR_ARG;
    build_structure(&hp);
    *out = hp.dims;
    return ECUDA_OK;
}

int ecuda_host_structure(const ecuda_problem_desc* desc, int32_t* iRow, int32_t* jCol, int32_t* group_of_col) {
    if (!desc) return ECUDA_ERR_ARG;
    HostProblem hp;
    std::string err;
    if (!build_layout(*desc, &hp, &err)) return ECUDA_ERR_ARG;
    build_structure(&hp);
    const int base = desc->index_base;
    for (int e = 0; e < hp.dims.nnz; ++e) {
        if (iRow) iRow[e] = hp.irow[e] + base;
        if (jCol) jCol[e] = hp.jcol[e] + base;
    }
    if (group_of_col) std::memcpy(group_of_col, hp.group_of_col.data(), sizeof(int32_t) * hp.dims.nvars);
    return ECUDA_OK;
}

int ecuda_host_compact_structure(const ecuda_problem_desc* desc, int32_t* nlocal, int32_t* local_index) {
    if (!desc) return ECUDA_ERR_ARG;
    HostProblem hp;
    std::string err;
    if (!build_layout(*desc, &hp, &err)) return ECUDA_ERR_ARG;
    build_structure(&hp);
    std::vector<int32_t> li;
    build_local_index(hp, &li);
    if (nlocal) *nlocal = static_cast<int32_t>(li.size());
    if (local_index) std::memcpy(local_index, li.data(), sizeof(int32_t) * li.size());
    return ECUDA_OK;
}

int ecuda_host_collocation(int kind, int nnodes, double* tau, double* w, double* D) {
    Collocation c;
    std::string err;
    if (!build_collocation(kind, nnodes, &c, &err)) return ECUDA_ERR_ARG;
    if (tau) std::memcpy(tau, c.tau.data(), sizeof(double) * nnodes);
    if (w) std::memcpy(w, c.w.data(), sizeof(double) * nnodes);
    if (D) std::memcpy(D, c.D.data(), sizeof(double) * nnodes * nnodes);
    return ECUDA_OK;
}

}  // extern "C"
