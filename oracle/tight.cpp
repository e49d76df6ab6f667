// tight.cpp (oracle) -- the same arithmetic as evaluate.cpp without the type erasure: plain double
// loops, obstacle geometry hoisted out of the node loop. TEST INFRASTRUCTURE (see oracle.hpp).
//
// Purpose: (1) a second, independently written CPU implementation that must agree BIT FOR BIT with
// the reference-style one (tests/test_oracle_styles.py); (2) the "CPU-tight" baseline of
// BASELINE.md section 2, reported next to the reference-style number for honesty.
// Model math: src/Examples/PSOPT/etol_psopt_example1.cpp:101-258; driver semantics:
// src/ePSOPT/ePSOPT.cpp:186-306; transcription: SURVEY.md Appendix A.
#include <cmath>
#include <stdexcept>

#include "../include/ecuda_detmath.h"
#include "oracle.hpp"

namespace oracle {

namespace {

struct EdgeRec {
    double xc, yc, ct, st, asq, bsq;
};
struct PhasePre {
    std::vector<EdgeRec> edges;
    std::vector<double> cyl;  // cx, cy, r2
};

std::vector<PhasePre> precompute(const Problem& P, const Instance& I) {
    std::vector<PhasePre> out(P.L.nphases);
    for (int p = 0; p < P.L.nphases; ++p) {
        for (const Border& bd : I.phases[p].borders) {
            size_t n = bd.size();
            for (size_t i = 0; i < n; ++i) {
                EdgeGeom g = edge_geometry(bd[i], bd[(i + 1) % n]);
                out[p].edges.push_back({g.xc, g.yc, std::cos(g.tt), std::sin(g.tt), g.asq, g.bsq});
            }
        }
        for (const Cylinder& c : I.phases[p].cylinders) {
            out[p].cyl.push_back(c.cx);
            out[p].cyl.push_back(c.cy);
            out[p].cyl.push_back(c.r * c.r);
        }
    }
    return out;
}

double interp(double t, const std::vector<double>& tv, const std::vector<double>& ref) {
    size_t j = 0;
    if (t > tv.back()) {
        j = tv.size() - 2;
    } else if (t >= tv.front()) {
        for (size_t c = 0; c + 1 < tv.size(); ++c)
            if (t >= tv[c] && t <= tv[c + 1]) j = c;
    }
    return (t - tv[j]) * (ref[j + 1] - ref[j]) / (tv[j + 1] - tv[j]) + ref[j];
}

// plain-double replay of a user tape (the tight counterpart of replay<T> in models.hpp)
void tape_values(const UserTape& ut, const double* x, const double* u, double t, std::vector<double>& v) {
    v.resize(ut.nodes.size());
    for (size_t i = 0; i < ut.nodes.size(); ++i) {
        const TapeNode& n = ut.nodes[i];
        switch (n.op) {
            case T_INPUT: v[i] = n.a < ut.ns ? x[n.a] : n.a < ut.ns + ut.nc ? u[n.a - ut.ns] : t; break;
            case T_CONST: v[i] = n.imm; break;
            case T_ADD: v[i] = v[n.a] + v[n.b]; break;
            case T_SUB: v[i] = v[n.a] - v[n.b]; break;
            case T_MUL: v[i] = v[n.a] * v[n.b]; break;
            case T_DIV: v[i] = v[n.a] / v[n.b]; break;
            case T_NEG: v[i] = -v[n.a]; break;
            case T_SQRT: v[i] = std::sqrt(v[n.a]); break;
            case T_EXP: v[i] = std::exp(v[n.a]); break;
            case T_POW: {
                const double e = n.imm;
                if (std::fabs(e) <= 8.0 && e == std::floor(e)) {
                    const int c = static_cast<int>(std::fabs(e));
                    double p = c == 0 ? 1.0 : v[n.a];
                    for (int r = 1; r < c; ++r) p = p * v[n.a];
                    v[i] = e < 0 ? 1.0 / p : p;
                } else {
                    v[i] = std::pow(v[n.a], e);
                }
                break;
            }
            default: {
                double s, c;
                ecuda_sincos(v[n.a], &s, &c);
                v[i] = n.op == T_SIN ? s : c;
            }
        }
    }
}

void dynamics(const Spec& spec, const double* x, const double* u, double t, double* f) {
    const int model = spec.model;
    if (model == USER) {
        static thread_local std::vector<double> v;
        tape_values(spec.user, x, u, t, v);
        for (int i = 0; i < spec.user.ns; ++i) f[i] = v[spec.user.f_out[i]];
    } else if (model == SI2D) {
        f[0] = u[0];
        f[1] = u[1];
    } else if (model == PM3D) {
        f[0] = x[3];
        f[1] = x[4];
        f[2] = x[5];
        f[3] = u[0];
        f[4] = u[1];
        f[5] = u[2];
    } else {
        double sg, cg, sp, cp;
        ecuda_sincos(x[4], &sg, &cg);
        ecuda_sincos(x[5], &sp, &cp);
        f[0] = (x[3] * cg) * cp;
        f[1] = (x[3] * cg) * sp;
        f[2] = x[3] * sg;
        f[3] = u[0] - 9.80665 * sg;
        f[4] = u[1];
        f[5] = u[2];
    }
}

void path_rows(const Problem& P, const PhasePre& pre, const Instance& I, const double* x, double t,
               double* out) {
    int q = 0;
    const bool user = P.spec.model == USER;
    if (P.spec.model == SI2D || user) {
        if (user)
            for (size_t c = 0; c < pre.cyl.size(); c += 3) {
                double dx = x[0] - pre.cyl[c], dy = x[1] - pre.cyl[c + 1];
                out[q++] = pre.cyl[c + 2] - (dx * dx + dy * dy);
            }
        for (const EdgeRec& e : pre.edges) {
            double dx = x[0] - e.xc, dy = x[1] - e.yc;
            double delx = e.ct * dx - e.st * dy;
            double dely = e.st * dx + e.ct * dy;
            out[q++] = e.asq * e.bsq - (e.bsq * (delx * delx) + e.asq * (dely * dely));
        }
        for (const Track& tr : I.tracks) {
            double xc = interp(t, tr.t, tr.x), yc = interp(t, tr.t, tr.y);
            double dx = x[0] - xc, dy = x[1] - yc;
            double dist = dx * dx + dy * dy;
            out[q++] = dist * (-1.) + tr.radius * tr.radius;
        }
        if (user && !P.spec.user.row_out.empty()) {  // traced rows read states 0, 1 and t only
            static thread_local std::vector<double> v;
            const double xs[8] = {x[0], x[1], 0, 0, 0, 0, 0, 0}, us[8] = {0};
            tape_values(P.spec.user, xs, us, t, v);
            for (int id : P.spec.user.row_out) out[q++] = v[id];
        }
    } else {
        for (size_t c = 0; c < pre.cyl.size(); c += 3) {
            double dx = x[0] - pre.cyl[c], dy = x[1] - pre.cyl[c + 1];
            out[q++] = pre.cyl[c + 2] - (dx * dx + dy * dy);
        }
    }
}

double running_cost(const Spec& spec, const double* x, const double* u, double t, bool maximize) {
    const int model = spec.model;
    double l;
    if (model == USER) {
        static thread_local std::vector<double> v;
        tape_values(spec.user, x, u, t, v);
        l = v[spec.user.cost_out];
    } else {
        l = (model == SI2D) ? u[0] * u[0] + u[1] * u[1] : (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2];
    }
    return maximize ? -1.0 * l : l;
}

void g_tight(const Problem& P, const std::vector<PhasePre>& pre, const Instance& I, const double* zs,
             double* g, std::vector<double>& z, std::vector<double>& F) {
    const Layout& L = P.L;
    const int ns = L.ns;
    for (int c = 0; c < L.nvars; ++c) z[c] = zs[c] * P.sc.isz[c];
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], np = L.npath[p];
        const Collocation& C = P.col[p];
        double t0 = z[L.it0(p)], tf = z[L.itf(p)];
        double h = 0.5 * (tf - t0), m = 0.5 * (tf + t0);
        double pathv[256];
        if (np > 256) throw std::invalid_argument("tight oracle supports <= 256 path rows per node");
        for (int k = 0; k < N; ++k) {
            double t = h * C.tau[k] + m;
            dynamics(P.spec, &z[L.ix(p, k, 0)], &z[L.iu(p, k, 0)], t, &F[static_cast<size_t>(k) * ns]);
            path_rows(P, pre[p], I, &z[L.ix(p, k, 0)], t, pathv);
            for (int q = 0; q < np; ++q) g[L.rpath(p, k, q)] = P.sc.sg[L.rpath(p, k, q)] * pathv[q];
        }
        const double* X = &z[L.ix(p, 0, 0)];
        for (int k = 0; k < N; ++k) {
            const double* Dr = &C.D[static_cast<size_t>(k) * N];
            for (int i = 0; i < ns; ++i) {
                double total = 0.0;
                for (int b0 = 0; b0 < N; b0 += DOT_BLOCK) {
                    int b1 = b0 + DOT_BLOCK < N ? b0 + DOT_BLOCK : N;
                    double s = 0.0;
                    for (int l = b0; l < b1; ++l) s = std::fma(Dr[l], X[static_cast<size_t>(l) * ns + i], s);
                    total = (b0 == 0) ? s : total + s;
                }
                double zeta = total - h * F[static_cast<size_t>(k) * ns + i];
                g[L.rdef(p, k, i)] = P.sc.sg[L.rdef(p, k, i)] * zeta;
            }
        }
        for (int i = 0; i < ns; ++i) {
            g[L.rev(p, i)] = P.sc.sg[L.rev(p, i)] * z[L.ix(p, 0, i)];
            g[L.rev(p, ns + i)] = P.sc.sg[L.rev(p, ns + i)] * z[L.ix(p, N - 1, i)];
        }
        g[L.rlast(p)] = P.sc.sg[L.rlast(p)] * (tf - t0);
    }
    for (int a = 0; a + 1 < L.nphases; ++a) {
        for (int i = 0; i < ns; ++i)
            g[L.rlink(a, i)] = P.sc.sg[L.rlink(a, i)] * (z[L.ix(a, L.N[a] - 1, i)] - z[L.ix(a + 1, 0, i)]);
        g[L.rlink(a, ns)] = P.sc.sg[L.rlink(a, ns)] * (z[L.itf(a)] - z[L.it0(a + 1)]);
    }
}

}  // namespace

// ---- mesh refinement support: relative local discretisation error per mesh interval and interpolation
// onto another mesh. PSOPT's own code is not in the reference tree (ePSOPT only sets mesh_refinement =
// "automatic", ode_tolerance, mr_max_iterations: src/ePSOPT/ePSOPT.cpp:69-71); restated is the estimate it
// documents (Betts): eta_ik = int_{t_k}^{t_k+1} |x~_i' - f_i(x~, u~)| dt by 4-point Gauss-Legendre, states
// and controls interpolated by the Lagrange polynomial through the nodes; eps_k = max_i eta_ik / (w_i + 1),
// w_i = max_k max(|x_ik|, |x'_ik|).
namespace {
const double kGX[4] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};
const double kGW[4] = {0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538};

std::vector<double> barycentric(const std::vector<double>& tau) {
    std::vector<double> w(tau.size());
    for (size_t l = 0; l < tau.size(); ++l) {
        double p = 1.0;
        for (size_t m = 0; m < tau.size(); ++m)
            if (m != l) p = p * (tau[l] - tau[m]);
        w[l] = 1.0 / p;
    }
    return w;
}
// values (and tau-derivatives) of the Lagrange basis at t
void basis_at(const std::vector<double>& tau, const std::vector<double>& bw, double t, std::vector<double>& L,
              std::vector<double>* dL) {
    const size_t N = tau.size();
    L.assign(N, 0.0);
    if (dL) dL->assign(N, 0.0);
    for (size_t hit = 0; hit < N; ++hit) {
        if (t != tau[hit]) continue;
        L[hit] = 1.0;
        if (dL) {
            double s = 0.0;
            for (size_t l = 0; l < N; ++l)
                if (l != hit) {
                    (*dL)[l] = (bw[l] / bw[hit]) / (tau[hit] - tau[l]);
                    s = s + (*dL)[l];
                }
            (*dL)[hit] = -s;
        }
        return;
    }
    double s = 0.0, s2 = 0.0;
    for (size_t l = 0; l < N; ++l) {
        double r = bw[l] / (t - tau[l]);
        s = s + r;
        s2 = s2 + r / (t - tau[l]);
    }
    for (size_t l = 0; l < N; ++l) {
        L[l] = (bw[l] / (t - tau[l])) / s;
        if (dL) (*dL)[l] = L[l] * (s2 / s - 1.0 / (t - tau[l]));
    }
}
}  // namespace

void Problem::ode_error(const Instance& I, const double* zs, double* err) const {
    (void)I;
    std::vector<double> z(L.nvars);
    for (int c = 0; c < L.nvars; ++c) z[c] = zs[c] * sc.isz[c];
    const int ns = L.ns, nc = L.nc;
    int eoff = 0;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p];
        const Collocation& C = col[p];
        const double t0 = z[L.it0(p)], tf = z[L.itf(p)], h = 0.5 * (tf - t0);
        const double* X = &z[L.ix(p, 0, 0)];
        const double* U = &z[L.iu(p, 0, 0)];
        std::vector<double> w(ns, 0.0);
        for (int i = 0; i < ns; ++i)
            for (int k = 0; k < N; ++k) {
                const double* Dr = &C.D[static_cast<size_t>(k) * N];
                double total = 0.0;
                for (int b0 = 0; b0 < N; b0 += DOT_BLOCK) {
                    int b1 = b0 + DOT_BLOCK < N ? b0 + DOT_BLOCK : N;
                    double s = 0.0;
                    for (int l = b0; l < b1; ++l) s = std::fma(Dr[l], X[static_cast<size_t>(l) * ns + i], s);
                    total = (b0 == 0) ? s : total + s;
                }
                w[i] = std::fmax(w[i], std::fabs(X[static_cast<size_t>(k) * ns + i]));
                w[i] = std::fmax(w[i], std::fabs(total / h));
            }
        std::vector<double> bw = barycentric(C.tau), E, dE;
        for (int k = 0; k + 1 < N; ++k) {
            const double half = 0.5 * (C.tau[k + 1] - C.tau[k]), mid = 0.5 * (C.tau[k + 1] + C.tau[k]);
            std::vector<double> eta(ns, 0.0);
            for (int q = 0; q < 4; ++q) {
                basis_at(C.tau, bw, mid + half * kGX[q], E, &dE);
                double xq[8] = {0}, dxq[8] = {0}, uq[8] = {0}, f[8];
                for (int l = 0; l < N; ++l) {
                    for (int i = 0; i < ns; ++i) {
                        xq[i] = std::fma(E[l], X[static_cast<size_t>(l) * ns + i], xq[i]);
                        dxq[i] = std::fma(dE[l], X[static_cast<size_t>(l) * ns + i], dxq[i]);
                    }
                    for (int j = 0; j < nc; ++j) uq[j] = std::fma(E[l], U[static_cast<size_t>(l) * nc + j], uq[j]);
                }
                dynamics(spec, xq, uq, h * (mid + half * kGX[q]) + 0.5 * (tf + t0), f);
                for (int i = 0; i < ns; ++i) eta[i] = std::fma(half * kGW[q], std::fabs(dxq[i] - h * f[i]), eta[i]);
            }
            double e = 0.0;
            for (int i = 0; i < ns; ++i) e = std::fmax(e, eta[i] / (w[i] + 1.0));
            err[eoff + k] = e;
        }
        eoff += N - 1;
    }
}

void Problem::resample(const double* zs, const std::vector<int>& nnew, const double* sz_new, std::vector<double>* out) const {
    const int ns = L.ns, nc = L.nc;
    out->clear();
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], Nn = nnew[p];
        Collocation to = make_collocation(spec.collocation, Nn);
        std::vector<double> bw = barycentric(col[p].tau), R;
        const size_t o = out->size();
        out->resize(o + static_cast<size_t>(ns + nc) * Nn + 2);
        double* zn = out->data() + o;
        for (int k = 0; k < Nn; ++k) {
            basis_at(col[p].tau, bw, to.tau[k], R, nullptr);
            for (int j = 0; j < nc; ++j) {
                double a = 0.0;
                for (int l = 0; l < N; ++l) a = std::fma(R[l], zs[L.iu(p, l, j)] * sc.isz[L.iu(p, l, j)], a);
                zn[k * nc + j] = a;
            }
            for (int i = 0; i < ns; ++i) {
                double a = 0.0;
                for (int l = 0; l < N; ++l) a = std::fma(R[l], zs[L.ix(p, l, i)] * sc.isz[L.ix(p, l, i)], a);
                zn[nc * Nn + k * ns + i] = a;
            }
        }
        zn[(ns + nc) * Nn] = zs[L.it0(p)] * sc.isz[L.it0(p)];
        zn[(ns + nc) * Nn + 1] = zs[L.itf(p)] * sc.isz[L.itf(p)];
    }
    if (sz_new)
        for (size_t c = 0; c < out->size(); ++c) (*out)[c] = (*out)[c] * sz_new[c];
}

void Problem::eval_g_tight(const Instance& I, const double* zs, double* g) const {
    auto pre = precompute(*this, I);
    std::vector<double> z(L.nvars), F;
    int maxN = 0;
    for (int n : L.N) maxN = n > maxN ? n : maxN;
    F.resize(static_cast<size_t>(maxN) * L.ns);
    g_tight(*this, pre, I, zs, g, z, F);
}

void Problem::eval_f_tight(const Instance& I, const double* zs, double* f) const {
    double total = 0.0;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p];
        double t0 = zs[L.it0(p)] * sc.isz[L.it0(p)], tf = zs[L.itf(p)] * sc.isz[L.itf(p)];
        double h = 0.5 * (tf - t0), tmid = 0.5 * (tf + t0);
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
            double u[8], x[8];
            for (int j = 0; j < L.nc; ++j) u[j] = zs[L.iu(p, k, j)] * sc.isz[L.iu(p, k, j)];
            for (int i = 0; i < L.ns; ++i) x[i] = zs[L.ix(p, k, i)] * sc.isz[L.ix(p, k, i)];
            acc = std::fma(col[p].w[k], running_cost(spec, x, u, h * col[p].tau[k] + tmid, spec.maximize), acc);
        }
        double fp = h * acc;
        total = (p == 0) ? fp : total + fp;
    }
    *f = sc.sf * total;
}

void Problem::eval_jac_fd_tight(const Instance& I, const double* zs, double* vals) const {
    auto pre = precompute(*this, I);
    const double sqrt_eps = 1.4901161193847656e-08;
    std::vector<double> zp(L.nvars), zm(L.nvars), gp(L.ncons), gm(L.ncons), rinv(L.nvars), z(L.nvars), F;
    int maxN = 0;
    for (int n : L.N) maxN = n > maxN ? n : maxN;
    F.resize(static_cast<size_t>(maxN) * L.ns);
    // columns of each group, built once
    std::vector<std::vector<int>> members(S.ngroups);
    for (int c = 0; c < L.nvars; ++c) members[S.group_of_col[c]].push_back(c);
    for (int c = 0; c < L.nvars; ++c) zp[c] = zm[c] = zs[c];
    for (int grp = 0; grp < S.ngroups; ++grp) {
        for (int c : members[grp]) {
            double delta = sqrt_eps * (1.0 + std::fabs(zs[c]));
            zp[c] = zs[c] + delta;
            zm[c] = zs[c] - delta;
            rinv[c] = 1.0 / (2.0 * delta);
        }
        g_tight(*this, pre, I, zp.data(), gp.data(), z, F);
        g_tight(*this, pre, I, zm.data(), gm.data(), z, F);
        for (int c : members[grp]) {
            for (int e = S.colptr[c]; e < S.colptr[c + 1]; ++e) vals[e] = (gp[S.irow[e]] - gm[S.irow[e]]) * rinv[c];
            zp[c] = zm[c] = zs[c];
        }
    }
}

}  // namespace oracle
