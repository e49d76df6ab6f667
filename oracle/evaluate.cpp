// evaluate.cpp (oracle) -- objective, constraint vector, exact and finite-difference Jacobian of
// the pseudospectral NLP, driven through per-node callbacks exactly the way PSOPT drives ePSOPT.
// TEST INFRASTRUCTURE (see oracle.hpp header).
//
// ETOL side restated from src/ePSOPT/ePSOPT.cpp:
//   node_dae       <- ePSOPT::dae            :218-276  (vector<any> of scalar pointers, one call per
//                                                        state derivative, one per constraint functor)
//   node_cost      <- ePSOPT::integrand_cost :186-216  (negated when maximising, :212-213)
//   events rows    <- ePSOPT::events         :281-291
//   endpoint cost  <- ePSOPT::endpoint_cost  :302-306  (always 0)
// PSOPT side (defects, quadrature, scaling, index-set differences) restated from SURVEY.md
// Appendix A.3-A.7. Canonical operation order: DESIGN.md section 3.
#include <cmath>
#include <stdexcept>

#include "models.hpp"
#include "oracle.hpp"

namespace oracle {

EdgeGeom edge_geometry(const Corner& a, const Corner& b) {
    // etol_psopt_example1.cpp:164-172,178-179 (plain double geometry of one polygon edge)
    double xa = a[0], ya = a[1], xb = b[0], yb = b[1];
    EdgeGeom g;
    g.xc = (xb + xa) / 2.;
    double m = (yb - ya) / (xb - xa);
    g.yc = ya + m * (g.xc - xa);
    g.radsq = std::pow(g.xc - xa, 2.0) + std::pow(g.yc - ya, 2.0);
    g.tt = -1.0 * std::atan2(g.yc - ya, g.xc - xa);
    g.asq = g.radsq;
    g.bsq = .2 * g.radsq;
    return g;
}

Problem::Problem(const Spec& s) : spec(s), L(make_layout(s)), S(make_structure(s, L)) {
    for (int p = 0; p < L.nphases; ++p) col.push_back(make_collocation(s.collocation, L.N[p]));
    set_scaling(nullptr, nullptr, 1.0);
}

void Problem::set_scaling(const double* sz, const double* sg, double sf) {
    sc.sz.assign(L.nvars, 1.0);
    sc.isz.assign(L.nvars, 1.0);
    sc.sg.assign(L.ncons, 1.0);
    if (sz)
        for (int c = 0; c < L.nvars; ++c) {
            sc.sz[c] = sz[c];
            sc.isz[c] = 1.0 / sz[c];
        }
    if (sg)
        for (int r = 0; r < L.ncons; ++r) sc.sg[r] = sg[r];
    sc.sf = sf;
}

namespace {

// D*X for one (row k, state i): blocks of DOT_BLOCK nodes, each a serial ascending fma chain from
// 0, block sums added serially in ascending order.
double blocked_dot(const double* Drow, const double* X, int stride, int N) {
    double total = 0.0;
    for (int b0 = 0, b = 0; b0 < N; b0 += DOT_BLOCK, ++b) {
        double p = 0.0;
        int b1 = b0 + DOT_BLOCK < N ? b0 + DOT_BLOCK : N;
        for (int l = b0; l < b1; ++l) p = std::fma(Drow[l], X[static_cast<size_t>(l) * stride], p);
        total = (b == 0) ? p : total + p;
    }
    return total;
}

// What ePSOPT::dae does at one node (ePSOPT.cpp:218-276), for scalar type T.
template <class T>
void node_dae(const Callbacks<T>& cb, int ns, int nc, T* derivatives, T* path, T* states, T* controls,
              T& t, double dt) {
    vector_t x, u;
    for (int i = 0; i < ns; ++i) x.push_back(&states[i]);
    for (int i = 0; i < nc; ++i) u.push_back(&controls[i]);
    T* tval = &t;
    for (int i = 0; i < ns; ++i) {
        const f_t& f = cb.gradient.at(i);
        vector_t params = {std::string()};
        std::vector<std::string> pnames = {std::string("")};
        scalar_t f_val = f(x, u, params, pnames, tval, dt);
        derivatives[i] = std::any_cast<T>(f_val);
    }
    size_t j = 0;
    for (size_t i = 0; i < cb.constraints.size(); ++i) {
        const f_t& p = cb.constraints.at(i);
        vector_t params = {std::string()};
        std::vector<std::string> pnames = {std::string("")};
        scalar_t p_val = p(x, u, params, pnames, tval, dt);
        std::vector<T> out = std::any_cast<std::vector<T>>(p_val);
        for (T val : out) path[j++] = val;
    }
}

// What ePSOPT::integrand_cost does (ePSOPT.cpp:186-216).
template <class T>
T node_cost(const Callbacks<T>& cb, int ns, int nc, T* states, T* controls, T& t, double dt,
            bool maximize) {
    vector_t x, u;
    for (int i = 0; i < ns; ++i) x.push_back(&states[i]);
    for (int i = 0; i < nc; ++i) u.push_back(&controls[i]);
    vector_t params = {std::string()};
    std::vector<std::string> pnames = {std::string("")};
    scalar_t fout = cb.objective(x, u, params, pnames, &t, dt);
    T f_val = std::any_cast<T>(fout);
    if (maximize) f_val = T(-1.0) * f_val;
    return f_val;
}

struct PhaseTimes {
    double t0, tf, h, m;
};

PhaseTimes phase_times(const Layout& L, int p, const std::vector<double>& z) {
    PhaseTimes pt;
    pt.t0 = z[L.it0(p)];
    pt.tf = z[L.itf(p)];
    pt.h = 0.5 * (pt.tf - pt.t0);
    pt.m = 0.5 * (pt.tf + pt.t0);
    return pt;
}

void unscale(const Problem& P, const double* zs, std::vector<double>* z) {
    z->resize(P.L.nvars);
    for (int c = 0; c < P.L.nvars; ++c) (*z)[c] = zs[c] * P.sc.isz[c];
}

void check_instance(const Problem& P, const Instance& I) {
    if (static_cast<int>(I.phases.size()) != P.L.nphases) throw std::invalid_argument("instance phases");
    for (int p = 0; p < P.L.nphases; ++p) {
        int nstat = 0;
        if (P.spec.model == SI2D || (P.spec.model == USER && P.spec.user.edges))
            for (auto& b : I.phases[p].borders) nstat += static_cast<int>(b.size());
        else
            nstat = static_cast<int>(I.phases[p].cylinders.size());
        if (nstat != P.spec.nstatic[p]) throw std::invalid_argument("instance static obstacle count");
    }
    if (static_cast<int>(I.tracks.size()) != P.spec.ntracks) throw std::invalid_argument("instance tracks");
}

// scaled constraint vector from a scaled decision vector, with prebuilt callbacks per phase
void g_with_callbacks(const Problem& P, const std::vector<Callbacks<double>>& cbs, const double* zs,
                      double* g) {
    const Layout& L = P.L;
    std::vector<double> z;
    unscale(P, zs, &z);
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], ns = L.ns, nc = L.nc, np = L.npath[p];
        const Collocation& C = P.col[p];
        PhaseTimes pt = phase_times(L, p, z);
        std::vector<double> F(static_cast<size_t>(N) * ns), path(np > 0 ? np : 1);
        for (int k = 0; k < N; ++k) {
            double t = pt.h * C.tau[k] + pt.m;
            node_dae<double>(cbs[p], ns, nc, &F[static_cast<size_t>(k) * ns], path.data(), &z[L.ix(p, k, 0)],
                             &z[L.iu(p, k, 0)], t, 0.0);
            for (int q = 0; q < np; ++q) g[L.rpath(p, k, q)] = P.sc.sg[L.rpath(p, k, q)] * path[q];
        }
        const double* X = &z[L.ix(p, 0, 0)];
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < ns; ++i) {
                double dot = blocked_dot(&C.D[static_cast<size_t>(k) * N], X + i, ns, N);
                double zeta = dot - pt.h * F[static_cast<size_t>(k) * ns + i];
                g[L.rdef(p, k, i)] = P.sc.sg[L.rdef(p, k, i)] * zeta;
            }
        for (int i = 0; i < ns; ++i) {
            g[L.rev(p, i)] = P.sc.sg[L.rev(p, i)] * z[L.ix(p, 0, i)];
            g[L.rev(p, ns + i)] = P.sc.sg[L.rev(p, ns + i)] * z[L.ix(p, N - 1, i)];
        }
        g[L.rlast(p)] = P.sc.sg[L.rlast(p)] * (pt.tf - pt.t0);
    }
    for (int a = 0; a + 1 < L.nphases; ++a) {
        for (int i = 0; i < L.ns; ++i)
            g[L.rlink(a, i)] = P.sc.sg[L.rlink(a, i)] * (z[L.ix(a, L.N[a] - 1, i)] - z[L.ix(a + 1, 0, i)]);
        g[L.rlink(a, L.ns)] = P.sc.sg[L.rlink(a, L.ns)] * (z[L.itf(a)] - z[L.it0(a + 1)]);
    }
}

template <class T>
std::vector<Callbacks<T>> build_callbacks(const Problem& P, const Instance& I) {
    std::vector<Callbacks<T>> cbs;
    for (int p = 0; p < P.L.nphases; ++p)
        cbs.push_back(make_callbacks<T>(P.spec.model, &I.phases[p], &I.tracks, &P.spec.user));
    return cbs;
}

}  // namespace

// Hessian of the Lagrangian sigma*f~ + sum_r lambda_r g~_r (scaled space), lower triangle, entries in the
// order of make_hess_structure. Every node's contribution is evaluated in second-order forward mode over
// the directions (x_k, u_k, t0, tf): the time t = h tau_k + m and the factor h = (tf - t0)/2 are themselves
// dual numbers, so no chain rule is written out by hand here.
void Problem::eval_hess(const Instance& I, const double* zs, double sigma, const double* lambda, double* vals) const {
    check_instance(*this, I);
    auto cbs = build_callbacks<Dual2>(*this, I);
    std::vector<double> z;
    unscale(*this, zs, &z);
    const int ns = L.ns, nc = L.nc, nd = ns + nc + 2;
    if (nd > MAXD2) throw std::invalid_argument("too many node variables for Dual2");
    size_t e = 0;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], np = L.npath[p];
        const Collocation& C = col[p];
        // node Lagrangians
        std::vector<Dual2> ell(N);
        for (int k = 0; k < N; ++k) {
            std::vector<Dual2> x(ns), u(nc), F(ns), path(np > 0 ? np : 1);
            for (int i = 0; i < ns; ++i) {
                x[i] = Dual2(z[L.ix(p, k, i)]);
                x[i].d[i] = 1.0;
            }
            for (int j = 0; j < nc; ++j) {
                u[j] = Dual2(z[L.iu(p, k, j)]);
                u[j].d[ns + j] = 1.0;
            }
            Dual2 t0(z[L.it0(p)]), tf(z[L.itf(p)]);
            t0.d[ns + nc] = 1.0;
            tf.d[ns + nc + 1] = 1.0;
            Dual2 h = Dual2(0.5) * (tf - t0), mid = Dual2(0.5) * (tf + t0);
            Dual2 t = h * Dual2(C.tau[k]) + mid;
            node_dae<Dual2>(cbs[p], ns, nc, F.data(), path.data(), x.data(), u.data(), t, 0.0);
            Dual2 Lk = node_cost<Dual2>(cbs[p], ns, nc, x.data(), u.data(), t, 0.0, spec.maximize);
            Dual2 acc = Dual2(sigma * sc.sf * C.w[k]) * (h * Lk);
            for (int i = 0; i < ns; ++i)
                acc = acc - Dual2(lambda[L.rdef(p, k, i)] * sc.sg[L.rdef(p, k, i)]) * (h * F[i]);
            for (int q = 0; q < np; ++q) acc = acc + Dual2(lambda[L.rpath(p, k, q)] * sc.sg[L.rpath(p, k, q)]) * path[q];
            ell[k] = acc;
        }
        auto isz = [&](int col) { return sc.isz[col]; };
        const int it0 = L.it0(p), itf = L.itf(p), T0 = ns + nc, TF = ns + nc + 1;
        for (int k = 0; k < N; ++k)
            for (int j = 0; j < nc; ++j) {
                const int c = L.iu(p, k, j), dj = ns + j;
                for (int j2 = j; j2 < nc; ++j2) vals[e++] = ell[k].h[dj][ns + j2] * isz(c) * isz(L.iu(p, k, j2));
                for (int i = 0; i < ns; ++i) vals[e++] = ell[k].h[dj][i] * isz(c) * isz(L.ix(p, k, i));
                vals[e++] = ell[k].h[dj][T0] * isz(c) * isz(it0);
                vals[e++] = ell[k].h[dj][TF] * isz(c) * isz(itf);
            }
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < ns; ++i) {
                const int c = L.ix(p, k, i);
                for (int i2 = i; i2 < ns; ++i2) vals[e++] = ell[k].h[i][i2] * isz(c) * isz(L.ix(p, k, i2));
                vals[e++] = ell[k].h[i][T0] * isz(c) * isz(it0);
                vals[e++] = ell[k].h[i][TF] * isz(c) * isz(itf);
            }
        double s00 = 0.0, s01 = 0.0, s11 = 0.0;
        for (int k = 0; k < N; ++k) {
            s00 += ell[k].h[T0][T0];
            s01 += ell[k].h[T0][TF];
            s11 += ell[k].h[TF][TF];
        }
        vals[e++] = s00 * isz(it0) * isz(it0);
        vals[e++] = s01 * isz(it0) * isz(itf);
        vals[e++] = s11 * isz(itf) * isz(itf);
    }
}

// pattern of eval_hess: per phase, column by column, rows ascending (lower triangle)
void Problem::hess_structure(std::vector<int32_t>* irow, std::vector<int32_t>* jcol) const {
    irow->clear();
    jcol->clear();
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p];
        for (int k = 0; k < N; ++k)
            for (int j = 0; j < L.nc; ++j) {
                for (int j2 = j; j2 < L.nc; ++j2) irow->push_back(L.iu(p, k, j2)), jcol->push_back(L.iu(p, k, j));
                for (int i = 0; i < L.ns; ++i) irow->push_back(L.ix(p, k, i)), jcol->push_back(L.iu(p, k, j));
                irow->push_back(L.it0(p)), jcol->push_back(L.iu(p, k, j));
                irow->push_back(L.itf(p)), jcol->push_back(L.iu(p, k, j));
            }
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < L.ns; ++i) {
                for (int i2 = i; i2 < L.ns; ++i2) irow->push_back(L.ix(p, k, i2)), jcol->push_back(L.ix(p, k, i));
                irow->push_back(L.it0(p)), jcol->push_back(L.ix(p, k, i));
                irow->push_back(L.itf(p)), jcol->push_back(L.ix(p, k, i));
            }
        irow->push_back(L.it0(p)), jcol->push_back(L.it0(p));
        irow->push_back(L.itf(p)), jcol->push_back(L.it0(p));
        irow->push_back(L.itf(p)), jcol->push_back(L.itf(p));
    }
}

void Problem::eval_g(const Instance& I, const double* zs, double* g) const {
    check_instance(*this, I);
    auto cbs = build_callbacks<double>(*this, I);
    g_with_callbacks(*this, cbs, zs, g);
}

void Problem::eval_f(const Instance& I, const double* zs, double* f) const {
    check_instance(*this, I);
    auto cbs = build_callbacks<double>(*this, I);
    std::vector<double> z;
    unscale(*this, zs, &z);
    double total = 0.0;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p];
        PhaseTimes pt = phase_times(L, p, z);
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
            double t = pt.h * col[p].tau[k] + pt.m;
            double Lk = node_cost<double>(cbs[p], L.ns, L.nc, &z[L.ix(p, k, 0)], &z[L.iu(p, k, 0)], t, 0.0,
                                          spec.maximize);
            acc = std::fma(col[p].w[k], Lk, acc);
        }
        double fp = pt.h * acc;  // endpoint cost is 0 (ePSOPT.cpp:302-306)
        total = (p == 0) ? fp : total + fp;
    }
    *f = sc.sf * total;
}

void Problem::eval_grad_f(const Instance& I, const double* zs, double* grad) const {
    check_instance(*this, I);
    auto cbs = build_callbacks<Dual>(*this, I);
    std::vector<double> z;
    unscale(*this, zs, &z);
    const int ns = L.ns, nc = L.nc, nd = ns + nc + 1;
    for (int c = 0; c < L.nvars; ++c) grad[c] = 0.0;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p];
        PhaseTimes pt = phase_times(L, p, z);
        double acc = 0.0, dt0 = 0.0, dtf = 0.0;
        for (int k = 0; k < N; ++k) {
            double tau = col[p].tau[k], w = col[p].w[k];
            std::vector<Dual> xs(ns), us(nc);
            for (int i = 0; i < ns; ++i) {
                xs[i] = Dual(z[L.ix(p, k, i)]);
                xs[i].d[i] = 1.0;
            }
            for (int j = 0; j < nc; ++j) {
                us[j] = Dual(z[L.iu(p, k, j)]);
                us[j].d[ns + j] = 1.0;
            }
            Dual t(pt.h * tau + pt.m);
            t.d[nd - 1] = 1.0;
            Dual Lk = node_cost<Dual>(cbs[p], ns, nc, xs.data(), us.data(), t, 0.0, spec.maximize);
            acc = std::fma(w, Lk.v, acc);
            for (int i = 0; i < ns; ++i) grad[L.ix(p, k, i)] = pt.h * (w * Lk.d[i]);
            for (int j = 0; j < nc; ++j) grad[L.iu(p, k, j)] = pt.h * (w * Lk.d[ns + j]);
            dt0 += w * (Lk.d[nd - 1] * (0.5 * (1.0 - tau)));
            dtf += w * (Lk.d[nd - 1] * (0.5 * (1.0 + tau)));
        }
        grad[L.it0(p)] = -0.5 * acc + pt.h * dt0;
        grad[L.itf(p)] = 0.5 * acc + pt.h * dtf;
    }
    for (int c = 0; c < L.nvars; ++c) grad[c] = (sc.sf * grad[c]) * sc.isz[c];
}

void Problem::eval_jac_exact(const Instance& I, const double* zs, double* vals) const {
    check_instance(*this, I);
    auto cbs = build_callbacks<Dual>(*this, I);
    std::vector<double> z;
    unscale(*this, zs, &z);
    const int ns = L.ns, nc = L.nc, nd = ns + nc + 1;
    // per phase: F[k][i], P[k][q] as duals w.r.t. (x_k, u_k, t_k)
    std::vector<std::vector<Dual>> F(L.nphases), PA(L.nphases);
    std::vector<PhaseTimes> pts;
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], np = L.npath[p];
        PhaseTimes pt = phase_times(L, p, z);
        pts.push_back(pt);
        F[p].resize(static_cast<size_t>(N) * ns);
        PA[p].resize(static_cast<size_t>(N) * (np > 0 ? np : 1));
        for (int k = 0; k < N; ++k) {
            std::vector<Dual> xs(ns), us(nc);
            for (int i = 0; i < ns; ++i) {
                xs[i] = Dual(z[L.ix(p, k, i)]);
                xs[i].d[i] = 1.0;
            }
            for (int j = 0; j < nc; ++j) {
                us[j] = Dual(z[L.iu(p, k, j)]);
                us[j].d[ns + j] = 1.0;
            }
            Dual t(pt.h * col[p].tau[k] + pt.m);
            t.d[nd - 1] = 1.0;
            node_dae<Dual>(cbs[p], ns, nc, &F[p][static_cast<size_t>(k) * ns],
                           &PA[p][static_cast<size_t>(k) * (np > 0 ? np : 1)], xs.data(), us.data(), t, 0.0);
        }
    }
    // row / column decoding
    struct RowId {
        int kind, p, k, a;  // kind: 0 defect,1 event,2 path,3 last,4 link
    };
    struct ColId {
        int kind, p, k, j;  // kind: 0 control,1 state,2 t0,3 tf
    };
    auto decode_row = [&](int r) {
        RowId id{-1, 0, 0, 0};
        if (r >= L.linkoff) {
            id.kind = 4;
            id.p = (r - L.linkoff) / (ns + 1);
            id.a = (r - L.linkoff) % (ns + 1);
            return id;
        }
        int p = 0;
        while (p + 1 < L.nphases && r >= L.goff[p + 1]) ++p;
        id.p = p;
        int q = r - L.goff[p], N = L.N[p], np = L.npath[p];
        if (q < ns * N) {
            id.kind = 0;
            id.k = q / ns;
            id.a = q % ns;
        } else if (q < ns * N + L.ne) {
            id.kind = 1;
            id.a = q - ns * N;
        } else if (q < ns * N + L.ne + np * N) {
            id.kind = 2;
            id.k = (q - ns * N - L.ne) / np;
            id.a = (q - ns * N - L.ne) % np;
        } else {
            id.kind = 3;
        }
        return id;
    };
    auto decode_col = [&](int c) {
        ColId id{-1, 0, 0, 0};
        int p = 0;
        while (p + 1 < L.nphases && c >= L.zoff[p + 1]) ++p;
        id.p = p;
        int q = c - L.zoff[p], N = L.N[p];
        if (q < nc * N) {
            id.kind = 0;
            id.k = q / nc;
            id.j = q % nc;
        } else if (q < (nc + ns) * N) {
            id.kind = 1;
            id.k = (q - nc * N) / ns;
            id.j = (q - nc * N) % ns;
        } else {
            id.kind = (q == (nc + ns) * N) ? 2 : 3;
        }
        return id;
    };
    const size_t nnz = S.irow.size();
    for (size_t e = 0; e < nnz; ++e) {
        int r = S.irow[e], c = S.jcol[e];
        RowId R = decode_row(r);
        ColId Cc = decode_col(c);
        double v = 0.0;
        bool ok = true;
        if (R.kind == 4) {  // linkage: x_a(tf) - x_{a+1}(t0) ; tf_a - t0_{a+1}
            v = (Cc.p == R.p) ? 1.0 : -1.0;
        } else if (R.p != Cc.p) {
            ok = false;
        } else {
            const int p = R.p, N = L.N[p], np = L.npath[p];
            const PhaseTimes& pt = pts[p];
            const Collocation& C = col[p];
            if (R.kind == 0) {
                const Dual& f = F[p][static_cast<size_t>(R.k) * ns + R.a];
                double tau = C.tau[R.k];
                if (Cc.kind == 1) {
                    if (Cc.k != R.k) {
                        v = C.D[static_cast<size_t>(R.k) * N + Cc.k];
                        ok = (Cc.j == R.a);
                    } else {
                        double dterm = (Cc.j == R.a) ? C.D[static_cast<size_t>(R.k) * N + R.k] : 0.0;
                        v = dterm - pt.h * f.d[Cc.j];
                    }
                } else if (Cc.kind == 0) {
                    ok = (Cc.k == R.k);
                    v = -(pt.h * f.d[ns + Cc.j]);
                } else if (Cc.kind == 2) {
                    v = 0.5 * f.v - pt.h * (f.d[nd - 1] * (0.5 * (1.0 - tau)));
                } else {
                    v = -0.5 * f.v - pt.h * (f.d[nd - 1] * (0.5 * (1.0 + tau)));
                }
            } else if (R.kind == 1) {
                v = 1.0;
                ok = (Cc.kind == 1);
            } else if (R.kind == 2) {
                const Dual& pa = PA[p][static_cast<size_t>(R.k) * np + R.a];
                double tau = C.tau[R.k];
                if (Cc.kind == 1) {
                    ok = (Cc.k == R.k);
                    v = pa.d[Cc.j];
                } else if (Cc.kind == 0) {
                    ok = (Cc.k == R.k);
                    v = pa.d[ns + Cc.j];
                } else if (Cc.kind == 2) {
                    v = pa.d[nd - 1] * (0.5 * (1.0 - tau));
                } else {
                    v = pa.d[nd - 1] * (0.5 * (1.0 + tau));
                }
            } else {
                v = (Cc.kind == 2) ? -1.0 : 1.0;
                ok = (Cc.kind >= 2);
            }
        }
        if (!ok) throw std::logic_error("pattern entry with no derivative rule");
        vals[e] = (sc.sg[r] * v) * sc.isz[c];
    }
}

void Problem::eval_jac_fd(const Instance& I, const double* zs, double* vals) const {
    // PSOPT derivatives="numerical" (SURVEY.md Appendix A.7): every column of a CPR group is
    // perturbed at once, the full constraint vector is re-evaluated at z+delta and z-delta, and
    // entry (r,c) is read from row r of the difference.
    check_instance(*this, I);
    auto cbs = build_callbacks<double>(*this, I);
    const double sqrt_eps = 1.4901161193847656e-08;  // 2^-26
    std::vector<double> zp(L.nvars), zm(L.nvars), gp(L.ncons), gm(L.ncons), rinv(L.nvars);
    for (int grp = 0; grp < S.ngroups; ++grp) {
        for (int c = 0; c < L.nvars; ++c) {
            zp[c] = zs[c];
            zm[c] = zs[c];
        }
        for (int c = 0; c < L.nvars; ++c) {
            if (S.group_of_col[c] != grp) continue;
            double delta = sqrt_eps * (1.0 + std::fabs(zs[c]));
            zp[c] = zs[c] + delta;
            zm[c] = zs[c] - delta;
            rinv[c] = 1.0 / (2.0 * delta);
        }
        g_with_callbacks(*this, cbs, zp.data(), gp.data());
        g_with_callbacks(*this, cbs, zm.data(), gm.data());
        for (int c = 0; c < L.nvars; ++c) {
            if (S.group_of_col[c] != grp) continue;
            for (int e = S.colptr[c]; e < S.colptr[c + 1]; ++e) {
                int r = S.irow[e];
                vals[e] = (gp[r] - gm[r]) * rinv[c];
            }
        }
    }
}

}  // namespace oracle
