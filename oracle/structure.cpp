// structure.cpp (oracle) -- NLP dimensions, Jacobian sparsity pattern, CPR column grouping.
// TEST INFRASTRUCTURE (see oracle.hpp header).
//
// Dimensions follow ePSOPT::setup (src/ePSOPT/ePSOPT.cpp:41-45,58): nevents = 2*nstates,
// nodes = nsteps+1, npath = number of path parameters. Vector layouts, the structural pattern and
// the index-set grouping restate SURVEY.md Appendix A.2, A.3, A.6, A.7.
#include <algorithm>
#include <stdexcept>

#include "oracle.hpp"

namespace oracle {

namespace {

struct ModelShape {
    int ns, nc_default;
    // which states / controls each state derivative reads (bit masks), MODEL_DEPS mode
    std::vector<unsigned> fx, fu;
};

ModelShape model_shape(const Spec& s) {
    ModelShape m;
    const int model = s.model;
    switch (model) {
        case USER: {  // read sets by walking the tape backwards from every output
            const UserTape& ut = s.user;
            m.ns = ut.ns;
            m.nc_default = ut.nc;
            for (int i = 0; i < ut.ns; ++i) {
                std::vector<char> live(ut.nodes.size(), 0);
                live.at(ut.f_out.at(i)) = 1;
                unsigned fx = 0, fu = 0;
                for (int k = static_cast<int>(ut.nodes.size()) - 1; k >= 0; --k) {
                    if (!live[k]) continue;
                    const TapeNode& n = ut.nodes[k];
                    if (n.op == T_INPUT) {
                        if (n.a < ut.ns) fx |= 1u << n.a;
                        else if (n.a < ut.ns + ut.nc) fu |= 1u << (n.a - ut.ns);
                        continue;
                    }
                    if (n.op == T_CONST) continue;
                    live.at(n.a) = 1;
                    if (n.op >= T_ADD && n.op <= T_DIV) live.at(n.b) = 1;
                }
                m.fx.push_back(fx);
                m.fu.push_back(fu);
            }
            break;
        }
        case SI2D:  // xdot = u0, ydot = u1   (etol_psopt_example1.cpp:116-138)
            m.ns = 2;
            m.nc_default = 2;
            m.fx = {0u, 0u};
            m.fu = {1u << 0, 1u << 1};
            break;
        case PM3D:  // p' = v, v' = a
            m.ns = 6;
            m.nc_default = 3;
            m.fx = {1u << 3, 1u << 4, 1u << 5, 0u, 0u, 0u};
            m.fu = {0u, 0u, 0u, 1u << 0, 1u << 1, 1u << 2};
            break;
        case FW6: {  // x,y,z,V,gamma,psi ; controls aT, gamma_dot, psi_dot
            m.ns = 6;
            m.nc_default = 3;
            const unsigned V = 1u << 3, G = 1u << 4, P = 1u << 5;
            m.fx = {V | G | P, V | G | P, V | G, G, 0u, 0u};
            m.fu = {0u, 0u, 0u, 1u << 0, 1u << 1, 1u << 2};
            break;
        }
        default:
            throw std::invalid_argument("unknown model");
    }
    return m;
}

}  // namespace

Layout make_layout(const Spec& s) {
    ModelShape m = model_shape(s);
    Layout L;
    L.ns = m.ns;
    L.nc = s.ncontrols > 0 ? s.ncontrols : m.nc_default;
    if (L.nc < m.nc_default) throw std::invalid_argument("too few controls for model");
    if (s.model != SI2D && L.nc != m.nc_default) throw std::invalid_argument("ncontrols fixed for this model");
    if (s.model != SI2D && s.model != USER && s.ntracks != 0) throw std::invalid_argument("tracks are si2d-only");
    L.ne = 2 * L.ns;
    L.nphases = s.nphases;
    if (s.nphases < 1 || static_cast<int>(s.nnodes.size()) != s.nphases ||
        static_cast<int>(s.nstatic.size()) != s.nphases)
        throw std::invalid_argument("bad phase description");
    int zo = 0, go = 0;
    for (int p = 0; p < s.nphases; ++p) {
        int N = s.nnodes[p];
        if (N < 2) throw std::invalid_argument("need >= 2 nodes per phase");
        int np = s.nstatic[p] + s.ntracks + (s.model == USER ? static_cast<int>(s.user.row_out.size()) : 0);
        L.N.push_back(N);
        L.npath.push_back(np);
        L.zoff.push_back(zo);
        L.goff.push_back(go);
        int nv = (L.ns + L.nc) * N + 2;
        int ng = L.ns * N + L.ne + np * N + 1;
        L.nvars_p.push_back(nv);
        L.ncons_p.push_back(ng);
        zo += nv;
        go += ng;
    }
    L.nvars = zo;
    L.linkoff = go;
    L.nlink = (s.nphases - 1) * (L.ns + 1);
    L.ncons = go + L.nlink;
    return L;
}

Structure make_structure(const Spec& s, const Layout& L) {
    ModelShape m = model_shape(s);
    const bool dense = (s.pattern_mode == DENSE_NODE);
    const unsigned all_x = (1u << L.ns) - 1u, all_u = (1u << L.nc) - 1u;
    std::vector<std::vector<int32_t>> rows(L.nvars);
    for (int p = 0; p < L.nphases; ++p) {
        const int N = L.N[p], np = L.npath[p], nstat = s.nstatic[p];
        // path-row dependencies: static rows read the two position states; track rows add time
        auto path_reads_state = [&](int q, int j) { (void)q; return j == 0 || j == 1; };
        auto path_reads_time = [&](int q) { return q >= nstat; };
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < L.nc; ++j) {  // control columns
                auto& r = rows[L.iu(p, k, j)];
                for (int i = 0; i < L.ns; ++i) {
                    unsigned dep = dense ? all_u : m.fu[i];
                    if (dep >> j & 1u) r.push_back(L.rdef(p, k, i));
                }
            }
            for (int j = 0; j < L.ns; ++j) {  // state columns
                auto& r = rows[L.ix(p, k, j)];
                for (int kk = 0; kk < N; ++kk) {
                    if (kk != k) {
                        r.push_back(L.rdef(p, kk, j));  // through D
                    } else {
                        for (int i = 0; i < L.ns; ++i) {
                            unsigned dep = dense ? all_x : m.fx[i];
                            if (i == j || (dep >> j & 1u)) r.push_back(L.rdef(p, k, i));
                        }
                    }
                }
                if (k == 0) r.push_back(L.rev(p, j));              // e[i] = x(t0), ePSOPT.cpp:287-288
                if (k == N - 1) r.push_back(L.rev(p, L.ns + j));   // e[ns+i] = x(tf), ePSOPT.cpp:289-290
                for (int q = 0; q < np; ++q)
                    if (path_reads_state(q, j)) r.push_back(L.rpath(p, k, q));
                if (k == N - 1 && p + 1 < L.nphases) r.push_back(L.rlink(p, j));
                if (k == 0 && p > 0) r.push_back(L.rlink(p - 1, j));
            }
        }
        for (int which = 0; which < 2; ++which) {  // t0, tf columns
            auto& r = rows[which == 0 ? L.it0(p) : L.itf(p)];
            for (int k = 0; k < N; ++k)
                for (int i = 0; i < L.ns; ++i) r.push_back(L.rdef(p, k, i));
            for (int k = 0; k < N; ++k)
                for (int q = 0; q < np; ++q)
                    if (path_reads_time(q)) r.push_back(L.rpath(p, k, q));
            r.push_back(L.rlast(p));
            if (which == 0 && p > 0) r.push_back(L.rlink(p - 1, L.ns));
            if (which == 1 && p + 1 < L.nphases) r.push_back(L.rlink(p, L.ns));
        }
    }
    Structure S;
    S.colptr.assign(L.nvars + 1, 0);
    for (int c = 0; c < L.nvars; ++c) {
        std::sort(rows[c].begin(), rows[c].end());
        if (std::adjacent_find(rows[c].begin(), rows[c].end()) != rows[c].end())
            throw std::logic_error("duplicate pattern entry");
        S.colptr[c + 1] = S.colptr[c] + static_cast<int32_t>(rows[c].size());
        for (int32_t r : rows[c]) {
            S.irow.push_back(r);
            S.jcol.push_back(c);
        }
    }
    // Curtis-Powell-Reid: first-fit over columns in natural order
    S.group_of_col.assign(L.nvars, -1);
    std::vector<std::vector<char>> used;  // per group: rows already covered
    for (int c = 0; c < L.nvars; ++c) {
        int g = 0;
        for (;; ++g) {
            if (g == static_cast<int>(used.size())) used.emplace_back(L.ncons, 0);
            bool clash = false;
            for (int32_t r : rows[c])
                if (used[g][r]) {
                    clash = true;
                    break;
                }
            if (!clash) break;
        }
        for (int32_t r : rows[c]) used[g][r] = 1;
        S.group_of_col[c] = g;
    }
    S.ngroups = static_cast<int>(used.size());
    return S;
}

}  // namespace oracle
