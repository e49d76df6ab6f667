// ecuda_nlp.cpp -- host NLP drivers for eCUDA::solve() (see ecuda_nlp.hpp).
//
// solve_builtin: primal-dual interior point on the slack formulation
//     min f(z)  s.t.  c_E(z) = b_E,   c_I(z) - s = 0,   gl_I <= s <= gu_I,   zl <= z <= zu
// with a damped-BFGS model B of the Lagrangian Hessian over the free variables (started from
// diag(d2f), which one extra gradient evaluation gives exactly for the separable running costs of the
// device models; Powell damping keeps B positive definite, so the nonconvex obstacle rows are
// handled like in an SQP method). With H = B + barrier terms positive definite the Newton system
// reduces to an m x m symmetric positive definite Schur complement
//     (J H^-1 J' + E S^-1 E' + dc I) dlambda = rhs
// built from the Jacobian triplets the GPU returns and factorised by dense Cholesky on the host --
// the KKT solve stays on the host, as it does with IPOPT in the reference
// (src/ePSOPT/ePSOPT.cpp:62). An l1 merit function with backtracking and the fraction-to-boundary
// rule globalises it; the barrier parameter follows the monotone Fiacco-McCormick schedule.
#include "ecuda_nlp.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>

namespace ecuda_nlp {

namespace {

const double INF = std::numeric_limits<double>::infinity();

// in-place dense Cholesky A = L L' (lower triangle), returns false if a pivot is not positive
bool cholesky(std::vector<double>& A, int n) {
    for (int j = 0; j < n; ++j) {
        double* Aj = &A[static_cast<size_t>(j) * n];
        double d = Aj[j];
        for (int k = 0; k < j; ++k) d -= Aj[k] * Aj[k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        Aj[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double* Ai = &A[static_cast<size_t>(i) * n];
            double s = Ai[j];
            for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];
            Ai[j] = s / d;
        }
    }
    return true;
}
void chol_solve(const std::vector<double>& L, int n, std::vector<double>& b) {
    for (int i = 0; i < n; ++i) {
        const double* Li = &L[static_cast<size_t>(i) * n];
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= Li[k] * b[k];
        b[i] = s / Li[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[static_cast<size_t>(k) * n + i] * b[k];
        b[i] = s / L[static_cast<size_t>(i) * n + i];
    }
}

struct Work {
    std::vector<double> g, jac, grad, grad2;
    double f = 0.0;
};

}  // namespace

int solve_builtin(const Problem& P, const Options& opt, std::vector<double>* zio, Result* out) {
    const int n = P.n, m = P.m, nnz = P.nnz;
    std::vector<double>& z = *zio;
    Result R;
    auto finish = [&](int rc, const std::string& msg) {
        R.message = msg;
        if (out) *out = R;
        return rc;
    };
    if (static_cast<long long>(m) * m > 64ll * 1000 * 1000)
        return finish(2, "builtin NLP driver: problem too large for the dense Schur complement (link IPOPT)");

    // classification of variables and rows
    std::vector<char> fixed(n, 0), ineq(m, 0);
    for (int c = 0; c < n; ++c) {
        fixed[c] = P.zl[c] == P.zu[c];
        if (fixed[c]) z[c] = P.zl[c];
    }
    for (int r = 0; r < m; ++r) ineq[r] = P.gl[r] != P.gu[r];
    // column pointers of the (col,row)-sorted triplets
    std::vector<int> colptr(n + 1, 0);
    for (int e = 0; e < nnz; ++e) ++colptr[P.jcol[e] + 1];
    for (int c = 0; c < n; ++c) colptr[c + 1] += colptr[c];

    // strictly interior start
    auto push_inside = [](double v, double lo, double hi) {
        if (lo == -INF && hi == INF) return v;
        const double k1 = 1e-2;
        double pl = lo == -INF ? 0.0 : std::min(k1 * std::max(1.0, std::fabs(lo)), hi == INF ? INF : k1 * (hi - lo));
        double pu = hi == INF ? 0.0 : std::min(k1 * std::max(1.0, std::fabs(hi)), lo == -INF ? INF : k1 * (hi - lo));
        if (lo != -INF) v = std::max(v, lo + pl);
        if (hi != INF) v = std::min(v, hi - pu);
        return v;
    };
    for (int c = 0; c < n; ++c)
        if (!fixed[c]) z[c] = push_inside(z[c], P.zl[c], P.zu[c]);

    Work W;
    W.g.resize(m);
    W.jac.resize(nnz);
    W.grad.resize(n);
    W.grad2.resize(n);
    std::vector<double> s(m, 0.0), lam(m, 0.0), hdiag(n, 0.0);
    if (!P.eval(z.data(), &W.f, W.g.data(), W.jac.data(), W.grad.data()))
        return finish(3, "builtin NLP driver: evaluation failed at the initial point");
    for (int r = 0; r < m; ++r)
        if (ineq[r]) s[r] = push_inside(W.g[r], P.gl[r], P.gu[r]);

    // diag(d2f) from one extra gradient (exact for a separable objective); refreshed every 10 iterations
    auto refresh_hessian = [&]() -> bool {
        const double eps = 1e-6;
        std::vector<double> zp(z);
        for (int c = 0; c < n; ++c)
            if (!fixed[c]) zp[c] += eps;
        if (!P.eval(zp.data(), nullptr, nullptr, nullptr, W.grad2.data())) return false;
        for (int c = 0; c < n; ++c) hdiag[c] = fixed[c] ? 0.0 : std::max(0.0, (W.grad2[c] - W.grad[c]) / eps);
        return true;
    };
    if (!refresh_hessian()) return finish(3, "builtin NLP driver: gradient evaluation failed");

    // free variables and the dense quasi-Newton matrix over them
    std::vector<int> fidx(n, -1), fvars;
    for (int c = 0; c < n; ++c)
        if (!fixed[c]) {
            fidx[c] = static_cast<int>(fvars.size());
            fvars.push_back(c);
        }
    const int nf = static_cast<int>(fvars.size());
    std::vector<double> Bq(static_cast<size_t>(nf) * nf, 0.0), Hq(static_cast<size_t>(nf) * nf), Jf(static_cast<size_t>(m) * nf),
        Tq(static_cast<size_t>(nf) * m), uq(nf), grad_old(n), jac_old(nnz), sk(nf), yk(nf), Bs(nf);
    for (int i = 0; i < nf; ++i) Bq[static_cast<size_t>(i) * nf + i] = std::max(hdiag[fvars[i]], 1e-2);

    double mu = 0.1;
    const double tau_min = 0.99, kappa_eps = 10.0, mu_min = opt.tol / 10.0, kappa_sigma = 1e10;
    double delta = 1e-4;   // primal (proximal) regularisation of the diagonal Hessian model
    std::vector<double> S(static_cast<size_t>(m) * m), rhs(m), dz(n), ds(m), dl(m), hz(n), hs(m), rz(n), rs(m), rc(m);
    std::vector<double> ztrial(n), strial(m), gtrial(m), bzv(n, 0.0), bsv(m, 0.0);
    // bound multipliers (primal-dual): vL/vU for z, wL/wU for the slacks
    std::vector<double> vL(n, 0.0), vU(n, 0.0), wL(m, 0.0), wU(m, 0.0), dvL(n), dvU(n), dwL(m), dwU(m);
    for (int c = 0; c < n; ++c) {
        if (fixed[c]) continue;
        if (P.zl[c] != -INF) vL[c] = mu / (z[c] - P.zl[c]);
        if (P.zu[c] != INF) vU[c] = mu / (P.zu[c] - z[c]);
    }
    for (int r = 0; r < m; ++r) {
        if (!ineq[r]) continue;
        if (P.gl[r] != -INF) wL[r] = mu / (s[r] - P.gl[r]);
        if (P.gu[r] != INF) wU[r] = mu / (P.gu[r] - s[r]);
    }

    auto barrier_terms = [&](const std::vector<double>& zz, const std::vector<double>& ss, double muv) {
        double phi = 0.0;
        for (int c = 0; c < n; ++c) {
            if (fixed[c]) continue;
            if (P.zl[c] != -INF) phi -= muv * std::log(zz[c] - P.zl[c]);
            if (P.zu[c] != INF) phi -= muv * std::log(P.zu[c] - zz[c]);
        }
        for (int r = 0; r < m; ++r) {
            if (!ineq[r]) continue;
            if (P.gl[r] != -INF) phi -= muv * std::log(ss[r] - P.gl[r]);
            if (P.gu[r] != INF) phi -= muv * std::log(P.gu[r] - ss[r]);
        }
        return phi;
    };
    auto infeasibility = [&](const std::vector<double>& gg, const std::vector<double>& ss) {
        double v = 0.0;
        for (int r = 0; r < m; ++r) v += std::fabs(ineq[r] ? gg[r] - ss[r] : gg[r] - P.gl[r]);
        return v;
    };

    const double theta_init = infeasibility(W.g, s);
    int it = 0, last_progress = 0;
    for (; it < opt.max_iter; ++it) {
        // a quasi-Newton matrix that has grown ill-conditioned shows up as steps that shrink while the
        // dual infeasibility does not: restart it from the diagonal when a barrier problem takes long
        if (it - last_progress >= 25) {
            std::fill(Bq.begin(), Bq.end(), 0.0);
            for (int i = 0; i < nf; ++i) Bq[static_cast<size_t>(i) * nf + i] = std::max(hdiag[fvars[i]], 1e-2);
            last_progress = it;
        }
        // ---- optimality error of the barrier problem (primal-dual form) and Newton data
        double err_dual = 0.0, err_prim = 0.0, err_comp = 0.0;
        int worst_dual = 0;  // variable index, or -1 - row for a slack
        for (int c = 0; c < n; ++c) {
            double v = W.grad[c];
            for (int e = colptr[c]; e < colptr[c + 1]; ++e) v += W.jac[e] * lam[P.irow[e]];
            double bz = 0.0, sig = 0.0;
            if (!fixed[c]) {
                if (P.zl[c] != -INF) {
                    const double dlo = z[c] - P.zl[c];
                    bz -= mu / dlo;
                    sig += vL[c] / dlo;
                    err_comp = std::max(err_comp, std::fabs(vL[c] * dlo - mu));
                }
                if (P.zu[c] != INF) {
                    const double dhi = P.zu[c] - z[c];
                    bz += mu / dhi;
                    sig += vU[c] / dhi;
                    err_comp = std::max(err_comp, std::fabs(vU[c] * dhi - mu));
                }
                if (std::fabs(v - vL[c] + vU[c]) > err_dual) {
                    err_dual = std::fabs(v - vL[c] + vU[c]);
                    worst_dual = c;
                }
            }
            rz[c] = fixed[c] ? 0.0 : v + bz;  // barrier form of the dual residual (multipliers eliminated)
            bzv[c] = bz;
            hz[c] = hdiag[c] + sig + delta;
        }
        for (int r = 0; r < m; ++r) {
            if (ineq[r]) {
                double bs = 0.0, sig = 0.0;
                if (P.gl[r] != -INF) {
                    const double dlo = s[r] - P.gl[r];
                    bs -= mu / dlo;
                    sig += wL[r] / dlo;
                    err_comp = std::max(err_comp, std::fabs(wL[r] * dlo - mu));
                }
                if (P.gu[r] != INF) {
                    const double dhi = P.gu[r] - s[r];
                    bs += mu / dhi;
                    sig += wU[r] / dhi;
                    err_comp = std::max(err_comp, std::fabs(wU[r] * dhi - mu));
                }
                if (std::fabs(-lam[r] - wL[r] + wU[r]) > err_dual) {
                    err_dual = std::fabs(-lam[r] - wL[r] + wU[r]);
                    worst_dual = -1 - r;
                }
                rs[r] = -lam[r] + bs;
                bsv[r] = bs;
                hs[r] = sig + 1e-12;
                rc[r] = W.g[r] - s[r];
            } else {
                rs[r] = 0.0;
                hs[r] = 0.0;
                rc[r] = W.g[r] - P.gl[r];
            }
            err_prim = std::max(err_prim, std::fabs(rc[r]));
        }
        const double err = std::max(std::max(err_dual, err_prim), err_comp);
        if (opt.print_level > 0)
            std::printf("iter %3d  f %.8e  inf_pr %.2e  inf_du %.2e  compl %.2e  mu %.1e  delta %.1e\n", it, W.f,
                        err_prim, err_dual, err_comp, mu, delta);
        if (opt.print_level > 2) {
            if (worst_dual >= 0)
                std::printf("          worst dual: var %d z %.6e [%.3e, %.3e] vL %.3e vU %.3e\n", worst_dual, z[worst_dual],
                            P.zl[worst_dual], P.zu[worst_dual], vL[worst_dual], vU[worst_dual]);
            else {
                const int r = -1 - worst_dual;
                std::printf("          worst dual: slack of row %d s %.6e [%.3e, %.3e] lam %.3e wL %.3e wU %.3e g %.6e\n", r, s[r],
                            P.gl[r], P.gu[r], lam[r], wL[r], wU[r], W.g[r]);
            }
        }
        if (err <= opt.tol && mu <= mu_min * 1.0001) break;
        if (err <= kappa_eps * mu && mu > mu_min) {
            mu = std::max(mu_min, std::min(0.2 * mu, std::pow(mu, 1.5)));
            last_progress = it;
            continue;  // re-evaluate the residuals with the new barrier parameter
        }

        // ---- H = B + barrier diagonal (+ proximal delta), Cholesky over the free variables
        for (int i = 0; i < nf; ++i) {
            for (int j2 = 0; j2 <= i; ++j2) Hq[static_cast<size_t>(i) * nf + j2] = Bq[static_cast<size_t>(i) * nf + j2];
            Hq[static_cast<size_t>(i) * nf + i] += hz[fvars[i]] - hdiag[fvars[i]];  // sigma + delta
        }
        if (!cholesky(Hq, nf)) {
            delta = std::max(1e-4, delta * 100.0);
            if (delta > 1e8) return finish(4, "builtin NLP driver: Hessian model not positive definite");
            continue;
        }
        // dense Jacobian over the free columns, T = L^-1 Jf'  (nf x m, column r = constraint row r)
        std::fill(Jf.begin(), Jf.end(), 0.0);
        for (int c = 0; c < n; ++c) {
            if (fixed[c]) continue;
            for (int e = colptr[c]; e < colptr[c + 1]; ++e) Jf[static_cast<size_t>(P.irow[e]) * nf + fidx[c]] = W.jac[e];
        }
        for (int r = 0; r < m; ++r) {
            const double* jr = &Jf[static_cast<size_t>(r) * nf];
            for (int i = 0; i < nf; ++i) {  // forward substitution with L
                const double* Li = &Hq[static_cast<size_t>(i) * nf];
                double v = jr[i];
                for (int k2 = 0; k2 < i; ++k2) v -= Li[k2] * Tq[static_cast<size_t>(k2) * m + r];
                Tq[static_cast<size_t>(i) * m + r] = v / Li[i];
            }
        }
        // ---- Schur complement S = T'T + E Hs^-1 E' + dc I   (lower triangle)
        std::fill(S.begin(), S.end(), 0.0);
        for (int i = 0; i < nf; ++i) {
            const double* Ti = &Tq[static_cast<size_t>(i) * m];
            for (int r1 = 0; r1 < m; ++r1) {
                const double a = Ti[r1];
                if (a == 0.0) continue;
                double* Sr = &S[static_cast<size_t>(r1) * m];
                for (int r2 = 0; r2 <= r1; ++r2) Sr[r2] += a * Ti[r2];
            }
        }
        for (int r = 0; r < m; ++r) S[static_cast<size_t>(r) * m + r] += (ineq[r] ? 1.0 / hs[r] : 0.0) + 1e-9;
        // rhs = rc - Jf H^-1 rz + E Hs^-1 rs
        for (int i = 0; i < nf; ++i) uq[i] = rz[fvars[i]];
        chol_solve(Hq, nf, uq);
        for (int r = 0; r < m; ++r) {
            const double* jr = &Jf[static_cast<size_t>(r) * nf];
            double v = rc[r] + (ineq[r] ? rs[r] / hs[r] : 0.0);
            for (int i = 0; i < nf; ++i) v -= jr[i] * uq[i];
            rhs[r] = v;
        }
        if (!cholesky(S, m)) {
            delta = std::max(1e-4, delta * 100.0);
            if (delta > 1e8) return finish(4, "builtin NLP driver: Schur complement not positive definite");
            continue;
        }
        dl = rhs;
        chol_solve(S, m, dl);
        // dz = -H^-1 (rz + Jf' dlambda)
        for (int i = 0; i < nf; ++i) uq[i] = rz[fvars[i]];
        for (int r = 0; r < m; ++r) {
            const double* jr = &Jf[static_cast<size_t>(r) * nf];
            const double d = dl[r];
            for (int i = 0; i < nf; ++i) uq[i] += jr[i] * d;
        }
        chol_solve(Hq, nf, uq);
        std::fill(dz.begin(), dz.end(), 0.0);
        for (int i = 0; i < nf; ++i) dz[fvars[i]] = -uq[i];
        for (int r = 0; r < m; ++r) ds[r] = ineq[r] ? -(rs[r] - dl[r]) / hs[r] : 0.0;
        // bound-multiplier steps from the linearised complementarity conditions
        for (int c = 0; c < n; ++c) {
            dvL[c] = dvU[c] = 0.0;
            if (fixed[c]) continue;
            if (P.zl[c] != -INF) {
                const double dlo = z[c] - P.zl[c];
                dvL[c] = mu / dlo - vL[c] - vL[c] / dlo * dz[c];
            }
            if (P.zu[c] != INF) {
                const double dhi = P.zu[c] - z[c];
                dvU[c] = mu / dhi - vU[c] + vU[c] / dhi * dz[c];
            }
        }
        for (int r = 0; r < m; ++r) {
            dwL[r] = dwU[r] = 0.0;
            if (!ineq[r]) continue;
            if (P.gl[r] != -INF) {
                const double dlo = s[r] - P.gl[r];
                dwL[r] = mu / dlo - wL[r] - wL[r] / dlo * ds[r];
            }
            if (P.gu[r] != INF) {
                const double dhi = P.gu[r] - s[r];
                dwU[r] = mu / dhi - wU[r] + wU[r] / dhi * ds[r];
            }
        }

        // ---- fraction to the boundary (primal and dual)
        const double tau = std::max(tau_min, 1.0 - mu);
        double amax = 1.0, adual = 1.0;
        auto limit = [&](double v, double d, double lo, double hi) {
            if (d < 0.0 && lo != -INF) amax = std::min(amax, -tau * (v - lo) / d);
            if (d > 0.0 && hi != INF) amax = std::min(amax, tau * (hi - v) / d);
        };
        auto limit_dual = [&](double v, double d) {
            if (d < 0.0) adual = std::min(adual, -tau * v / d);
        };
        for (int c = 0; c < n; ++c)
            if (!fixed[c]) {
                limit(z[c], dz[c], P.zl[c], P.zu[c]);
                limit_dual(vL[c], dvL[c]);
                limit_dual(vU[c], dvU[c]);
            }
        for (int r = 0; r < m; ++r)
            if (ineq[r]) {
                limit(s[r], ds[r], P.gl[r], P.gu[r]);
                limit_dual(wL[r], dwL[r]);
                limit_dual(wU[r], dwU[r]);
            }

        // ---- backtracking line search with a filter-style acceptance test (no filter history):
        // away from feasibility a trial point must reduce the infeasibility theta or the barrier
        // objective phi sufficiently; once theta is negligible it must satisfy the Armijo condition on
        // phi while staying (nearly) feasible -- an l1 merit with a penalty tied to |lambda| rejects
        // good steps there (Maratos effect)
        const double theta0 = infeasibility(W.g, s);
        const double phi0 = W.f + barrier_terms(z, s, mu);
        double dphi = 0.0;  // directional derivative of the barrier objective
        for (int c = 0; c < n; ++c)
            if (!fixed[c]) dphi += (W.grad[c] + bzv[c]) * dz[c];
        for (int r = 0; r < m; ++r)
            if (ineq[r]) dphi += bsv[r] * ds[r];
        const double theta_small = 1e-7 * std::max(1.0, theta_init);
        const double dmerit = dphi;
        double alpha = amax;
        bool accepted = false;
        double ftrial = 0.0;
        for (int ls = 0; ls < 25; ++ls) {
            for (int c = 0; c < n; ++c) ztrial[c] = z[c] + alpha * dz[c];
            for (int r = 0; r < m; ++r) strial[r] = s[r] + alpha * ds[r];
            if (!P.eval(ztrial.data(), &ftrial, gtrial.data(), nullptr, nullptr))
                return finish(3, "builtin NLP driver: evaluation failed in the line search");
            const double phit = ftrial + barrier_terms(ztrial, strial, mu);
            const double thetat = infeasibility(gtrial, strial);
            if (std::isfinite(phit) && std::isfinite(thetat)) {
                if (theta0 > theta_small || dphi >= 0.0) {  // (no descent in phi: the step is a feasibility step)
                    if (thetat <= (1.0 - 1e-5) * theta0 || phit <= phi0 - 1e-5 * theta0) accepted = true;
                } else if (phit <= phi0 + 1e-4 * alpha * std::min(dphi, 0.0) + 1e-13 * std::fabs(phi0) &&
                           thetat <= 10.0 * theta_small) {
                    accepted = true;
                }
            }
            if (accepted) break;
            alpha *= 0.5;
        }
        if (!accepted) {
            // no progress along the Newton direction: regularise more and try again
            delta = std::max(1e-4, delta * 10.0);
            if (delta > 1e6) return finish(5, "builtin NLP driver: line search failed");
            continue;
        }
        if (opt.print_level > 1) {
            double ndz = 0.0, ndl = 0.0;
            for (int c = 0; c < n; ++c) ndz = std::max(ndz, std::fabs(dz[c]));
            for (int r = 0; r < m; ++r) ndl = std::max(ndl, std::fabs(dl[r]));
            std::printf("          step: amax %.3e  alpha %.3e  adual %.3e  |dz| %.3e  |dlam| %.3e  dmerit %.3e\n", amax, alpha,
                        adual, ndz, ndl, dmerit);
        }
        delta = std::max(1e-6, delta * 0.5);
        z = ztrial;
        s = strial;
        for (int r = 0; r < m; ++r) lam[r] += alpha * dl[r];
        // dual step + the usual safeguard keeping the multipliers near the central path
        auto upd = [&](double& v, double dv, double dist) {
            v += adual * dv;
            v = std::min(std::max(v, mu / (kappa_sigma * dist)), kappa_sigma * mu / dist);
        };
        for (int c = 0; c < n; ++c) {
            if (fixed[c]) continue;
            if (P.zl[c] != -INF) upd(vL[c], dvL[c], z[c] - P.zl[c]);
            if (P.zu[c] != INF) upd(vU[c], dvU[c], P.zu[c] - z[c]);
        }
        for (int r = 0; r < m; ++r) {
            if (!ineq[r]) continue;
            if (P.gl[r] != -INF) upd(wL[r], dwL[r], s[r] - P.gl[r]);
            if (P.gu[r] != INF) upd(wU[r], dwU[r], P.gu[r] - s[r]);
        }
        grad_old = W.grad;
        jac_old = W.jac;
        if (!P.eval(z.data(), &W.f, W.g.data(), W.jac.data(), W.grad.data()))
            return finish(3, "builtin NLP driver: evaluation failed");
        // damped BFGS update with s = step, y = change of the Lagrangian gradient at the new multipliers
        double sy = 0.0, sBs = 0.0;
        for (int i = 0; i < nf; ++i) {
            const int c = fvars[i];
            sk[i] = alpha * dz[c];
            double yn = W.grad[c], yo = grad_old[c];
            for (int e = colptr[c]; e < colptr[c + 1]; ++e) {
                yn += W.jac[e] * lam[P.irow[e]];
                yo += jac_old[e] * lam[P.irow[e]];
            }
            yk[i] = yn - yo;
        }
        for (int i = 0; i < nf; ++i) {
            double v = 0.0;
            for (int j2 = 0; j2 < nf; ++j2) {
                const double bij = j2 <= i ? Bq[static_cast<size_t>(i) * nf + j2] : Bq[static_cast<size_t>(j2) * nf + i];
                v += bij * sk[j2];
            }
            Bs[i] = v;
            sBs += sk[i] * v;
            sy += sk[i] * yk[i];
        }
        if (sBs > 1e-16) {
            double theta = 1.0;
            if (sy < 0.2 * sBs) theta = 0.8 * sBs / (sBs - sy);  // Powell damping
            double sr = 0.0;
            for (int i = 0; i < nf; ++i) {
                yk[i] = theta * yk[i] + (1.0 - theta) * Bs[i];
                sr += sk[i] * yk[i];
            }
            if (sr > 1e-16)
                for (int i = 0; i < nf; ++i)
                    for (int j2 = 0; j2 <= i; ++j2)
                        Bq[static_cast<size_t>(i) * nf + j2] += yk[i] * yk[j2] / sr - Bs[i] * Bs[j2] / sBs;
        }
    }

    R.iterations = it;
    R.objective = W.f;
    double viol = 0.0;
    for (int r = 0; r < m; ++r) viol = std::max(viol, std::max(P.gl[r] - W.g[r], W.g[r] - P.gu[r]));
    for (int c = 0; c < n; ++c) viol = std::max(viol, std::max(P.zl[c] - z[c], z[c] - P.zu[c]));
    R.max_violation = std::max(viol, 0.0);
    if (it >= opt.max_iter) {
        // IPOPT's "maximum iterations" is reported as a failure by PSOPT; accept the point when it is
        // feasible to a loose tolerance so that a near-converged trajectory is still returned
        if (R.max_violation <= 1e3 * opt.tol) return finish(0, "maximum number of iterations reached (feasible point returned)");
        return finish(1, "builtin NLP driver: maximum number of iterations exceeded");
    }
    return finish(0, "optimal solution found");
}

#ifndef ECUDA_HAVE_IPOPT
bool have_ipopt() { return false; }
int solve_ipopt(const Problem&, const Options&, std::vector<double>*, Result* out) {
    if (out) out->message = "this build of eCUDA was configured without IPOPT";
    return 6;
}
#endif

}  // namespace ecuda_nlp
