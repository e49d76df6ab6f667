// stub_ipopt.cpp -- TEST STUB behind tests/ipopt_stub/IpStdCInterface.h, plus a small driver (stub_run) that hands
// src/eCUDA/ecuda_nlp_ipopt.cpp a toy NLP and reports which callbacks the adapter served. Not an optimiser.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "IpStdCInterface.h"
#include "../../src/eCUDA/ecuda_nlp.hpp"

struct IpoptProblemInfo {
    Index n, m, nele_jac, nele_hess;
    std::vector<Number> xl, xu, gl, gu;
    Eval_F_CB f;
    Eval_G_CB g;
    Eval_Grad_F_CB grad;
    Eval_Jac_G_CB jac;
    Eval_H_CB h;
    std::map<std::string, std::string> sopt;
    std::map<std::string, double> nopt;
};
static IpoptProblemInfo g_last;  // what the adapter configured, for the test to inspect
static std::vector<double> g_seen;  // values the "solver" received from the callbacks

extern "C" {
IpoptProblem CreateIpoptProblem(Index n, Number* x_L, Number* x_U, Index m, Number* g_L, Number* g_U, Index nele_jac,
                                Index nele_hess, Index index_style, Eval_F_CB eval_f, Eval_G_CB eval_g,
                                Eval_Grad_F_CB eval_grad_f, Eval_Jac_G_CB eval_jac_g, Eval_H_CB eval_h) {
    if (index_style != 0) return nullptr;
    IpoptProblemInfo* p = new IpoptProblemInfo;
    p->n = n; p->m = m; p->nele_jac = nele_jac; p->nele_hess = nele_hess;
    p->xl.assign(x_L, x_L + n); p->xu.assign(x_U, x_U + n); p->gl.assign(g_L, g_L + m); p->gu.assign(g_U, g_U + m);
    p->f = eval_f; p->g = eval_g; p->grad = eval_grad_f; p->jac = eval_jac_g; p->h = eval_h;
    return p;
}
void FreeIpoptProblem(IpoptProblem p) { g_last = *p; delete p; }
Bool AddIpoptStrOption(IpoptProblem p, char* k, char* v) { p->sopt[k] = v; return TRUE; }
Bool AddIpoptNumOption(IpoptProblem p, char* k, Number v) { p->nopt[k] = v; return TRUE; }
Bool AddIpoptIntOption(IpoptProblem p, char* k, Int v) { p->nopt[k] = v; return TRUE; }
int IpoptSolve(IpoptProblem p, Number* x, Number* g, Number* obj_val, Number* mult_g, Number*, Number*, UserDataPtr ud) {
    g_seen.clear();
    std::vector<Index> ir(p->nele_jac), jc(p->nele_jac), hr(p->nele_hess), hc(p->nele_hess);
    std::vector<Number> jv(p->nele_jac), hv(p->nele_hess), grad(p->n), lam(p->m, 0.5);
    if (!p->jac(p->n, x, TRUE, p->m, p->nele_jac, ir.data(), jc.data(), nullptr, ud)) return -1;   // structure query
    if (!p->f(p->n, x, TRUE, obj_val, ud) || !p->grad(p->n, x, FALSE, grad.data(), ud) || !p->g(p->n, x, FALSE, p->m, g, ud) ||
        !p->jac(p->n, x, FALSE, p->m, p->nele_jac, nullptr, nullptr, jv.data(), ud))
        return -2;
    if (p->nele_hess > 0) {
        if (!p->h(p->n, x, FALSE, 1.0, p->m, lam.data(), TRUE, p->nele_hess, hr.data(), hc.data(), nullptr, ud)) return -3;
        if (!p->h(p->n, x, FALSE, 1.0, p->m, lam.data(), TRUE, p->nele_hess, nullptr, nullptr, hv.data(), ud)) return -4;
    }
    for (Index r = 0; r < p->m; ++r) mult_g[r] = lam[r];
    g_seen.push_back(*obj_val);
    for (double v : grad) g_seen.push_back(v);
    for (Index e = 0; e < p->nele_jac; ++e) { g_seen.push_back(ir[e]); g_seen.push_back(jc[e]); g_seen.push_back(jv[e]); }
    for (Index e = 0; e < p->nele_hess; ++e) { g_seen.push_back(hr[e]); g_seen.push_back(hc[e]); g_seen.push_back(hv[e]); }
    return 0;  // Solve_Succeeded
}

// toy NLP: min x0^2 + x1^2  s.t.  x0 * x1 >= 1 (one row, two Jacobian entries), optional exact Hessian.
// out: [rc, objective, max violation, n values the stub solver saw..., hessian option flag (1 exact / 0 limited-memory)]
int stub_run(int with_hessian, double* out, int cap) {
    using namespace ecuda_nlp;
    static const int32_t irow[2] = {0, 0}, jcol[2] = {0, 1}, hrow[3] = {0, 1, 1}, hcol[3] = {0, 0, 1};
    Problem P;
    P.n = 2; P.m = 1; P.nnz = 2;
    P.zl = {-10.0, -10.0}; P.zu = {10.0, INFINITY};
    P.gl = {1.0}; P.gu = {INFINITY};
    P.irow = irow; P.jcol = jcol;
    P.eval = [](const double* z, double* f, double* g, double* jac, double* grad) {
        if (f) *f = z[0] * z[0] + z[1] * z[1];
        if (g) g[0] = z[0] * z[1];
        if (jac) { jac[0] = z[1]; jac[1] = z[0]; }
        if (grad) { grad[0] = 2 * z[0]; grad[1] = 2 * z[1]; }
        return true;
    };
    if (with_hessian) {
        P.hnnz = 3; P.hrow = hrow; P.hcol = hcol;
        P.eval_h = [](const double*, double sigma, const double* lam, double* h) {
            h[0] = 2 * sigma; h[1] = lam[0]; h[2] = 2 * sigma;
            return true;
        };
    }
    std::vector<double> z = {2.0, 3.0};
    Result res;
    Options opt;
    opt.max_iter = 77;
    const int rc = solve_ipopt(P, opt, &z, &res);
    int n = 0;
    auto put = [&](double v) { if (n < cap) out[n++] = v; };
    put(rc); put(res.objective); put(res.max_violation); put(have_ipopt() ? 1 : 0);
    put(g_last.sopt["hessian_approximation"] == "exact" ? 1 : 0);
    put(g_last.nopt["max_iter"]); put(g_last.xu[1]);  // +inf must have become IPOPT's 2e19
    for (double v : g_seen) put(v);
    return n;
}
}
