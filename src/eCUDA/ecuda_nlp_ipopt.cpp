// ecuda_nlp_ipopt.cpp -- IPOPT driver for eCUDA::solve(), compiled only when the build found IPOPT
// (CMake: find_package(IPOPT) -> -DECUDA_HAVE_IPOPT, like src/ePSOPT/CMakeLists.txt:5-19 requires it
// for ePSOPT). The reference reaches IPOPT through PSOPT (src/ePSOPT/ePSOPT.cpp:62,84); here the five
// TNLP callbacks are thin forwards to the GPU evaluator. Hessian: the exact Lagrangian Hessian from
// the device (ecuda_eval_hess) when the problem carries eval_h -- the reference asks PSOPT for
// hessian="exact" from ADOL-C, ePSOPT.cpp:65 -- otherwise IPOPT's limited-memory quasi-Newton.
//
// Not compiled in this repository's container (no IPOPT headers); kept warning-free against
// IpStdCInterface.h of IPOPT 3.11-3.14.
#ifdef ECUDA_HAVE_IPOPT
#include <IpStdCInterface.h>

#include <algorithm>
#include <cstring>
#include <limits>
#include <string>

#include "ecuda_nlp.hpp"

namespace ecuda_nlp {

namespace {
struct Ctx {
    const Problem* P;
};

Bool cb_f(Index, Number* x, Bool, Number* obj, UserDataPtr ud) {
    return static_cast<Ctx*>(ud)->P->eval(x, obj, nullptr, nullptr, nullptr) ? TRUE : FALSE;
}
Bool cb_grad(Index, Number* x, Bool, Number* grad, UserDataPtr ud) {
    return static_cast<Ctx*>(ud)->P->eval(x, nullptr, nullptr, nullptr, grad) ? TRUE : FALSE;
}
Bool cb_g(Index, Number* x, Bool, Index, Number* g, UserDataPtr ud) {
    return static_cast<Ctx*>(ud)->P->eval(x, nullptr, g, nullptr, nullptr) ? TRUE : FALSE;
}
Bool cb_jac(Index, Number* x, Bool, Index, Index nele, Index* iRow, Index* jCol, Number* values, UserDataPtr ud) {
    const Problem* P = static_cast<Ctx*>(ud)->P;
    if (!values) {  // structure query
        for (Index e = 0; e < nele; ++e) {
            iRow[e] = P->irow[e];
            jCol[e] = P->jcol[e];
        }
        return TRUE;
    }
    return P->eval(x, nullptr, nullptr, values, nullptr) ? TRUE : FALSE;
}
Bool cb_h(Index, Number* x, Bool, Number obj_factor, Index, Number* lambda, Bool, Index nele, Index* iRow, Index* jCol,
          Number* values, UserDataPtr ud) {
    const Problem* P = static_cast<Ctx*>(ud)->P;
    if (!P->eval_h) return FALSE;  // limited-memory Hessian approximation: never called
    if (!values) {                 // structure query: lower triangle
        for (Index e = 0; e < nele; ++e) {
            iRow[e] = P->hrow[e];
            jCol[e] = P->hcol[e];
        }
        return TRUE;
    }
    return P->eval_h(x, obj_factor, lambda, values) ? TRUE : FALSE;
}
}  // namespace

bool have_ipopt() { return true; }

int solve_ipopt(const Problem& P, const Options& opt, std::vector<double>* z, Result* out) {
    std::vector<Number> xl(P.zl), xu(P.zu), gl(P.gl), gu(P.gu);
    for (auto* v : {&xl, &gl})
        for (Number& b : *v)
            if (b == -std::numeric_limits<double>::infinity()) b = -2e19;
    for (auto* v : {&xu, &gu})
        for (Number& b : *v)
            if (b == std::numeric_limits<double>::infinity()) b = 2e19;
    const bool exact_h = static_cast<bool>(P.eval_h) && P.hnnz > 0;
    IpoptProblem nlp = CreateIpoptProblem(P.n, xl.data(), xu.data(), P.m, gl.data(), gu.data(), P.nnz, exact_h ? P.hnnz : 0,
                                          0, &cb_f, &cb_g, &cb_grad, &cb_jac, &cb_h);
    if (!nlp) {
        if (out) out->message = "CreateIpoptProblem failed";
        return 7;
    }
    AddIpoptStrOption(nlp, const_cast<char*>("hessian_approximation"),
                      const_cast<char*>(exact_h ? "exact" : "limited-memory"));
    AddIpoptNumOption(nlp, const_cast<char*>("tol"), opt.tol);
    AddIpoptIntOption(nlp, const_cast<char*>("max_iter"), opt.max_iter);
    AddIpoptIntOption(nlp, const_cast<char*>("print_level"), opt.print_level);
    Ctx ctx{&P};
    std::vector<Number> g(P.m), lam(P.m), zL(P.n), zU(P.n);
    Number obj = 0.0;
    const int status = IpoptSolve(nlp, z->data(), g.data(), &obj, lam.data(), zL.data(), zU.data(), &ctx);
    FreeIpoptProblem(nlp);
    if (out) {
        out->objective = obj;
        double viol = 0.0;
        for (int r = 0; r < P.m; ++r) viol = std::max(viol, std::max(P.gl[r] - g[r], g[r] - P.gu[r]));
        out->max_violation = std::max(viol, 0.0);
        out->message = "IPOPT return status " + std::to_string(status);
    }
    return (status == 0 || status == 1) ? 0 : 8;  // Solve_Succeeded / Solved_To_Acceptable_Level
}

}  // namespace ecuda_nlp
#endif  // ECUDA_HAVE_IPOPT
