// ecuda_nlp.hpp -- the host NLP drivers eCUDA::solve() can run over the GPU callbacks.
//
// In the reference the NLP is solved by IPOPT inside PSOPT (src/ePSOPT/ePSOPT.cpp:62,84); IPOPT's
// KKT factorisation stays on the host and is not part of the accelerated path. Two drivers:
//   solve_ipopt    IPOPT through its C interface (IpStdCInterface.h), compiled only when the build
//                  found IPOPT (ECUDA_HAVE_IPOPT); the callbacks are the ecuda_ipopt_* shims' twins
//   solve_builtin  a small dense primal-dual interior-point method shipped with the plugin so that a
//                  VGP can be solved where IPOPT is not installed; meant for problems of the size of
//                  the shipped examples (a few hundred constraints), not for large meshes
// Both see the problem in the scaled space the device evaluates in.
#ifndef SRC_ECUDA_ECUDA_NLP_HPP_
#define SRC_ECUDA_ECUDA_NLP_HPP_

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace ecuda_nlp {

struct Problem {
    int n = 0, m = 0, nnz = 0;
    std::vector<double> zl, zu, gl, gu;          // bounds; +-infinity allowed; zl == zu fixes a variable
    const int32_t* irow = nullptr;               // [nnz] triplet pattern, 0-based, sorted by (col,row)
    const int32_t* jcol = nullptr;
    // any output may be null. jac in triplet order. Returns false on evaluation failure.
    std::function<bool(const double* z, double* f, double* g, double* jac, double* grad)> eval;
    // optional exact Hessian of the Lagrangian sigma f + lambda'g: lower triangle in the given pattern (0-based).
    // Used by solve_ipopt (hessian_approximation = exact). solve_builtin keeps its damped-BFGS model: its
    // Schur-complement factorisation needs a positive definite Hessian block, and the exact Hessian of an
    // obstacle-avoidance problem is indefinite (tried: convexifying it with a multiple of I diverges).
    int hnnz = 0;
    const int32_t* hrow = nullptr;
    const int32_t* hcol = nullptr;
    std::function<bool(const double* z, double sigma, const double* lambda, double* hvals)> eval_h;
};

struct Options {
    int max_iter = 200;
    double tol = 1e-6;
    int print_level = 0;
};

struct Result {
    int iterations = 0;
    double objective = 0.0;
    double max_violation = 0.0;  // max bound violation of g and z at the returned point
    double kkt_error = 0.0;
    std::string message;
};

bool have_ipopt();
// return 0 on success (a point satisfying the tolerances), non-zero otherwise with Result::message
int solve_ipopt(const Problem& P, const Options& opt, std::vector<double>* z, Result* out);
int solve_builtin(const Problem& P, const Options& opt, std::vector<double>* z, Result* out);

}  // namespace ecuda_nlp
#endif  // SRC_ECUDA_ECUDA_NLP_HPP_
