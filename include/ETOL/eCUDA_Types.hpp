// eCUDA_Types.hpp -- solver-side types of the eCUDA plugin, the counterpart of the reference's
// include/ETOL/ePSOPT_Types.hpp (which pulls in psopt.h and names fout_psopt_t).
//
// ePSOPT exposes PSOPT's own Prob / Alg / Sol structs through getProblem() / getAlgorithm() /
// getSolution() (include/ETOL/ePSOPT.hpp:56-68) and the example edits them between setup() and
// solve() (src/Examples/PSOPT/etol_psopt_example1.cpp:86-99). eCUDA does the same with the three
// plain structs below; field names follow PSOPT's where the meaning is the same.
#ifndef INCLUDE_ETOL_ECUDA_TYPES_HPP_
#define INCLUDE_ETOL_ECUDA_TYPES_HPP_

#include <cstdint>
#include <string>
#include <vector>

#include <ecuda.h>

#include <utility>

namespace ETOL {

// algorithm options; defaults mirror ePSOPT::setup() (src/ePSOPT/ePSOPT.cpp:62-72)
struct ecuda_alg_t {
    std::string nlp_method = "IPOPT";          // "IPOPT" (when linked) or "builtin"
    std::string scaling = "automatic";         // "automatic" (PSOPT-like variable scaling) or "none"
    std::string derivatives = "automatic";     // "automatic" = exact Jacobian, "numerical" = index-set FD
    std::string collocation_method = "Legendre";  // or "Chebyshev"
    std::string hessian = "exact";             // IPOPT driver: "exact" = device Hessian of the Lagrangian (what ePSOPT.cpp:65
                                               // asks PSOPT for) or "limited-memory"; the built-in driver always uses
                                               // its own quasi-Newton model
    int nlp_iter_max = 200;
    double nlp_tolerance = 1.e-6;
    // mesh refinement between NLP solves, as ePSOPT configures PSOPT (ePSOPT.cpp:69-71): "automatic" re-solves
    // on more nodes until the relative local discretisation error (ecuda_ode_error) is below ode_tolerance;
    // "manual" solves once on nsteps+1 nodes
    std::string mesh_refinement = "automatic";
    int mr_max_iterations = 10;
    double ode_tolerance = 1.e-4;
    int mr_initial_increment = 10;          // nodes added by the first refinement
    double mr_max_increment_factor = 0.4;   // later refinements add at most this fraction of the node count
    int print_level = 0;
    int device = 0;                            // CUDA device ordinal
};

// the transcribed NLP of one instance: sizes, bounds, structure (filled by setup())
struct ecuda_prob_t {
    ecuda_problem_desc desc{};
    ecuda_dims dims{};
    std::vector<double> zl, zu;        // [nvars]  variable bounds (unscaled)
    std::vector<double> gl, gu;        // [ncons]  constraint bounds (unscaled)
    std::vector<double> guess;         // [nvars]  initial point (unscaled); zeros + time grid like ePSOPT
    std::vector<double> sz, sg;        // scaling handed to the device (all ones when scaling == "none")
    double sf = 1.0;
    std::vector<int32_t> iRow, jCol;   // [nnz]    Jacobian triplet pattern, sorted by (col,row)
    std::vector<int32_t> group_of_col; // [nvars]  Curtis-Powell-Reid column groups
    std::vector<double> tau, w;        // collocation nodes and quadrature weights
    std::vector<std::string> path_names;  // parameter name of every path row of a node
};

// result of solve()
struct ecuda_sol_t {
    int error_flag = 0;
    std::string error_msg;
    double cost = 0.0;
    int nlp_iterations = 0;
    double max_violation = 0.0;
    std::vector<double> z;  // final decision vector (unscaled)
    // one entry per NLP solve: node count and the largest relative local error of its solution
    std::vector<std::pair<int, double>> mesh_history;
};

}  // namespace ETOL
#endif  // INCLUDE_ETOL_ECUDA_TYPES_HPP_
