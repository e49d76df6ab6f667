// oracle.hpp -- CPU oracle for the eCUDA hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// PARITY: the reference (olasanni1/ETOL) ships no tests, golden vectors or expected outputs for this path
// (SURVEY.md section 4), and part of the arithmetic lives in PSOPT 5.0.0 / ADOL-C / IPOPT, which are not
// vendored and cannot be built here. This oracle restates
//   * the ETOL side from the in-tree sources it cites (src/ePSOPT/ePSOPT.cpp,
//     src/Examples/PSOPT/etol_psopt_example1.cpp, include/ETOL/TrajectoryOptimizer.hpp). PINNED: those two
//     reference files are compiled unmodified against a stub psopt.h (oracle/refstub, `make ref` ->
//     oracle/_ref/libetol_ref.so) and their dae / integrand_cost / events / endpoint_cost / addBounds and the
//     obs / saa lambdas are compared with this oracle at every node (tests/test_oracle_vs_reference.py; the same
//     outputs are committed as tests/golden/c0_ref_pernode.npz);
//   * the PSOPT side from its published algorithm (Legendre/Chebyshev pseudospectral transcription, layout,
//     scaling, colouring, finite-difference step: SURVEY.md Appendix A). PARITY UNPINNED for this part: it is
//     checked only against the known answers derivable by hand from the in-tree formulas (SURVEY.md Appendix B;
//     tests/golden/) and closed-form identities of the method.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/. The product (etol_b200/, src/) never includes, links or calls it.
#ifndef ORACLE_HPP_
#define ORACLE_HPP_

#include <any>
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace oracle {

// ---- callback ABI, as in include/ETOL/ETOL_Types.hpp:25-27,111-117 ------------------------------
using scalar_t = std::any;
using vector_t = std::vector<scalar_t>;
using f_t = std::function<scalar_t(vector_t x, vector_t u, vector_t params,
                                   std::vector<std::string> pnames, std::any k, std::any dt)>;

enum Model { SI2D = 0, PM3D = 1, FW6 = 2, USER = 3 };
// USER: dynamics and running cost given as a recorded tape (semantics of include/ecuda.h,
// "user models"): what a user's ePSOPT-style lambdas compute, replayed node by node
enum TapeOp { T_INPUT = 0, T_CONST, T_ADD, T_SUB, T_MUL, T_DIV, T_NEG, T_POW, T_SQRT, T_SIN, T_COS, T_EXP };
struct TapeNode {
    int op, a, b;
    double imm;
};
struct UserTape {
    int ns = 0, nc = 0;
    bool edges = false;  // static path rows: ellipse per polygon edge (true) or cylinders (false)
    std::vector<TapeNode> nodes;
    std::vector<int> f_out;
    int cost_out = -1;
    // traced path rows: one row per entry, evaluated at every node after the static and the moving-zone rows; may
    // read states 0, 1 and t (the read set of a moving-zone row, whose sparsity they share)
    std::vector<int> row_out;
};
enum CollocationKind { LEGENDRE = 0, CHEBYSHEV = 1 };
enum PatternMode { DENSE_NODE = 0, MODEL_DEPS = 1 };
enum JacMode { JAC_EXACT = 0, JAC_FD_INDEXSET = 1 };

constexpr int DOT_BLOCK = 8;  // canonical blocked summation of D*X

struct Collocation {
    int N = 0;
    std::vector<double> tau, w, D;  // D row-major N x N
};
Collocation make_collocation(int kind, int N);

// ---- VGP data, shaped like what the reference example captures -----------------------------------
using Corner = std::array<double, 3>;   // corner_t, ETOL_Types.hpp:60
using Border = std::vector<Corner>;     // border_t (a std::list in ETOL; order is what matters)
struct Track {                          // track_t, ETOL_Types.hpp:102-105 (2-D datums)
    double radius = 0.0;
    std::vector<double> t, x, y;
};
struct Cylinder {
    double cx, cy, r;
};
struct PhaseData {
    std::vector<Border> borders;      // si2d
    std::vector<Cylinder> cylinders;  // pm3d / fw6
};
struct Instance {
    std::vector<PhaseData> phases;
    std::vector<Track> tracks;  // si2d, shared by the phases
};

struct Spec {
    int model = SI2D;
    int nphases = 1;
    std::vector<int> nnodes;
    std::vector<int> nstatic;  // static path rows per phase (edges or cylinders)
    int ncontrols = 0;         // 0 = model default
    int ntracks = 0, nwaypoints = 0;
    int collocation = LEGENDRE;
    int pattern_mode = DENSE_NODE;
    bool maximize = false;
    int index_base = 0;
    UserTape user;  // model == USER
};

// ---- NLP layout ------------------------------------------------------
struct Layout {
    int ns = 0, nc = 0, ne = 0, nphases = 0;
    std::vector<int> N, npath, zoff, goff, nvars_p, ncons_p;
    int nvars = 0, ncons = 0, nlink = 0, linkoff = 0;
    int iu(int p, int k, int j) const { return zoff[p] + k * nc + j; }
    int ix(int p, int k, int i) const { return zoff[p] + nc * N[p] + k * ns + i; }
    int it0(int p) const { return zoff[p] + (ns + nc) * N[p]; }
    int itf(int p) const { return it0(p) + 1; }
    int rdef(int p, int k, int i) const { return goff[p] + k * ns + i; }
    int rev(int p, int e) const { return goff[p] + ns * N[p] + e; }
    int rpath(int p, int k, int q) const { return goff[p] + ns * N[p] + ne + k * npath[p] + q; }
    int rlast(int p) const { return goff[p] + ns * N[p] + ne + npath[p] * N[p]; }
    int rlink(int a, int i) const { return linkoff + a * (ns + 1) + i; }
};
Layout make_layout(const Spec& s);

struct Structure {
    std::vector<int32_t> irow, jcol;    // sorted by (col,row), 0-based
    std::vector<int32_t> colptr;        // nvars+1
    std::vector<int32_t> group_of_col;  // CPR first-fit groups, natural column order
    int ngroups = 0;
};
Structure make_structure(const Spec& s, const Layout& L);

struct Scaling {
    std::vector<double> sz, isz, sg;  // isz = 1/sz rounded once
    double sf = 1.0;
};

class Problem {
 public:
    explicit Problem(const Spec& s);
    Spec spec;
    Layout L;
    Structure S;
    std::vector<Collocation> col;  // per phase
    Scaling sc;
    void set_scaling(const double* sz, const double* sg, double sf);

    // Reference-style evaluation: per-node std::function callbacks like ePSOPT::dae.
    void eval_f(const Instance& I, const double* zs, double* f) const;
    void eval_g(const Instance& I, const double* zs, double* g) const;
    void eval_grad_f(const Instance& I, const double* zs, double* grad) const;
    void eval_jac_exact(const Instance& I, const double* zs, double* vals) const;
    void eval_jac_fd(const Instance& I, const double* zs, double* vals) const;
    // Tight style: same arithmetic, plain loops (no std::any / std::function). Bit-identical.
    void eval_g_tight(const Instance& I, const double* zs, double* g) const;
    void eval_f_tight(const Instance& I, const double* zs, double* f) const;
    void eval_jac_fd_tight(const Instance& I, const double* zs, double* vals) const;
    // Hessian of the Lagrangian (evaluate.cpp), lower triangle in the order of hess_structure
    void eval_hess(const Instance& I, const double* zs, double sigma, const double* lambda, double* vals) const;
    void hess_structure(std::vector<int32_t>* irow, std::vector<int32_t>* jcol) const;
    // mesh refinement support (tight.cpp): relative local error per mesh interval [sum_p (N_p - 1)];
    // decision vector interpolated onto meshes of nnew[p] nodes (scaled by sz_new when given)
    void ode_error(const Instance& I, const double* zs, double* err) const;
    void resample(const double* zs, const std::vector<int>& nnew, const double* sz_new, std::vector<double>* out) const;
};

// static geometry of one polygon edge, etol_psopt_example1.cpp:164-172,178-179
struct EdgeGeom {
    double xc, yc, radsq, tt, asq, bsq;
};
EdgeGeom edge_geometry(const Corner& a, const Corner& b);

}  // namespace oracle
#endif
