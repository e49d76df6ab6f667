// ecuda_internal.hpp -- data structures shared by the host side of libecuda.so and the kernels.
// Product code. Never includes anything from oracle/.
#ifndef ECUDA_INTERNAL_HPP_
#define ECUDA_INTERNAL_HPP_

#ifndef __CUDACC_RTC__
#include <cstdint>
#include <string>
#include <vector>
#endif

#include "../../include/ecuda.h"

#if defined(__CUDACC__)
#define ECUDA_HD __host__ __device__ __forceinline__
#else
#define ECUDA_HD inline
#endif

namespace ecuda {

// a / d for 0 <= a < 65536 and 1 <= d < 65536 with magic = floor(2^32 / d) + 1 (0 when d <= 1)
ECUDA_HD int fast_div(int a, unsigned magic) {
#if defined(__CUDA_ARCH__)
    return magic ? static_cast<int>(__umulhi(static_cast<unsigned>(a), magic)) : a;
#else
    return magic ? static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(a)) * magic) >> 32) : a;
#endif
}
#ifndef __CUDACC_RTC__
inline unsigned fast_div_magic(int d) { return d <= 1 ? 0u : static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(d)) + 1u; }
#endif

// ---- what the kernels see (passed by value as a __grid_constant__ kernel parameter) -------------
struct PhaseDev {
    int N;          // collocation nodes
    int npath;      // path rows per node = nstat + ntracks
    int nstat;      // static obstacle records of this phase
    int nb;         // ceil(N / ECUDA_DOT_BLOCK)
    int zoff;       // first decision variable of the phase
    int goff;       // first constraint row of the phase
    int nvars;      // (ns+nc)*N + 2
    int inst_off;   // offset (doubles) of this phase's static records inside the instance block
    unsigned mN, mnp, m2np;  // multiply-high reciprocals of N, npath, 2*npath (fast_div; 0 = divisor <= 1)
    unsigned mns, mnt;       // ... of nstat and of npath - nstat (the track rows per node): path_item
    const double* D;    // [N][N] row-major
    const double* Dt;   // [N][N] transposed (Dt[l*N+k] = D[k][l]) : coalesced over rows k
    const double* tau;  // [N]
    const double* w;    // [N]
};

// Path rows as work items: item `it` of the phase's N * npath path rows -> (node k, row q of the node). The static
// rows of all nodes come first, then the track rows, so that the lanes of a warp evaluate the same kind of row
// (a moving-obstacle row costs two divisions and a waypoint search; in node-major order every warp paid for both).
ECUDA_HD void path_item(const PhaseDev& ph, int it, int& k, int& q) {
    const int nS = ph.nstat * ph.N;
    if (it < nS) {
        k = fast_div(it, ph.mns);
        q = it - k * ph.nstat;
    } else {
        const int i2 = it - nS;
        k = fast_div(i2, ph.mnt);
        q = ph.nstat + (i2 - k * (ph.npath - ph.nstat));
    }
}

struct ProbDev {
    int model, ns, nc, ne, nphases;
    int nvars, ncons, nnz, nlink, linkoff;
    int ntracks, nway, track_off, track_size, rec_size, inst_stride;
    int maximize, dense;
    double sf;
    const int* colptr;    // [nvars+1]
    const double* jtmpl;  // [nnz] exact-mode template: the instance-independent D-coupled entries
                          //       sg[row] * D[k][l] / sz[col]; other slots are 0
    const double* isz;    // [nvars]  1/sz
    // [nnz] exact-mode triplet descriptors (ecuda_stream.cuh): every exact triplet is (sg[row] * T) / sz[col] with T
    // either an entry of D or an entry of a small per-instance table; desc packs {T index, row, col}
    const unsigned long long* desc;
    const double* sg;     // [ncons]
    // position of defect row (k,i) inside the node-local part of column X(k,j) / U(k,j); -1 = absent
    signed char xrank[ECUDA_MAX_STATES][ECUDA_MAX_STATES];
    signed char urank[ECUDA_MAX_CONTROLS][ECUDA_MAX_STATES];
    int xcnt[ECUDA_MAX_STATES];
    int ucnt[ECUDA_MAX_CONTROLS];
    PhaseDev ph[ECUDA_MAX_PHASES];
};

// ---- exact-mode triplet descriptors ----------------------------------------------------------------------------------
// One 64-bit word per triplet: low 32 bits = index into the instance's table, bits 32..47 = constraint row (global),
// bits 48..63 = decision variable (phase-local). Table of a phase (ns states, nc controls, N nodes, np path rows):
//   [0] = 1.0, [1] = -1.0,
//   [2 + (k*ns + i)*W + j]            j < ns: d zeta_ki / d x_kj (unscaled, includes D_kk for j == i); ns <= j < ns+nc:
//                                     d zeta_ki / d u_k(j-ns); j = ns+nc, ns+nc+1: d zeta_ki / d t0, d tf;  W = ns+nc+2
//   [TP + (k*np + q)*4 + j]           path row (k,q): d/dx_0, d/dx_1, d/dt0, d/dtf
//   [TD + l*N + k]                    D[k][l] (instance independent, copied in by every CTA)
ECUDA_HD unsigned long long desc_pack(unsigned tab, unsigned row, unsigned lcol) {
    return static_cast<unsigned long long>(tab) | (static_cast<unsigned long long>(row) << 32) |
           (static_cast<unsigned long long>(lcol) << 48);
}
ECUDA_HD int desc_rowtab_width(int ns, int nc) { return ns + nc + 2; }
ECUDA_HD int desc_path_off(int ns, int nc, int N) { return 2 + ns * N * desc_rowtab_width(ns, nc); }
ECUDA_HD int desc_d_off(int ns, int nc, int N, int np) { return desc_path_off(ns, nc, N) + np * N * 4; }
ECUDA_HD int desc_table_size(int ns, int nc, int N, int np) { return desc_d_off(ns, nc, N, np) + N * N; }

// constraint rows of one phase: defects, events, path rows, duration
ECUDA_HD int phase_ncons(const ProbDev& pb, const PhaseDev& ph) { return (pb.ns + ph.npath) * ph.N + pb.ne + 1; }

// mesh-refinement support (ecuda_ode_error): interpolation data of every phase, device pointers
#define ECUDA_MESH_Q 4  // Gauss-Legendre points per mesh interval
struct MeshDev {
    const double* E[ECUDA_MAX_PHASES];   // [(N-1)*Q][N] Lagrange basis at the quadrature points
    const double* dE[ECUDA_MAX_PHASES];  // [(N-1)*Q][N] its derivative with respect to tau
    const double* wq[ECUDA_MAX_PHASES];  // [(N-1)*Q]    quadrature weights (tau units)
    const double* tq[ECUDA_MAX_PHASES];  // [(N-1)*Q]    quadrature points (tau)
    int eoff[ECUDA_MAX_PHASES];          // first interval of the phase in the output row
    int nint;                            // intervals per instance = sum (N_p - 1)
    double* out;                         // [B][nint]
};

// Hessian of the Lagrangian (ecuda_eval_hess): sigma * f + sum_r lambda_r g_r in the solver's scaled space.
// Lower triangle, sorted by (column, row). Per phase: the dense block of every node's variables
// [u_k | x_k], their couplings with t0 and tf, and the 3 time-time entries; everything else is
// structurally zero (D X, events, duration and linkages are linear).
struct HessIO {
    const double* lambda;  // [B][ncons]
    const double* sigma;   // [B] or null (sigma0 for every instance)
    double sigma0;
    double* vals;          // [B][nnz_h]
    int nnz_h;
    int hoff[ECUDA_MAX_PHASES];  // first entry of each phase
};
ECUDA_HD int hess_cu(int ns, int nc) { return nc * (nc + 1) / 2 + nc * (ns + 2); }  // entries of a node's control columns
ECUDA_HD int hess_cx(int ns) { return ns * (ns + 1) / 2 + 2 * ns; }                  // ... of its state columns
ECUDA_HD int hess_phase_nnz(int ns, int nc, int N) { return N * (hess_cu(ns, nc) + hess_cx(ns)) + 3; }

// per-call pointers (device memory)
struct EvalIO {
    const double* x;     // [B][nvars] scaled decision vectors
    const double* inst;  // [B][inst_stride]
    double* f;           // [B] or null
    double* fpart;       // [B][nphases] scratch for multi-phase objective
    double* g;           // [B][ncons] or null
    double* jac;         // [B][nnz] or null
    double* grad;        // [B][nvars] or null
    int jac_mode;
    int batch;
    // fused per-instance summary {f, max bound violation} + all-gather over peer memory (single-phase
    // problems, specialised kernels): active when nranks > 0
    const double* bl;    // [B][ncons] constraint bounds
    const double* bu;
    // compact bounds (single phase; found by ecuda_upload_bounds when every instance has defect bounds 0 and
    // the same pair of path-row bounds): only the event and duration rows are read per instance
    const double* bev;   // [B][2 * (ne + 1)]: lower bounds of the ne event rows and the duration row, then upper
    double plo, phi;     // bounds of every path row
    double* peer[16];    // rank r's gathered buffer [nranks*B][2]
    int nranks, rank;
};

#ifndef __CUDACC_RTC__
// ---- host-side problem description ------------------------------------------------------------------
struct ModelInfo {
    int ns, nc_default, nc_used, rec_size;
    unsigned fx[ECUDA_MAX_STATES];  // states read by f_i   (MODEL_DEPS)
    unsigned fu[ECUDA_MAX_STATES];  // controls read by f_i
    unsigned path_x;                // states read by every path row
};
bool model_info(int model, ModelInfo* out);

// ---- user models (ecuda_usermodel.cpp) -------------------------------------------------------------------
struct UserModel {
    int id = 0, ns = 0, nc = 0, static_kind = 0;
    int nregistered = 0;                 // length of the registered tape (the first nodes)
    std::vector<ecuda_tape_node> nodes;  // the registered tape followed by the derivative nodes
    int f_out[ECUDA_MAX_STATES];
    int cost_out = -1;
    // node ids of the partial derivatives, -1 = identically zero
    int dfdx[ECUDA_MAX_STATES][ECUDA_MAX_STATES], dfdu[ECUDA_MAX_STATES][ECUDA_MAX_CONTROLS];
    int dcdx[ECUDA_MAX_STATES], dcdu[ECUDA_MAX_CONTROLS];
    // traced path rows (may read states 0, 1 and t): output nodes and their partials (-1 = identically zero)
    std::vector<int> row_out, drdx, drdy, drdt;
    std::vector<int> rhess;              // per traced row: d2/dx2, dxdy, dy2, dxdt, dydt, dt2
    // time part of the second derivatives (o < ns: f_o, o == ns: cost): d2/dv dt over [x | u], d2/dt2
    std::vector<int> dvt, dtt;
    bool tdep = false;                   // some f_i or the running cost reads t
    int dfdt[ECUDA_MAX_STATES], dcdt = -1;
    // second derivatives over [x | u]: d2[(o * nv + a) * nv + b], a <= b, o < ns: f_o, o == ns: cost
    std::vector<int> d2;
    unsigned fx[ECUDA_MAX_STATES], fu[ECUDA_MAX_STATES];  // states / controls read by f_i
    std::string source;                                    // generated Model<ECUDA_MODEL_USER>
};
// compiled kernels of one user model for one dot-block count
struct UserImage {
    enum { GENERIC = 0, GRAD = 1, ROWS_FD = 2, ROWS_EXACT = 3, ODE_ERROR = 4, HESS = 5, ROWSN_FD = 6, NKERNELS = 7 };
    std::vector<char> cubin;
    std::string name[NKERNELS];  // lowered kernel names ("" = not compiled)
    std::string log;
};
const UserModel* user_model(int model_id);  // null when the id is not registered
int register_user_model(const ecuda_user_model* um, int nrows, const int32_t* row_out, int32_t* model_id, std::string* err);
// values of the traced path rows of a user model at one point (host)
void user_model_rows(const UserModel& m, double x0, double x1, double t, double* rows);
void user_model_eval(const UserModel& m, const double* x, const double* u, double t, double* f_out, double* cost_out);
void user_model_partials(const UserModel& m, const double* x, const double* u, double t, double* dfdx, double* dfdu,
                         double* dcdx, double* dcdu);
// nb: template argument NB of the kernels (0 = generic block count); rows: also compile k_eval_rows; rowsn_N > 0: also
// the N-specialised finite-difference kernel k_rows_n<USER, rowsn_N, FD, trk> (trk: the problem has path rows beyond the
// static records)
bool user_model_compile(const UserModel& m, int nb, bool rows, int rowsn_N, bool trk, UserImage* out, std::string* err);

struct Collocation {
    int N = 0;
    std::vector<double> tau, w, D;
};
bool build_collocation(int kind, int N, Collocation* out, std::string* err);
// ecuda_mesh.cpp: quadrature points inside every mesh interval with the Lagrange basis (and its
// derivative) there; interpolation matrix between two meshes
struct MeshHost {
    std::vector<double> E, dE, wq, tq;
};
void build_error_mesh(const Collocation& c, MeshHost* out);
void build_resample(const Collocation& from, const Collocation& to, std::vector<double>* R);  // [to.N][from.N]

struct HostProblem {
    ecuda_problem_desc desc{};
    ecuda_dims dims{};
    ModelInfo mi{};
    int ns = 0, nc = 0, ne = 0, nphases = 0, linkoff = 0;
    std::vector<int> N, npath, nstat, zoff, goff, nvars_p, inst_off;
    int track_off = 0;
    std::vector<int32_t> irow, jcol, colptr, group_of_col;
    std::vector<uint64_t> tdesc;  // exact-mode triplet descriptors (see desc_pack), built with the pattern
    std::vector<Collocation> col;
    signed char xrank[ECUDA_MAX_STATES][ECUDA_MAX_STATES];
    signed char urank[ECUDA_MAX_CONTROLS][ECUDA_MAX_STATES];
    int xcnt[ECUDA_MAX_STATES];
    int ucnt[ECUDA_MAX_CONTROLS];
};
// validates desc, fills dims/layout (no pattern). Returns false + err on bad input.
bool build_layout(const ecuda_problem_desc& d, HostProblem* hp, std::string* err);
// pattern (CSC, rows ascending per column) + CPR grouping
void build_structure(HostProblem* hp);
// lower triangle of the Lagrangian Hessian (0-based), sorted by (column, row)
void build_hess_structure(const HostProblem& hp, std::vector<int32_t>* irow, std::vector<int32_t>* jcol);
// layout part of the kernel parameter block (everything except device pointers and sf); needs
// build_layout + build_structure
void fill_probdev(const HostProblem& hp, ProbDev* pd);
// exact-mode Jacobian template: tmpl[e] = (sg[row] * D[k][l]) * isz[col] for every D-coupled triplet
// (defect row (k,j) x state column X(l,j), k != l), 0 elsewhere. Needs the collocation data in hp.col.
void build_jac_template(const HostProblem& hp, const double* isz, const double* sg, std::vector<double>* tmpl);
// ascending indices of the triplets the template does not cover (the per-instance part of an exact Jacobian)
void build_local_index(const HostProblem& hp, std::vector<int32_t>* local);

#endif  // !__CUDACC_RTC__

}  // namespace ecuda
#endif
