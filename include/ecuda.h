/*
 * ecuda.h -- extern "C" boundary of the eCUDA collocation-NLP evaluator (B200 / sm_100a).
 *
 * This is the drop-in boundary for the hot path that ETOL's ePSOPT plugin hands to IPOPT:
 * objective, constraint vector and sparse constraint Jacobian of the pseudospectral NLP, evaluated
 * for a batch of independent vehicle guidance problems (VGPs) per call.
 *
 * Reference interfaces each entry point replaces (paths relative to the ETOL tree):
 *   - ecuda_set_problem            <- ePSOPT::setup()                 src/ePSOPT/ePSOPT.cpp:40-81
 *                                     (nstates/ncontrols/nevents/nodes/npath, collocation method)
 *   - ecuda_upload_instances       <- the VGP data the example lambdas capture:
 *                                     obstacles  src/Examples/PSOPT/etol_psopt_example1.cpp:140-151
 *                                     tracks     src/Examples/PSOPT/etol_psopt_example1.cpp:199-223
 *   - ecuda_eval (f)               <- ePSOPT::integrand_cost + endpoint_cost
 *                                     src/ePSOPT/ePSOPT.cpp:186-216,302-306   (PSOPT quadrature)
 *   - ecuda_eval (g)               <- ePSOPT::dae + ePSOPT::events + ePSOPT::linkages
 *                                     src/ePSOPT/ePSOPT.cpp:218-297           (PSOPT defect assembly)
 *   - ecuda_eval (jac)             <- PSOPT/ADOL-C sparse Jacobian selected at
 *                                     src/ePSOPT/ePSOPT.cpp:64 ("automatic" = exact,
 *                                     "numerical" = column-grouped finite differences)
 *   - ecuda_get_structure          <- the (iRow,jCol) triplet pattern PSOPT hands IPOPT, plus the
 *                                     column grouping ("index sets") of the perturbation scheme
 *   - ecuda_si2d_edge_records      <- edge -> ellipse geometry,
 *                                     src/Examples/PSOPT/etol_psopt_example1.cpp:164-179
 *
 * Conventions: opaque handle, int status (0 = ok, <0 = error, text via ecuda_last_error), no C++
 * types, no exceptions, caller-owned buffers, one handle per (host thread, device). There is NO
 * CPU fallback: every compute entry point fails with ECUDA_ERR_CUDA when no sm_100 device can be
 * used. All floating point is IEEE binary64; all indices are 32-bit.
 *
 * NLP layout (normative). Per phase p with ns states, nc controls, N nodes:
 *   z_p = [ U (node-major, z[k*nc+j]) | X (node-major, z[nc*N + k*ns + i]) | t0 | tf ]
 *   g_p = [ defects (g[k*ns+i]) | events (2*ns) | path (g[ns*N+2*ns + k*npath + q]) | tf - t0 ]
 * Phases are concatenated; linkage rows (state + time continuity between consecutive phases)
 * follow the last phase block. Jacobian triplets are sorted by (column, row).
 */
#ifndef ECUDA_H_
#define ECUDA_H_

#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#else /* runtime compilation of the kernels for a user model: no host headers */
typedef signed char int8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long uintptr_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ECUDA_ABI_VERSION 1
#define ECUDA_MAX_PHASES 8
#define ECUDA_MAX_STATES 8
#define ECUDA_MAX_CONTROLS 8
#define ECUDA_DOT_BLOCK 8 /* block length of the canonical blocked summation used for D*X */

/* status codes */
#define ECUDA_OK 0
#define ECUDA_ERR_ARG (-1)     /* bad argument / unsupported configuration */
#define ECUDA_ERR_STATE (-2)   /* call order violated (e.g. eval before set_problem) */
#define ECUDA_ERR_CUDA (-3)    /* CUDA runtime failure or no usable device */
#define ECUDA_ERR_ALLOC (-4)   /* host or device allocation failed */
#define ECUDA_ERR_PEER (-5)    /* a cross-GPU barrier timed out: a peer did not arrive (gathered rows incomplete) */

/* device models of the VGP callbacks (dynamics, running cost, path constraints) */
enum ecuda_model {
    ECUDA_MODEL_SI2D = 0, /* 2-D single integrator + ellipse-per-edge + moving circles (reference VGP) */
    ECUDA_MODEL_PM3D = 1, /* 3-D point mass (6 states, 3 controls) + vertical cylinders */
    ECUDA_MODEL_FW6 = 2,  /* 6-state fixed-wing kinematics + vertical cylinders */
    ECUDA_MODEL_USER = 3  /* internal tag of the runtime-compiled model; callers use the id returned by
                             ecuda_register_user_model (>= ECUDA_MODEL_USER_BASE) */
};
#define ECUDA_MODEL_USER_BASE 16
#define ECUDA_MAX_USER_MODELS 64
#define ECUDA_MAX_USER_ROWS 16 /* traced path rows per node (ecuda_register_user_model_rows) */

/* ---- user models: dynamics and running cost recorded from callbacks -------------------------------
 * What ePSOPT gets by running the VGP callbacks on ADOL-C adoubles (src/ePSOPT/ePSOPT.cpp:186-276),
 * eCUDA gets as a tape: a straight-line program over the inputs [x_0..x_{ns-1} | u_0..u_{nc-1} | t].
 * Node i may only refer to nodes < i. Semantics are normative (each node is one IEEE-754 double
 * operation, no contraction): POW with an integral exponent e, |e| <= 8, is the left-to-right
 * product a*a*...*a (1/(...) for e < 0, 1 for e = 0); SIN / COS are ecuda_sincos of
 * include/ecuda_detmath.h; other POW exponents and EXP use the platform's pow / exp and are
 * reproducible to rounding only. Dynamics and cost may read t (the ePSOPT callbacks receive the node time as `k`,
 * src/ePSOPT/ePSOPT.cpp:218-260): values, both Jacobian modes (d/dt0, d/dtf through t_k = t0 + (tf - t0)(tau_k + 1)/2)
 * the objective gradient and the Lagrangian Hessian (d2/dv dt, d2/dt2 of f and L) follow. */
enum ecuda_tape_op {
    ECUDA_OP_INPUT = 0, ECUDA_OP_CONST = 1, ECUDA_OP_ADD = 2, ECUDA_OP_SUB = 3, ECUDA_OP_MUL = 4, ECUDA_OP_DIV = 5,
    ECUDA_OP_NEG = 6, ECUDA_OP_POW = 7, ECUDA_OP_SQRT = 8, ECUDA_OP_SIN = 9, ECUDA_OP_COS = 10, ECUDA_OP_EXP = 11
};
typedef struct {
    int32_t op;   /* enum ecuda_tape_op */
    int32_t a, b; /* operand nodes (INPUT: a = input slot; unary: b = -1) */
    int32_t reserved;
    double imm;   /* CONST value / POW exponent */
} ecuda_tape_node;
enum ecuda_static_kind { ECUDA_STATIC_CYLINDER = 0, ECUDA_STATIC_EDGE = 1 };
typedef struct {
    int32_t nstates, ncontrols;        /* 2..ECUDA_MAX_STATES (path rows read states 0 and 1), 1..ECUDA_MAX_CONTROLS */
    int32_t static_kind;               /* record type of the static path rows: enum ecuda_static_kind */
    int32_t nnodes;                    /* tape length */
    const ecuda_tape_node* nodes;
    int32_t f_out[ECUDA_MAX_STATES];   /* node holding dx_i/dt */
    int32_t cost_out;                  /* node holding the running cost */
} ecuda_user_model;

enum ecuda_collocation { ECUDA_LEGENDRE = 0, ECUDA_CHEBYSHEV = 1 };

/* DENSE_NODE: every defect row depends on all states/controls of its node (PSOPT's assumption);
 * MODEL_DEPS: only the variables the model's dynamics actually read. */
enum ecuda_pattern { ECUDA_PATTERN_DENSE_NODE = 0, ECUDA_PATTERN_MODEL_DEPS = 1 };

/* EXACT       ~ PSOPT derivatives="automatic" (the reference default, ePSOPT.cpp:64)
 * FD_INDEXSET ~ PSOPT derivatives="numerical": central differences, step 2^-26*(1+|z|), all
 *               columns of a Curtis-Powell-Reid group perturbed together. */
enum ecuda_jac_mode { ECUDA_JAC_EXACT = 0, ECUDA_JAC_FD_INDEXSET = 1 };

enum ecuda_mem { ECUDA_MEM_HOST = 0, ECUDA_MEM_DEVICE = 1 };

typedef struct ecuda_ctx* ecuda_handle;

typedef struct {
    int32_t model;                      /* enum ecuda_model */
    int32_t nphases;                    /* 1..ECUDA_MAX_PHASES (ePSOPT itself uses 1) */
    int32_t nnodes[ECUDA_MAX_PHASES];   /* collocation nodes per phase (= nsteps+1) */
    int32_t nstatic[ECUDA_MAX_PHASES];  /* static obstacle records per phase (edges / cylinders) */
    int32_t ncontrols;                  /* 0 = model default; si2d accepts >= 2 (extra controls unused) */
    int32_t ntracks;                    /* moving-obstacle tracks per instance (si2d only) */
    int32_t nwaypoints;                 /* waypoints per track (>= 2 when ntracks > 0) */
    int32_t collocation;                /* enum ecuda_collocation */
    int32_t pattern_mode;               /* enum ecuda_pattern */
    int32_t maximize;                   /* 1: running cost is negated (ePSOPT.cpp:212-213) */
    int32_t batch;                      /* number of independent VGP instances B */
    int32_t index_base;                 /* 0 (C) or 1 (Fortran) for ecuda_get_structure */
} ecuda_problem_desc;

typedef struct {
    int32_t nvars;        /* decision variables per instance */
    int32_t ncons;        /* constraint rows per instance */
    int32_t nnz;          /* structural Jacobian non-zeros per instance */
    int32_t ngroups;      /* CPR column groups */
    int32_t nstates;
    int32_t ncontrols;
    int32_t nlinkages;    /* linkage rows */
    int32_t inst_stride;  /* doubles per instance in the instance-data block */
    int32_t rec_size;     /* doubles per static obstacle record */
    int32_t track_size;   /* doubles per track record (1 + 3*nwaypoints) */
} ecuda_dims;

#ifndef __CUDACC_RTC__ /* entry points: host code only (the kernels include this file for the types) */
/* ---- lifetime ------------------------------------------------------------------------------ */
int ecuda_abi_version(void);
int ecuda_create(int device, ecuda_handle* out);
int ecuda_destroy(ecuda_handle h);
const char* ecuda_last_error(ecuda_handle h); /* h may be NULL: error of the last failed create */

/* ---- problem structure (host work, done once per structure) ----------------------------------- */
int ecuda_set_problem(ecuda_handle h, const ecuda_problem_desc* desc);
int ecuda_get_dims(ecuda_handle h, ecuda_dims* out);
/* iRow/jCol: nnz entries sorted by (col,row); group_of_col: nvars entries. Any pointer may be NULL. */
int ecuda_get_structure(ecuda_handle h, int32_t* iRow, int32_t* jCol, int32_t* group_of_col);
/* nodes tau[N], quadrature weights w[N], differentiation matrix D[N*N] (row-major) of one phase */
int ecuda_get_collocation(ecuda_handle h, int phase, double* tau, double* w, double* D);
/* override the collocation data of one phase (e.g. to inject an externally computed D) */
int ecuda_set_collocation(ecuda_handle h, int phase, const double* tau, const double* w,
                          const double* D);
/* PSOPT-style scaling: solver sees z~ = z*sz, g~ = g*sg, f~ = f*sf. NULL = all ones. Shared by
 * the whole batch. Host pointers. */
int ecuda_set_scaling(ecuda_handle h, const double* sz, const double* sg, double sf);

/* ---- per-instance problem data --------------------------------------------------------------- */
/* inst: [B][inst_stride] doubles. Per instance: for each phase p, nstatic[p] static records of
 * rec_size doubles (si2d edge: xc,yc,cos,sin,asq,bsq; cylinder: cx,cy,r^2,0), then ntracks track
 * records (radius, then nwaypoints x (t,x,y)), zero-padded to inst_stride. */
int ecuda_upload_instances(ecuda_handle h, const double* inst, int memkind);
/* optional constraint bounds [B][ncons] used only by ecuda_summary (violation measure) */
int ecuda_upload_bounds(ecuda_handle h, const double* gl, const double* gu, int memkind);

/* ---- evaluation (the hot path) --------------------------------------------------------------- */
/* x: [B][nvars] scaled decision vectors. Outputs (each may be NULL = not requested):
 * f [B], g [B][ncons], jac [B][nnz] in the triplet order of ecuda_get_structure.
 * memkind says where x/f/g/jac live (all the same kind). HOST buffers are staged through device
 * buffers owned by the handle (pinned host memory makes the copies asynchronous). The work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = the handle's own non-blocking stream,
 * NOT the legacy default stream -- pass cudaStreamLegacy explicitly if that is wanted) and is
 * asynchronous with respect to the host for DEVICE buffers; HOST calls return after the results
 * have landed. */
int ecuda_eval(ecuda_handle h, const double* x, double* f, double* g, double* jac, int jac_mode,
               int memkind, void* stream);
/* ---- compact exact Jacobian -----------------------------------------------------------------------
 * In ECUDA_JAC_EXACT mode the D-coupled off-diagonal triplets (defect row (k,j) x state column X(l,j), k != l:
 * 74 % of the triplets of the benchmark shape) are (sg*D[k][l])*(1/sz): the same for every instance. A host
 * consumer does not need them B times over PCIe. ecuda_get_compact_structure returns the split:
 * local_index [nlocal] = ascending triplet indices of the per-instance part; shared_vals [nnz] = the shared
 * values (0 at the per-instance positions; they depend on the scaling and on D: fetch again after
 * ecuda_set_scaling / ecuda_set_collocation). Any pointer may be NULL. No reference counterpart: PSOPT hands
 * IPOPT one instance at a time (src/ePSOPT/ePSOPT.cpp:84). */
int ecuda_get_compact_structure(ecuda_handle h, int32_t* nlocal, int32_t* local_index, double* shared_vals);
/* as ecuda_eval(..., ECUDA_JAC_EXACT, ...), but the Jacobian output is jac_local [B][nlocal]: the per-instance
 * triplets in the order of local_index. The full array is written to handle-owned device scratch by the same
 * evaluation kernel and gathered on the device. batch <= 65535. */
int ecuda_eval_compact(ecuda_handle h, const double* x, double* f, double* g, double* jac_local, int memkind, void* stream);
/* host helper (no GPU): jac_full [batch][nnz] from the compact form; bit-identical to what ecuda_eval returns */
int ecuda_splice_jacobian(const double* shared_vals, const int32_t* local_index, int32_t nnz, int32_t nlocal,
                          const double* jac_local, int32_t batch, double* jac_full);
/* gradient of the (scaled) objective, [B][nvars]; exact. */
int ecuda_eval_grad_f(ecuda_handle h, const double* x, double* grad, int memkind, void* stream);
/* per-instance summary [B][2] = { f, max bound violation of g } (needs ecuda_upload_bounds) */
int ecuda_summary(ecuda_handle h, const double* x, double* out, int memkind, void* stream);
/* same reduction over f [B] and g [B][ncons] already on the device (e.g. the outputs of a previous
 * ecuda_eval on the same stream); out [B][2] on the device. One warp per instance, shuffle max. */
int ecuda_summarize(ecuda_handle h, const double* f_dev, const double* g_dev, double* out_dev, void* stream);
/* sharded runs (one handle per GPU, one process per GPU): the same reduction fused with the all-gather
 * of the rows over NVLink peer memory. peer_out[r] is rank r's gathered buffer [nranks*B][2] (device
 * pointers valid on this device, e.g. from a symmetric-memory rendezvous; 16-byte aligned); this rank's
 * rows land at [rank*B, (rank+1)*B) of every buffer. The caller issues a cross-GPU barrier on the same
 * stream afterwards. nranks <= 16. */
int ecuda_summarize_allgather(ecuda_handle h, const double* f_dev, const double* g_dev, void* const* peer_out,
                              int nranks, int rank, void* stream);
/* evaluation and exchange in ONE kernel (single-phase problems on the specialised kernels): as ecuda_eval
 * with device pointers, and every CTA finishes by storing its instance's {f, max bound violation} row
 * into all ranks' gathered buffers, as ecuda_summarize_allgather does. f and g are required. */
int ecuda_eval_allgather(ecuda_handle h, const double* x, double* f, double* g, double* jac, int jac_mode,
                         void* const* peer_out, int nranks, int rank, void* stream);
/* cross-GPU barrier for the fused exchanges above: peer_flags[r] is rank r's array of nranks 64-bit
 * counters in peer-visible memory (zero-initialised once, e.g. a symmetric-memory buffer); step must
 * grow by one per call (first call: 1). Enqueued on `stream` after the kernel whose peer stores it
 * publishes; when it completes on every rank, every rank's rows of this step are visible everywhere.
 * The wait is bounded (10 s, or ECUDA_PEER_TIMEOUT_MS in the environment): a peer that never arrives does not hang
 * the GPU; the failure is recorded in the handle (sticky) and reported by ecuda_sync (ECUDA_ERR_PEER) and by
 * ecuda_peer_barrier_status. */
int ecuda_peer_barrier(ecuda_handle h, void* const* peer_flags, int nranks, int rank, uint64_t step, void* stream);
/* synchronises `stream` (NULL: the handle's own) and reports whether any peer barrier issued through this handle has
 * timed out since the status was last cleared: *timed_out 0/1, *step the first failed step, *late_rank the first
 * rank that had not arrived (-1: none). reset != 0 clears the status. There is no reference counterpart (the
 * reference has no parallelism, src/ePSOPT/CMakeLists.txt:10,25-27). */
int ecuda_peer_barrier_status(ecuda_handle h, void* stream, int32_t* timed_out, uint64_t* step, int32_t* late_rank, int reset);
/* waits for the handle's own stream; ECUDA_ERR_PEER when a peer barrier of this handle has timed out (sticky) */
int ecuda_sync(ecuda_handle h);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches evidence) */
int64_t ecuda_launch_count(ecuda_handle h);
/* measured FP64 FMA throughput of the device (TFLOP/s, register-only microbenchmark): the FP64 roof
 * the finite-difference Jacobian kernel is reported against */
int ecuda_fp64_peak(ecuda_handle h, double* tflops);

/* IPOPT TNLP-shaped single-instance shims (B must be 1; host pointers). values==NULL in
 * eval_jac_g returns the structure, as IPOPT's convention requires. */
int ecuda_ipopt_eval_f(ecuda_handle h, int n, const double* x, int new_x, double* obj);
int ecuda_ipopt_eval_grad_f(ecuda_handle h, int n, const double* x, int new_x, double* grad);
int ecuda_ipopt_eval_g(ecuda_handle h, int n, const double* x, int new_x, int m, double* g);
int ecuda_ipopt_eval_jac_g(ecuda_handle h, int n, const double* x, int new_x, int m, int nele_jac,
                           int32_t* iRow, int32_t* jCol, double* values);
int ecuda_set_ipopt_jac_mode(ecuda_handle h, int jac_mode);
/* TNLP::eval_h; values == NULL: structure query */
int ecuda_ipopt_eval_h(ecuda_handle h, int n, const double* x, int new_x, double obj_factor, int m, const double* lambda,
                       int new_lambda, int nele_hess, int32_t* iRow, int32_t* jCol, double* values);

/* ---- Hessian of the Lagrangian (ePSOPT asks PSOPT / IPOPT for hessian = "exact", ePSOPT.cpp:65) -------- */
/* sigma * f + sum_r lambda_r g_r in the solver's scaled space (f, g as ecuda_eval returns them). Lower
 * triangle, sorted by (column, row): per phase the dense block of every node's variables [u_k | x_k],
 * their couplings with t0 / tf and the 3 time-time entries -- everything else is structurally zero.
 * nnz_h = sum_p N_p * (nc(nc+1)/2 + nc(ns+2) + ns(ns+1)/2 + 2 ns) + 3. */
int ecuda_get_hess_structure(ecuda_handle h, int32_t* nnz_h, int32_t* iRow, int32_t* jCol); /* any pointer may be NULL */
/* sigma: [B] objective factors, or NULL to use sigma0 for every instance; lambda: [B][ncons]; vals: [B][nnz_h] */
int ecuda_eval_hess(ecuda_handle h, const double* x, const double* sigma, double sigma0, const double* lambda,
                    double* vals, int memkind, void* stream);

/* ---- mesh refinement support (what PSOPT does between NLP solves, ePSOPT.cpp:69-71) ---------------- */
/* Relative local discretisation error of the collocation solution x on every mesh interval (Betts'
 * estimate, see etol_b200/csrc/ecuda_mesh.cpp): err[B][sum_p (nnodes[p] - 1)], phases back to back. */
int ecuda_ode_error(ecuda_handle h, const double* x, double* err, int memkind, void* stream);
/* The decision vectors x (this problem's layout and scaling) interpolated onto meshes with
 * nnodes_new[p] nodes per phase: x_new[B][sum_p ((ns+nc)*nnodes_new[p] + 2)], multiplied by sz_new
 * (same memory kind as x; NULL = unscaled). The warm start of the next solve. */
int ecuda_resample(ecuda_handle h, const double* x, const int32_t* nnodes_new, const double* sz_new, double* x_new,
                   int memkind, void* stream);

/* ---- host-side helpers (no GPU needed) ------------------------------------------------------- */
/* polygon (ncorners x,y pairs, closed implicitly) -> ncorners edge records of 6 doubles each */
int ecuda_si2d_edge_records(const double* corners_xy, int ncorners, double* rec6);
/* dims/structure without a device: same results as set_problem + get_dims/get_structure */
int ecuda_host_dims(const ecuda_problem_desc* desc, ecuda_dims* out);
int ecuda_host_structure(const ecuda_problem_desc* desc, int32_t* iRow, int32_t* jCol,
                         int32_t* group_of_col);
int ecuda_host_collocation(int kind, int nnodes, double* tau, double* w, double* D);
/* the index part of ecuda_get_compact_structure without a device (local_index may be NULL to query nlocal) */
int ecuda_host_compact_structure(const ecuda_problem_desc* desc, int32_t* nlocal, int32_t* local_index);
int ecuda_host_hess_structure(const ecuda_problem_desc* desc, int32_t* nnz_h, int32_t* iRow, int32_t* jCol);
/* interpolation data of ecuda_ode_error: quadrature points tq / weights wq [(N-1)*4] inside the mesh
 * intervals, Lagrange basis E and its derivative dE there [(N-1)*4][N]; matrix R [N_to][N_from] of
 * ecuda_resample. Any output may be NULL. */
int ecuda_host_error_mesh(int kind, int nnodes, double* tq, double* wq, double* E, double* dE);
int ecuda_host_resample_matrix(int kind, int nnodes_from, int nnodes_to, double* R);
/* the device models evaluated on the host at one point: state derivatives f_out[nstates] and running
 * cost. Used by the eCUDA plugin to verify that user callbacks and device model agree (eCUDA.hpp); not
 * an evaluation path. u holds the model's controls (2 for si2d, 3 for pm3d / fw6). */
/* Registers a user model (process-wide, thread-safe) and returns its id for ecuda_problem_desc.model.
 * The library differentiates the tape symbolically, derives the dependency masks used by
 * ECUDA_PATTERN_MODEL_DEPS and generates the CUDA source of the model; the kernels are compiled for it
 * with NVRTC (libnvrtc.so.12 must be loadable) at the first ecuda_set_problem that uses the id.
 * No GPU is needed to register. err (may be NULL) receives a message on failure. */
int ecuda_register_user_model(const ecuda_user_model* m, int32_t* model_id, char* err, size_t errlen);
/* as ecuda_register_user_model, with traced path constraints: row_out[nrows] are tape nodes, each one a path row that
 * is evaluated at every node after the static and the moving-zone rows of the problem (npath = nstatic + ntracks +
 * nrows) -- what ePSOPT does with every entry of _constraints (src/ePSOPT/ePSOPT.cpp:262-270) when it is none of the
 * built-in zone rows. A traced row may read states 0, 1 and t: the read set of a moving-zone row, whose sparsity
 * (columns x_0, x_1 of the node, t0, tf) it shares. Values, both Jacobian modes and the Lagrangian Hessian. */
int ecuda_register_user_model_rows(const ecuda_user_model* m, int32_t nrows, const int32_t* row_out, int32_t* model_id,
                                   char* err, size_t errlen);
/* the generated CUDA source of a registered model (NUL-terminated; *needed = bytes incl. NUL; buf may be NULL) */
int ecuda_user_model_source(int32_t model_id, char* buf, size_t buflen, size_t* needed);
/* compiles the kernels of a registered model for `nnodes` collocation nodes without a device and returns the
 * size of the sm_100a image (0 on failure; log, if not NULL, receives the compiler log). Build check only. */
int ecuda_user_model_compile_check(int32_t model_id, int nnodes, size_t* image_bytes, char* log, size_t loglen);
int ecuda_host_model_eval(int model, const double* x, const double* u, double t, double* f_out, double* cost_out);
/* path rows of phase 0 of one instance block (layout of ecuda_upload_instances) at position (x,y), time t */
int ecuda_host_path_eval(const ecuda_problem_desc* desc, const double* inst, double x, double y, double t, double* rows);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* ECUDA_H_ */
