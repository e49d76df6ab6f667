// capi.cpp (oracle) -- extern "C" surface of the CPU oracle for ctypes (tests/, bench.py cpu_baseline).
// TEST INFRASTRUCTURE (see oracle.hpp header). Nothing here is part of the product.
#include <omp.h>

#include <chrono>
#include <cstring>
#include <exception>
#include <string>

#include "../include/ecuda_detmath.h"
#include "oracle.hpp"

using namespace oracle;

namespace {
struct Handle {
    Problem* P = nullptr;
    std::vector<Instance> inst;
    std::string err;
};
thread_local std::string g_err;
UserTape g_staged;  // consumed by the next oracle_create with model == USER
}  // namespace

extern "C" {

struct oracle_desc {  // field-for-field the same meaning as ecuda_problem_desc (include/ecuda.h)
    int32_t model, nphases;
    int32_t nnodes[8];
    int32_t nstatic[8];
    int32_t ncontrols, ntracks, nwaypoints, collocation, pattern_mode, maximize, batch, index_base;
};

const char* oracle_last_error() { return g_err.c_str(); }

// the tape of a user model (op / a / b / imm per node, include/ecuda.h "user models")
void oracle_stage_user_tape(int ns, int nc, int edges, int n, const int32_t* op, const int32_t* a, const int32_t* b,
                            const double* imm, const int32_t* f_out, int cost_out) {
    UserTape t;
    t.ns = ns;
    t.nc = nc;
    t.edges = edges != 0;
    for (int i = 0; i < n; ++i) t.nodes.push_back(TapeNode{op[i], a[i], b[i], imm[i]});
    t.f_out.assign(f_out, f_out + ns);
    t.cost_out = cost_out;
    g_staged = t;
}

// traced path rows of the staged tape (call after oracle_stage_user_tape)
void oracle_stage_user_rows(int n, const int32_t* row_out) { g_staged.row_out.assign(row_out, row_out + n); }

void* oracle_create(const oracle_desc* d) {
    try {
        Spec s;
        s.model = d->model;
        s.nphases = d->nphases;
        for (int p = 0; p < d->nphases; ++p) {
            s.nnodes.push_back(d->nnodes[p]);
            s.nstatic.push_back(d->nstatic[p]);
        }
        s.ncontrols = d->ncontrols;
        s.ntracks = d->ntracks;
        s.nwaypoints = d->nwaypoints;
        s.collocation = d->collocation;
        s.pattern_mode = d->pattern_mode;
        s.maximize = d->maximize != 0;
        s.index_base = d->index_base;
        if (s.model == USER) s.user = g_staged;
        Handle* h = new Handle;
        h->P = new Problem(s);
        h->inst.resize(d->batch);
        for (auto& I : h->inst) I.phases.resize(s.nphases);
        return h;
    } catch (std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void oracle_destroy(void* hv) {
    Handle* h = static_cast<Handle*>(hv);
    if (!h) return;
    delete h->P;
    delete h;
}

// dims: nvars, ncons, nnz, ngroups, ns, nc, nlink
void oracle_dims(void* hv, int32_t* out7) {
    Handle* h = static_cast<Handle*>(hv);
    out7[0] = h->P->L.nvars;
    out7[1] = h->P->L.ncons;
    out7[2] = static_cast<int32_t>(h->P->S.irow.size());
    out7[3] = h->P->S.ngroups;
    out7[4] = h->P->L.ns;
    out7[5] = h->P->L.nc;
    out7[6] = h->P->L.nlink;
}

void oracle_structure(void* hv, int32_t* irow, int32_t* jcol, int32_t* group) {
    Handle* h = static_cast<Handle*>(hv);
    const Structure& S = h->P->S;
    int base = h->P->spec.index_base;
    for (size_t e = 0; e < S.irow.size(); ++e) {
        if (irow) irow[e] = S.irow[e] + base;
        if (jcol) jcol[e] = S.jcol[e] + base;
    }
    if (group) std::memcpy(group, S.group_of_col.data(), sizeof(int32_t) * S.group_of_col.size());
}

void oracle_collocation(void* hv, int phase, double* tau, double* w, double* D) {
    Handle* h = static_cast<Handle*>(hv);
    const Collocation& c = h->P->col[phase];
    std::memcpy(tau, c.tau.data(), sizeof(double) * c.N);
    std::memcpy(w, c.w.data(), sizeof(double) * c.N);
    std::memcpy(D, c.D.data(), sizeof(double) * c.N * c.N);
}

int oracle_make_collocation(int kind, int N, double* tau, double* w, double* D) {
    try {
        Collocation c = make_collocation(kind, N);
        std::memcpy(tau, c.tau.data(), sizeof(double) * N);
        std::memcpy(w, c.w.data(), sizeof(double) * N);
        std::memcpy(D, c.D.data(), sizeof(double) * N * N);
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

void oracle_set_scaling(void* hv, const double* sz, const double* sg, double sf) {
    static_cast<Handle*>(hv)->P->set_scaling(sz, sg, sf);
}

// raw VGP data -------------------------------------------------------------------------------------
// one polygon (ncorners x (x,y,z)) appended to instance b, phase p
void oracle_add_border(void* hv, int b, int p, const double* xyz, int ncorners) {
    Handle* h = static_cast<Handle*>(hv);
    Border bd;
    for (int i = 0; i < ncorners; ++i) bd.push_back({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
    h->inst[b].phases[p].borders.push_back(bd);
}
void oracle_add_track(void* hv, int b, double radius, int nway, const double* t, const double* x,
                      const double* y) {
    Handle* h = static_cast<Handle*>(hv);
    Track tr;
    tr.radius = radius;
    tr.t.assign(t, t + nway);
    tr.x.assign(x, x + nway);
    tr.y.assign(y, y + nway);
    h->inst[b].tracks.push_back(tr);
}
// cylinders for the whole batch: cyl[B][sum_p nstatic[p]][3] = (cx, cy, r)
void oracle_set_cylinders(void* hv, const double* cyl) {
    Handle* h = static_cast<Handle*>(hv);
    const Spec& s = h->P->spec;
    size_t o = 0;
    for (auto& I : h->inst)
        for (int p = 0; p < s.nphases; ++p) {
            I.phases[p].cylinders.clear();
            for (int c = 0; c < s.nstatic[p]; ++c, o += 3)
                I.phases[p].cylinders.push_back({cyl[o], cyl[o + 1], cyl[o + 2]});
        }
}

// edge geometry KAT surface: returns xc, yc, radsq, tt, asq, bsq
void oracle_edge_geometry(const double* a_xyz, const double* b_xyz, double* out6) {
    EdgeGeom g = edge_geometry({a_xyz[0], a_xyz[1], a_xyz[2]}, {b_xyz[0], b_xyz[1], b_xyz[2]});
    out6[0] = g.xc;
    out6[1] = g.yc;
    out6[2] = g.radsq;
    out6[3] = g.tt;
    out6[4] = g.asq;
    out6[5] = g.bsq;
}

// batched evaluation ---------------------------------------------------------------------------------
// style: 0 = reference-style (std::function/std::any per node), 1 = tight loops.
// jac_mode: 0 exact, 1 FD index-set. Any output pointer may be NULL. Returns wall seconds, <0 on error.
double oracle_eval_batch(void* hv, int first, int count, const double* x, double* f, double* g,
                         double* jac, double* grad, int jac_mode, int style, int nthreads) {
    Handle* h = static_cast<Handle*>(hv);
    const Problem& P = *h->P;
    const size_t nv = P.L.nvars, ng = P.L.ncons, nz = P.S.irow.size();
    std::string err;
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int b = 0; b < count; ++b) {
        try {
            const Instance& I = h->inst[first + b];
            const double* xb = x + b * nv;
            if (style == 0) {
                if (f) P.eval_f(I, xb, f + b);
                if (g) P.eval_g(I, xb, g + b * ng);
                if (jac) {
                    if (jac_mode == JAC_EXACT)
                        P.eval_jac_exact(I, xb, jac + b * nz);
                    else
                        P.eval_jac_fd(I, xb, jac + b * nz);
                }
            } else {
                if (f) P.eval_f_tight(I, xb, f + b);
                if (g) P.eval_g_tight(I, xb, g + b * ng);
                if (jac) {
                    if (jac_mode == JAC_EXACT)
                        P.eval_jac_exact(I, xb, jac + b * nz);
                    else
                        P.eval_jac_fd_tight(I, xb, jac + b * nz);
                }
            }
            if (grad) P.eval_grad_f(I, xb, grad + b * nv);
        } catch (std::exception& e) {
#pragma omp critical
            err = e.what();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (!err.empty()) {
        g_err = err;
        return -1.0;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Hessian of the Lagrangian: structure (nnz returned; arrays may be null) and values [count][nnz_h]
int oracle_hess_structure(void* hv, int32_t* irow, int32_t* jcol) {
    Handle* h = static_cast<Handle*>(hv);
    std::vector<int32_t> ir, jc;
    h->P->hess_structure(&ir, &jc);
    for (size_t e = 0; e < ir.size(); ++e) {
        if (irow) irow[e] = ir[e] + h->P->spec.index_base;
        if (jcol) jcol[e] = jc[e] + h->P->spec.index_base;
    }
    return static_cast<int>(ir.size());
}
int oracle_eval_hess(void* hv, int first, int count, const double* x, const double* sigma, const double* lambda,
                     double* vals) {
    Handle* h = static_cast<Handle*>(hv);
    const Problem& P = *h->P;
    const size_t nnz = oracle_hess_structure(hv, nullptr, nullptr);
    std::string err;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < count; ++b) {
        try {
            P.eval_hess(h->inst[first + b], x + b * P.L.nvars, sigma[b], lambda + b * P.L.ncons, vals + b * nnz);
        } catch (std::exception& e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) {
        g_err = err;
        return -1;
    }
    return 0;
}

// mesh refinement support: err[count][sum_p (N_p - 1)]; x_new[count][nvars_new]
int oracle_ode_error(void* hv, int first, int count, const double* x, double* err) {
    Handle* h = static_cast<Handle*>(hv);
    const Problem& P = *h->P;
    size_t nint = 0;
    for (int n : P.L.N) nint += n - 1;
    try {
        for (int b = 0; b < count; ++b) P.ode_error(h->inst[first + b], x + b * P.L.nvars, err + b * nint);
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return 0;
}
int oracle_resample(void* hv, int count, const double* x, const int32_t* nnew, const double* sz_new, double* x_new) {
    Handle* h = static_cast<Handle*>(hv);
    const Problem& P = *h->P;
    try {
        std::vector<int> nn(nnew, nnew + P.L.nphases);
        std::vector<double> out;
        for (int b = 0; b < count; ++b) {
            P.resample(x + static_cast<size_t>(b) * P.L.nvars, nn, sz_new, &out);
            std::memcpy(x_new + b * out.size(), out.data(), sizeof(double) * out.size());
        }
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return 0;
}

int oracle_max_threads() { return omp_get_max_threads(); }

// host build of the shared deterministic sincos, for tests/test_detmath.py
void oracle_sincos(const double* x, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) ecuda_sincos(x[i], s + i, c + i);
}

}  // extern "C"
