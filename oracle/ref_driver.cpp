// ref_driver.cpp -- TEST INFRASTRUCTURE: runs the reference's OWN ePSOPT callbacks.
//
// Linked with /root/reference/src/ePSOPT/ePSOPT.cpp and /root/reference/src/Examples/PSOPT/etol_psopt_example1.cpp,
// both compiled UNMODIFIED from where they lie (the example with -Dmain=ref_example_main) against
// oracle/refstub/psopt.h, and with this repository's VGP container (src/TrajectoryOptimizer). The functions
// below repeat the call sequence of the example's main() (etol_psopt_example1.cpp:41-66: loadConfigs,
// setMaximize, setObjective, setGradient, obsConstraint / saaConstraint + setConstraints, setup) and then call the
// static callbacks PSOPT would call -- ePSOPT::dae, ::integrand_cost, ::events, ::endpoint_cost
// (ePSOPT.cpp:186-306) -- at node inputs chosen by the tests, and read back the bounds ePSOPT::addBounds
// (ePSOPT.cpp:125-155) wrote into the Prob structure. Values come with their tangents (the stub's adouble is a
// dual number), which is what derivatives="automatic" (ePSOPT.cpp:64) differentiates.
// Only tests/ and tests/golden/make_golden.py load the resulting oracle/_ref/libetol_ref.so.
#include <ETOL/ePSOPT.hpp>

#include <cstring>
#include <string>

// free functions of the reference example (external linkage there)
ETOL::scalar_t objFunction(F_ARGS);
ETOL::scalar_t dxdt(F_ARGS);
ETOL::scalar_t dydt(F_ARGS);
ETOL::f_t obsConstraint(ETOL::TrajectoryOptimizer*);
ETOL::f_t saaConstraint(ETOL::TrajectoryOptimizer*);

namespace {
struct RefProblem {
    ETOL::ePSOPT tp;
    ETOL::f_t obj, xdot, ydot, obs, saa;
    Workspace ws;
};
}  // namespace

extern "C" {

// the example's main() up to and including setup()
void* ref_open(const char* xml_path, int maximize) {
    RefProblem* r = new RefProblem;
    ETOL::TrajectoryOptimizer* t = &r->tp;
    t->loadConfigs(xml_path);
    t->setMaximize(maximize != 0);
    r->obj = &objFunction;
    t->setObjective(&r->obj);
    r->xdot = &dxdt;
    r->ydot = &dydt;
    t->setGradient({&r->xdot, &r->ydot});
    r->obs = obsConstraint(t);
    r->saa = saaConstraint(t);
    t->setConstraints({&r->obs, &r->saa});
    t->setup();
    r->ws.problem = r->tp.getProblem();
    return r;
}
void ref_close(void* h) { delete static_cast<RefProblem*>(h); }

// nstates, ncontrols, nevents, npath, nodes
void ref_dims(void* h, int* out5) {
    Prob* p = static_cast<RefProblem*>(h)->tp.getProblem();
    out5[0] = p->phases(1).nstates;
    out5[1] = p->phases(1).ncontrols;
    out5[2] = p->phases(1).nevents;
    out5[3] = p->phases(1).npath;
    out5[4] = p->phases(1).nodes(0);
}
// bounds written by ePSOPT::addBounds: lower then upper of states, controls, events, path; {t0 lo, t0 hi, tf lo, tf hi}
void ref_bounds(void* h, double* xs, double* us, double* ev, double* path, double* times4) {
    Prob* p = static_cast<RefProblem*>(h)->tp.getProblem();
    Phase& ph = p->phases(1);
    for (int i = 0; i < ph.nstates; ++i) { xs[i] = ph.bounds.lower.states(i); xs[ph.nstates + i] = ph.bounds.upper.states(i); }
    for (int i = 0; i < ph.ncontrols; ++i) { us[i] = ph.bounds.lower.controls(i); us[ph.ncontrols + i] = ph.bounds.upper.controls(i); }
    for (int i = 0; i < ph.nevents; ++i) { ev[i] = ph.bounds.lower.events(i); ev[ph.nevents + i] = ph.bounds.upper.events(i); }
    for (int i = 0; i < ph.npath; ++i) { path[i] = ph.bounds.lower.path(i); path[ph.npath + i] = ph.bounds.upper.path(i); }
    times4[0] = ph.bounds.lower.StartTime;
    times4[1] = ph.bounds.upper.StartTime;
    times4[2] = ph.bounds.lower.EndTime;
    times4[3] = ph.bounds.upper.EndTime;
}
// algorithm options chosen by ePSOPT::setup (strings joined with '|')
int ref_algorithm(void* h, char* out, int cap, double* nums5) {
    Alg* a = static_cast<RefProblem*>(h)->tp.getAlgorithm();
    std::string s = a->nlp_method + "|" + a->scaling + "|" + a->derivatives + "|" + a->hessian + "|" + a->collocation_method +
                    "|" + a->mesh_refinement;
    std::strncpy(out, s.c_str(), static_cast<size_t>(cap) - 1);
    out[cap - 1] = 0;
    nums5[0] = a->nlp_iter_max;
    nums5[1] = a->nlp_tolerance;
    nums5[2] = a->mr_max_iterations;
    nums5[3] = a->ode_tolerance;
    nums5[4] = a->print_level;
    return 0;
}

// ePSOPT::dae at one node. Tangent directions: states 0..ns-1, then controls, then t (ns + nc + 1 <= ADOUBLE_NDIR).
// derivatives[ns], path[npath]; d_derivatives[ns][ndir], d_path[npath][ndir] (may be null)
int ref_dae(void* h, const double* states, const double* controls, double t, double* derivatives, double* path,
            double* d_derivatives, double* d_path) {
    RefProblem* r = static_cast<RefProblem*>(h);
    Phase& ph = r->tp.getProblem()->phases(1);
    const int ns = ph.nstates, nc = ph.ncontrols, np = ph.npath, ndir = ns + nc + 1;
    if (ndir > ADOUBLE_NDIR) return -1;
    std::vector<adouble> x(ns), u(nc), dx(ns), p(np > 0 ? np : 1);
    for (int i = 0; i < ns; ++i) { x[i] = states[i]; x[i].d[i] = 1.0; }
    for (int i = 0; i < nc; ++i) { u[i] = controls[i]; u[i].d[ns + i] = 1.0; }
    adouble tt = t;
    tt.d[ns + nc] = 1.0;
    ETOL::ePSOPT::dae(dx.data(), p.data(), x.data(), u.data(), nullptr, tt, nullptr, 1, &r->ws);
    for (int i = 0; i < ns; ++i) {
        derivatives[i] = dx[i].value();
        if (d_derivatives) for (int k = 0; k < ndir; ++k) d_derivatives[i * ndir + k] = dx[i].d[k];
    }
    for (int q = 0; q < np; ++q) {
        path[q] = p[q].value();
        if (d_path) for (int k = 0; k < ndir; ++k) d_path[q * ndir + k] = p[q].d[k];
    }
    return 0;
}
// ePSOPT::integrand_cost at one node (+ tangents over the same directions), and endpoint_cost
double ref_integrand_cost(void* h, const double* states, const double* controls, double t, double* d_cost) {
    RefProblem* r = static_cast<RefProblem*>(h);
    Phase& ph = r->tp.getProblem()->phases(1);
    const int ns = ph.nstates, nc = ph.ncontrols, ndir = ns + nc + 1;
    std::vector<adouble> x(ns), u(nc);
    for (int i = 0; i < ns; ++i) { x[i] = states[i]; x[i].d[i] = 1.0; }
    for (int i = 0; i < nc; ++i) { u[i] = controls[i]; u[i].d[ns + i] = 1.0; }
    adouble tt = t;
    tt.d[ns + nc] = 1.0;
    adouble L = ETOL::ePSOPT::integrand_cost(x.data(), u.data(), nullptr, tt, nullptr, 1, &r->ws);
    if (d_cost) for (int k = 0; k < ndir; ++k) d_cost[k] = L.d[k];
    return L.value();
}
double ref_endpoint_cost(void* h, const double* x0, const double* xf, double t0, double tf) {
    RefProblem* r = static_cast<RefProblem*>(h);
    const int ns = r->tp.getProblem()->phases(1).nstates;
    std::vector<adouble> a(ns), b(ns);
    for (int i = 0; i < ns; ++i) { a[i] = x0[i]; b[i] = xf[i]; }
    adouble T0 = t0, TF = tf;
    return ETOL::ePSOPT::endpoint_cost(a.data(), b.data(), nullptr, T0, TF, nullptr, 1, &r->ws).value();
}
// ePSOPT::events: e[nevents]
void ref_events(void* h, const double* x0, const double* xf, double t0, double tf, double* e) {
    RefProblem* r = static_cast<RefProblem*>(h);
    Phase& ph = r->tp.getProblem()->phases(1);
    const int ns = ph.nstates;
    std::vector<adouble> a(ns), b(ns), ev(ph.nevents);
    for (int i = 0; i < ns; ++i) { a[i] = x0[i]; b[i] = xf[i]; }
    adouble T0 = t0, TF = tf;
    ETOL::ePSOPT::events(ev.data(), a.data(), b.data(), nullptr, T0, TF, nullptr, 1, &r->ws);
    for (int i = 0; i < ph.nevents; ++i) e[i] = ev[i].value();
}
// default guess written by ePSOPT::setup: time[nodes] (controls and states are all zero, :47-53)
void ref_guess_time(void* h, double* time) {
    Phase& ph = static_cast<RefProblem*>(h)->tp.getProblem()->phases(1);
    for (long j = 0; j < ph.guess.time.cols(); ++j) time[j] = ph.guess.time(0, j);
}

}  // extern "C"
