// shim.cpp -- TEST-ONLY C entry points over the C++ plugin layer so that the python test-suite can
// drive it with ctypes: XML -> VGP -> transcription (no GPU needed), CSV/XML writers, the built-in
// NLP driver with a caller-supplied evaluation callback, and the full eCUDA setup/evaluate/solve
// path (GPU needed). Nothing in the product links this file.
#include <ETOL/eCUDA.hpp>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ecuda_nlp.hpp"
#include "vgp_si2d_callbacks.hpp"

using ETOL::eCUDA;

extern "C" {

void* shim_create() { return new eCUDA(); }
void shim_destroy(void* h) { delete static_cast<eCUDA*>(h); }

// load + register constraints + transcribe (host only). flags: bit0 obstacles, bit1 tracks
int shim_load(void* h, const char* xml, int model, int flags, int batch, const char* scaling, const char* derivatives) {
    eCUDA* t = static_cast<eCUDA*>(h);
    t->loadConfigs(xml);
    t->setMaximize(false);
    if (model >= 0) t->setModel(model);
    if (flags & 1) t->addObstacleConstraints();
    if (flags & 2) t->addTrackConstraints();
    t->setBatch(batch);
    if (scaling) t->getAlgorithm()->scaling = scaling;
    if (derivatives) t->getAlgorithm()->derivatives = derivatives;
    t->transcribe();
    return 0;
}
// load + register USER CALLBACKS (ecuda::var) like the reference example does, then try to match them.
// variant 0: the example's callbacks; 1: a different objective; 2: exclusion zones only; 3: zones
// registered in the opposite order (moving zones first); 4: dynamics no built-in model has (wind field
// depending on the position) -> user model; 5: an objective holding a static ecuda::var constant; 6: dynamics that
// read the node time -> time-dependent user model; 7: a third constraint callback that is none of the zone
// constraints (a disc that grows with time) -> traced path row of a user model; 8: a constraint row that reads a
// control -> refused. Returns 1 when matched; *model, *flags
// (bit0 obstacles, bit1 tracks) report what was recognised, why (<= 255 chars) the reason otherwise.
int shim_load_callbacks(void* h, const char* xml, int variant, int* model, int* flags, char* why) {
    eCUDA* t = static_cast<eCUDA*>(h);
    t->loadConfigs(xml);
    t->setMaximize(false);
    static std::vector<std::unique_ptr<ETOL::f_t>> keep;  // callbacks must outlive the optimizer calls
    auto hold = [&](ETOL::f_t f) {
        keep.push_back(std::make_unique<ETOL::f_t>(std::move(f)));
        return keep.back().get();
    };
    ETOL::f_t other = [](F_ARGS) -> ETOL::scalar_t {
        ecuda::var a = vgp_si2d::at(u, 0), b = vgp_si2d::at(u, 1);
        return a * a + 2.0 * (b * b);
    };
    // variant 5: the example's objective written with a constant that OUTLIVES one recording (static ecuda::var):
    // every transcription records the callback on a fresh tape, the constant must be materialised again on each
    ETOL::f_t kept_constant = [](F_ARGS) -> ETOL::scalar_t {
        static const ecuda::var one(1.0);
        ecuda::var a = vgp_si2d::at(u, 0), b = vgp_si2d::at(u, 1);
        return one * (a * a) + one * (b * b);
    };
    t->setObjective(hold(variant == 1 ? other : variant == 5 ? kept_constant : ETOL::f_t(&vgp_si2d::effort)));
    if (variant == 4)
        t->setGradient({hold(&vgp_si2d::windyXdot), hold(&vgp_si2d::windyYdot)});
    else if (variant == 6)  // dynamics that read the node time
        t->setGradient({hold(&vgp_si2d::gustXdot), hold(&vgp_si2d::gustYdot)});
    else
        t->setGradient({hold(&vgp_si2d::xdot), hold(&vgp_si2d::ydot)});
    ETOL::f_t* zones = hold(vgp_si2d::exclusionZones(t));
    if (variant == 2) {
        t->setConstraints({zones});
    } else {
        ETOL::f_t* movers = hold(vgp_si2d::movingZones(t));
        if (variant == 3)
            t->setConstraints({movers, zones});
        else if (variant == 8) {  // a constraint row that reads a control: cannot be a traced path row
            t->addParams({ETOL::param_t("ctl_0_0_0", {ETOL::var_t::CONTINUOUS, -1., 1., 0., t->getDt() * t->getNSteps()})});
            t->setConstraints({zones, movers, hold([](F_ARGS) -> ETOL::scalar_t {
                                   return ETOL::fout_ecuda_t{vgp_si2d::at(u, 0) * vgp_si2d::at(x, 0)};
                               })});
        } else if (variant == 7)  // one more constraint that is none of the VGP's zones: a traced path row
            t->setConstraints({zones, movers, hold(vgp_si2d::growingDisc(t, 3.0, 3.5, 0.2, 0.01))});
        else
            t->setConstraints({zones, movers});
    }
    std::string reason;
    const bool ok = t->matchCallbacks(&reason);
    std::strncpy(why, reason.c_str(), 255);
    why[255] = 0;
    if (ok) {
        t->transcribe();
        *model = t->getProblem()->desc.model;
        *flags = (t->getProblem()->desc.nstatic[0] > 0 ? 1 : 0) | (t->getProblem()->desc.ntracks > 0 ? 2 : 0);
    }
    return ok ? 1 : 0;
}
void shim_vgp(void* h, int* out /*nsteps,nstates,ncontrols,nzones,ntracks,nparams*/, double* dt) {
    eCUDA* t = static_cast<eCUDA*>(h);
    out[0] = (int)t->getNSteps(); out[1] = (int)t->getNStates(); out[2] = (int)t->getNControls();
    out[3] = (int)t->getNExclZones(); out[4] = (int)t->getNTracks(); out[5] = (int)t->getParams()->size();
    *dt = t->getDt();
}
void shim_dims(void* h, ecuda_dims* d) { *d = static_cast<eCUDA*>(h)->getProblem()->dims; }
void shim_desc(void* h, ecuda_problem_desc* d) { *d = static_cast<eCUDA*>(h)->getProblem()->desc; }
void shim_bounds(void* h, double* zl, double* zu, double* gl, double* gu, double* guess, double* sz, double* sg) {
    ETOL::ecuda_prob_t* p = static_cast<eCUDA*>(h)->getProblem();
    std::memcpy(zl, p->zl.data(), sizeof(double) * p->zl.size());
    std::memcpy(zu, p->zu.data(), sizeof(double) * p->zu.size());
    std::memcpy(gl, p->gl.data(), sizeof(double) * p->gl.size());
    std::memcpy(gu, p->gu.data(), sizeof(double) * p->gu.size());
    std::memcpy(guess, p->guess.data(), sizeof(double) * p->guess.size());
    std::memcpy(sz, p->sz.data(), sizeof(double) * p->sz.size());
    std::memcpy(sg, p->sg.data(), sizeof(double) * p->sg.size());
}
void shim_instance(void* h, int b, double* out) {
    std::vector<double>& v = static_cast<eCUDA*>(h)->instanceData(b);
    std::memcpy(out, v.data(), sizeof(double) * v.size());
}
// edit one value of instance b's data block through instanceData(), then transcribe again the way
// solve() does when it refines the mesh (more nodes); returns the value found afterwards. again_setup != 0 re-runs
// transcription the way setup() does instead (starts from the loaded VGP again).
double shim_edit_instance_and_remesh(void* h, int b, int index, double value, int more_nodes, int as_setup) {
    eCUDA* t = static_cast<eCUDA*>(h);
    t->instanceData(static_cast<size_t>(b)).at(static_cast<size_t>(index)) = value;
    if (as_setup)
        t->retranscribeForTest(0, true);
    else
        t->retranscribeForTest(t->getProblem()->desc.nnodes[0] + more_nodes, false);
    const std::vector<std::vector<double>>& inst = t->instanceBlocksForTest();
    return inst.at(static_cast<size_t>(b)).at(static_cast<size_t>(index));
}
// convex partition of a polygon (TrajectoryOptimizer::genRegion + calcSlopes): returns the number of pieces; out holds,
// per piece, [nlower, nupper, lower xy..., upper xy..., lower slopes (nlower-1), upper slopes (nupper-1)]
int shim_gen_region(const double* xy, int n, double* out, int cap) {
    ETOL::border_t border;
    for (int i = 0; i < n; ++i) border.push_back({xy[2 * i], xy[2 * i + 1], 0.});
    ETOL::region_t region = ETOL::TrajectoryOptimizer::genRegion(&border);
    std::vector<ETOL::seg_t> lowers, uppers;
    ETOL::TrajectoryOptimizer::calcSlopes(region, &lowers, &uppers);
    int w = 0, piece = 0;
    auto put = [&](double v) {
        if (w < cap) out[w] = v;
        ++w;
    };
    for (const ETOL::boundary_t& bd : region) {
        put(static_cast<double>(bd.lower.size()));
        put(static_cast<double>(bd.upper.size()));
        for (const ETOL::corner_t& c : bd.lower) { put(c[0]); put(c[1]); }
        for (const ETOL::corner_t& c : bd.upper) { put(c[0]); put(c[1]); }
        for (const ETOL::edge_t& e : lowers[piece]) put(e.second.slope);
        for (const ETOL::edge_t& e : uppers[piece]) put(e.second.slope);
        ++piece;
    }
    return w <= cap ? piece : -w;
}
// number of partitioned exclusion zones the loaded VGP holds (addExclZone fills them while loading)
int shim_num_partitioned_zones(void* h) { return static_cast<int>(static_cast<eCUDA*>(h)->getObstacles()->size()); }
void shim_structure(void* h, int32_t* irow, int32_t* jcol, int32_t* grp) {
    ETOL::ecuda_prob_t* p = static_cast<eCUDA*>(h)->getProblem();
    std::memcpy(irow, p->iRow.data(), sizeof(int32_t) * p->iRow.size());
    std::memcpy(jcol, p->jCol.data(), sizeof(int32_t) * p->jCol.size());
    std::memcpy(grp, p->group_of_col.data(), sizeof(int32_t) * p->group_of_col.size());
}
int shim_save_xml(void* h, const char* path) {
    static_cast<eCUDA*>(h)->saveConfigs(path);
    return 0;
}
// CSV writer on a synthetic trajectory: n rows (t = i*0.5, values i + 0.25*c)
int shim_save_csv(const char* path, int n, int width, char* out_path, int out_len) {
    ETOL::traj_t traj;
    for (int i = 0; i < n; ++i) {
        ETOL::state_t v;
        for (int c = 0; c < width; ++c) v.push_back(i + 0.25 * c);
        traj.push_back(ETOL::traj_elem_t(0.5 * i, v));
    }
    std::string p = ETOL::TrajectoryOptimizer::save(&traj, path);
    std::strncpy(out_path, p.c_str(), out_len - 1);
    out_path[out_len - 1] = 0;
    return 0;
}
double shim_interp(double t, int n, const double* tv, const double* ref) {
    ETOL::state_t a(tv, tv + n), b(ref, ref + n);
    return ETOL::TrajectoryOptimizer::linear_interpolation<double>(t, a, b);
}

// ---- built-in NLP driver with a python callback -------------------------------------------------------
typedef int (*eval_cb)(const double* z, double* f, double* g, double* jac, double* grad);
int shim_nlp_solve(int n, int m, int nnz, const double* zl, const double* zu, const double* gl, const double* gu,
                   const int32_t* irow, const int32_t* jcol, eval_cb cb, int max_iter, double tol, int print_level,
                   double* z, double* result /*iters, objective, max_violation*/) {
    ecuda_nlp::Problem P;
    P.n = n; P.m = m; P.nnz = nnz;
    P.zl.assign(zl, zl + n); P.zu.assign(zu, zu + n); P.gl.assign(gl, gl + m); P.gu.assign(gu, gu + m);
    P.irow = irow; P.jcol = jcol;
    P.eval = [cb](const double* zz, double* f, double* g, double* jac, double* grad) { return cb(zz, f, g, jac, grad) == 0; };
    ecuda_nlp::Options opt;
    opt.max_iter = max_iter; opt.tol = tol; opt.print_level = print_level;
    std::vector<double> zz(z, z + n);
    ecuda_nlp::Result R;
    int rc = ecuda_nlp::solve_builtin(P, opt, &zz, &R);
    std::memcpy(z, zz.data(), sizeof(double) * n);
    result[0] = R.iterations; result[1] = R.objective; result[2] = R.max_violation;
    return rc;
}

// ---- the device path (GPU needed) --------------------------------------------------------------------------
int shim_setup(void* h) { static_cast<eCUDA*>(h)->setup(); return 0; }
int shim_evaluate(void* h, const double* z, double* f, double* g, double* jac) {
    return static_cast<eCUDA*>(h)->evaluate(z, f, g, jac);
}
int shim_solve(void* h, int max_iter, int print_level, double* score, int* iters, double* viol) {
    eCUDA* t = static_cast<eCUDA*>(h);
    t->getAlgorithm()->nlp_iter_max = max_iter;
    t->getAlgorithm()->print_level = print_level;
    t->solve();
    *score = t->getScore();
    *iters = t->getSolution()->nlp_iterations;
    *viol = t->getSolution()->max_violation;
    return t->getSolution()->error_flag;
}
void shim_set_hessian(void* h, const char* mode) { static_cast<eCUDA*>(h)->getAlgorithm()->hessian = mode; }
// mesh refinement controls / report (mode "automatic" | "manual")
void shim_set_mesh(void* h, const char* mode, double ode_tolerance, int max_iterations) {
    eCUDA* t = static_cast<eCUDA*>(h);
    t->getAlgorithm()->mesh_refinement = mode;
    t->getAlgorithm()->ode_tolerance = ode_tolerance;
    t->getAlgorithm()->mr_max_iterations = max_iterations;
}
int shim_mesh_history(void* h, int cap, int* nodes, double* err) {
    const auto& hist = static_cast<eCUDA*>(h)->getSolution()->mesh_history;
    for (size_t i = 0; i < hist.size() && static_cast<int>(i) < cap; ++i) {
        nodes[i] = hist[i].first;
        err[i] = hist[i].second;
    }
    return static_cast<int>(hist.size());
}
int shim_next_mesh_size(int n, const int* nodes, const double* err, double tol, int initial_increment, double factor) {
    std::vector<std::pair<int, double>> hist;
    for (int i = 0; i < n; ++i) hist.push_back({nodes[i], err[i]});
    ETOL::ecuda_alg_t alg;
    alg.ode_tolerance = tol;
    alg.mr_initial_increment = initial_increment;
    alg.mr_max_increment_factor = factor;
    return eCUDA::nextMeshSize(hist, alg);
}
int shim_traj(void* h, int which, double* out /* [N][1+width] */) {
    eCUDA* t = static_cast<eCUDA*>(h);
    ETOL::traj_t* tr = which == 0 ? t->getXtraj() : t->getUtraj();
    size_t o = 0;
    for (auto& e : *tr) {
        out[o++] = e.first;
        for (double v : e.second) out[o++] = v;
    }
    return (int)tr->size();
}
void shim_close(void* h) { static_cast<eCUDA*>(h)->close(); }

}  // extern "C"
