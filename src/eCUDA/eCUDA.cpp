// eCUDA.cpp -- the eCUDA eSolver: VGP -> collocation NLP on a B200 through the C ABI of include/ecuda.h.
//
// Mirrors src/ePSOPT/ePSOPT.cpp of the reference method by method:
//   eCUDA::setup        <- ePSOPT::setup      :40-81   dimensions (nevents = 2*nstates, nodes = nsteps+1,
//                                                       npath = #parameters), zero guess + linspace time
//   eCUDA::buildBounds  <- ePSOPT::addBounds  :125-155 box / event / path / fixed-time bounds
//   eCUDA::solve        <- ePSOPT::solve      :83-94   run the NLP solver, score with the sign undone
//   extractTrajectories <- ePSOPT::getTraj    :157-182 one (t, values) pair per collocation node
//   eCUDA::debug/close  <- ePSOPT::debug/close :100-107
// The per-node callbacks ePSOPT::dae / integrand_cost / events (:186-291) are the device kernels.
#include <ETOL/eCUDA.hpp>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>

#include "ecuda_nlp.hpp"

namespace ETOL {

namespace {
std::string paramName(const std::string& name, size_t i, size_t j, size_t k) {
    return name + "_" + std::to_string(i) + "_" + std::to_string(j) + "_" + std::to_string(k);
}
}  // namespace

eCUDA::eCUDA()
    : TrajectoryOptimizer(), _handle(nullptr), _model(ECUDA_MODEL_SI2D), _model_set(false), _obstacles_on(false),
      _tracks_on(false), _is_setup(false), _user_edges(true), _batch(1), _nodes(0) {}

eCUDA::~eCUDA() { close(); }

void eCUDA::fail(const std::string& what) {
    // ETOL's error convention: message, then exit (TrajectoryOptimizer.cpp:1812-1817)
    std::cerr << "eCUDA: " << what;
    if (_handle) std::cerr << " [" << ecuda_last_error(_handle) << "]";
    std::cerr << std::endl;
    exit(EXIT_FAILURE);
}

ecuda_alg_t* eCUDA::getAlgorithm() { return &_algorithm; }
ecuda_prob_t* eCUDA::getProblem() { return &_problem; }
ecuda_sol_t* eCUDA::getSolution() { return &_solution; }
ecuda_handle eCUDA::handle() { return _handle; }
size_t eCUDA::getBatch() const { return _batch; }

void eCUDA::setModel(int model) {
    _model = model;
    _model_set = true;
}

void eCUDA::setBatch(size_t n) { _batch = n ? n : 1; }

// static path rows are one ellipse per polygon edge (si2d, user models without explicit cylinders) or
// vertical cylinders
bool eCUDA::usesEdges(int model) const {
    return model == ECUDA_MODEL_SI2D || (model >= ECUDA_MODEL_USER_BASE && _user_edges);
}
bool eCUDA::isUserModel() const { return _model_set && _model >= ECUDA_MODEL_USER_BASE; }

void eCUDA::addCylinder(double cx, double cy, double radius) {
    addParams({param_t(paramName("cyl", _cylinders.size(), 0, 0),
                       {var_t::CONTINUOUS, -1000., 0., 0., getDt() * getNSteps()})});
    _cylinders.push_back({cx, cy, radius});
}

// same parameters as obsConstraint() of the reference example (etol_psopt_example1.cpp:140-151)
void eCUDA::addObstacleConstraints() {
    const double tspan = getDt() * getNSteps();
    size_t i = 0;
    for (const border_t& border : *getObstacles_Raw()) {
        const int model = _model_set ? _model : (getNStates() == 2 ? ECUDA_MODEL_SI2D : ECUDA_MODEL_PM3D);
        const size_t rows = usesEdges(model) ? border.size() : 1;  // one row per edge / per cylinder
        for (size_t j = 0; j < rows; ++j)
            addParams({param_t(paramName("side", i, j, 0), {var_t::CONTINUOUS, -1000., 0., 0., tspan})});
        ++i;
    }
    _obstacles_on = true;
}

// same parameters as saaConstraint() of the reference example (etol_psopt_example1.cpp:199-223)
void eCUDA::addTrackConstraints() {
    const double tspan = getDt() * getNSteps();
    size_t i = 0;
    for (const track_t& track : *getTracks()) {
        (void)track;
        addParams({param_t(paramName("ball", i, 0, 0), {var_t::CONTINUOUS, -1000., 0., 0., tspan})});
        ++i;
    }
    _tracks_on = true;
}

// problem description for a given choice of device model and constraint sets
void eCUDA::fillDesc(ecuda_problem_desc* out, int model, bool obstacles, bool tracks) {
    ecuda_problem_desc& d = *out;
    d = ecuda_problem_desc{};
    d.model = model;
    d.nphases = 1;  // ePSOPT.cpp:27-28: one phase, no linkages
    d.nnodes[0] = static_cast<int32_t>(_nodes > 0 ? _nodes : getNSteps() + 1);
    size_t nstatic = 0;
    if (obstacles) {
        if (usesEdges(model))
            for (const border_t& b : *getObstacles_Raw()) nstatic += b.size();
        else
            nstatic += getObstacles_Raw()->size();
    }
    if (!usesEdges(model)) nstatic += _cylinders.size();
    d.nstatic[0] = static_cast<int32_t>(nstatic);
    d.ncontrols = static_cast<int32_t>(getNControls());
    d.ntracks = 0;
    d.nwaypoints = 0;
    if (tracks && !getTracks()->empty()) {
        if (model != ECUDA_MODEL_SI2D && model < ECUDA_MODEL_USER_BASE)
            fail("moving exclusion zones are only modelled for the si2d device model and for user models");
        d.ntracks = static_cast<int32_t>(getTracks()->size());
        d.nwaypoints = static_cast<int32_t>(getTracks()->front().trajectory.size());
        for (const track_t& t : *getTracks())
            if (static_cast<int32_t>(t.trajectory.size()) != d.nwaypoints)
                fail("all tracks must have the same number of waypoints");
    }
    d.collocation = _algorithm.collocation_method == "Chebyshev" ? ECUDA_CHEBYSHEV : ECUDA_LEGENDRE;
    d.pattern_mode = ECUDA_PATTERN_DENSE_NODE;
    d.maximize = isMaximized() ? 1 : 0;
    d.batch = static_cast<int32_t>(_batch);
    d.index_base = 0;
}

// ---- callbacks -> device model ------------------------------------------------------------------------------
namespace {
// deterministic sample points (SplitMix64)
struct Rng {
    uint64_t s;
    double uniform() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (z >> 11) * (1.0 / 9007199254740992.0);
    }
};
bool agree(double a, double b) {
    const double d = std::fabs(a - b), m = std::max(std::fabs(a), std::fabs(b));
    return d <= 1e-11 * std::max(m, 1e-3);
}
}  // namespace

// the recording of the objective, the state derivatives and (optionally) constraint rows -> a user model of the library
bool eCUDA::registerTape(const ecuda::Tape& tape, const std::vector<int>& f_ids, int cost_id, const std::vector<int>& row_ids,
                         int32_t* id, std::string* msg) {
    const size_t ns = getNStates(), nc = getNControls();
    int last = cost_id;
    for (int v : f_ids) last = std::max(last, v);
    for (int v : row_ids) last = std::max(last, v);
    std::vector<ecuda_tape_node> nodes(static_cast<size_t>(last) + 1);
    for (int k = 0; k <= last; ++k) {
        const ecuda::Node& n = tape.nodes[k];
        nodes[k] = ecuda_tape_node{static_cast<int32_t>(n.op), n.a, n.b, 0, n.imm};
    }
    ecuda_user_model um{};
    um.nstates = static_cast<int32_t>(ns);
    um.ncontrols = static_cast<int32_t>(nc);
    _user_edges = _cylinders.empty();
    um.static_kind = _user_edges ? ECUDA_STATIC_EDGE : ECUDA_STATIC_CYLINDER;
    um.nnodes = static_cast<int32_t>(nodes.size());
    um.nodes = nodes.data();
    for (size_t i = 0; i < ns && i < ECUDA_MAX_STATES; ++i) um.f_out[i] = f_ids[i];
    um.cost_out = cost_id;
    if (ns > ECUDA_MAX_STATES) {
        *msg = "too many states";
        return false;
    }
    char buf[256] = {0};
    std::vector<int32_t> rows(row_ids.begin(), row_ids.end());
    const int rc = rows.empty() ? ecuda_register_user_model(&um, id, buf, sizeof buf)
                                : ecuda_register_user_model_rows(&um, static_cast<int32_t>(rows.size()), rows.data(), id, buf,
                                                                 sizeof buf);
    if (rc != ECUDA_OK) *msg = buf;
    return rc == ECUDA_OK;
}

bool eCUDA::matchCallbacks(std::string* why) {
    auto no = [&](const std::string& msg) {
        if (why) *why = msg;
        return false;
    };
    const size_t ns = getNStates(), nc = getNControls();
    if (!_objective || _gradient.size() != ns) return no("setObjective and one setGradient entry per state are required");

    // ---- run every callback once on symbolic inputs (the calling convention of ePSOPT::dae, :225-270)
    ecuda::Tape tape;
    ecuda::Tape::active() = &tape;
    std::vector<ecuda::var> xs(ns), us(nc);
    for (size_t i = 0; i < ns; ++i) xs[i] = ecuda::var::input(static_cast<int>(i));
    for (size_t j = 0; j < nc; ++j) us[j] = ecuda::var::input(static_cast<int>(ns + j));
    ecuda::var tv = ecuda::var::input(static_cast<int>(ns + nc));
    vector_t x, u;
    for (size_t i = 0; i < ns; ++i) x.push_back(&xs[i]);
    for (size_t j = 0; j < nc; ++j) u.push_back(&us[j]);
    const vector_t params = {std::string()};
    const std::vector<std::string> pnames = {std::string("")};
    int cost_id = -1;
    std::vector<int> f_ids, row_ids;
    try {
        cost_id = std::any_cast<ecuda::var>((*_objective)(x, u, params, pnames, &tv, getDt())).id();
        for (size_t i = 0; i < ns; ++i)
            f_ids.push_back(std::any_cast<ecuda::var>((*_gradient[i])(x, u, params, pnames, &tv, getDt())).id());
        for (f_t* c : _constraints)
            for (const ecuda::var& r : std::any_cast<fout_ecuda_t>((*c)(x, u, params, pnames, &tv, getDt())))
                row_ids.push_back(r.id());
    } catch (std::bad_any_cast& e) {
        ecuda::Tape::active() = nullptr;
        return no(std::string("a callback did not take/return ecuda::var values (") + e.what() + ")");
    }
    ecuda::Tape::active() = nullptr;

    // ---- sample points inside the state / control box and the time span
    const int npts = 24;
    Rng rng{0xE701u};
    std::vector<std::vector<double>> pts;
    auto pick = [&](const state_t& lo, const state_t& hi, size_t i) {
        const double a = i < lo.size() ? lo[i] : -1.0, b = i < hi.size() ? hi[i] : 1.0;
        return a + rng.uniform() * (b - a);
    };
    for (int p = 0; p < npts; ++p) {
        std::vector<double> in(ns + nc + 1);
        for (size_t i = 0; i < ns; ++i) in[i] = pick(getXlower(), getXupper(), i);
        for (size_t j = 0; j < nc; ++j) in[ns + j] = pick(getUlower(), getUupper(), j);
        in[ns + nc] = rng.uniform() * getNSteps() * getDt();
        pts.push_back(in);
    }
    std::vector<std::vector<double>> vals(npts);
    for (int p = 0; p < npts; ++p) tape.eval(pts[p].data(), &vals[p]);

    // ---- dynamics + running cost against the device models
    int found = -1;
    for (int model : {ECUDA_MODEL_SI2D, ECUDA_MODEL_PM3D, ECUDA_MODEL_FW6}) {
        if (_model_set && model != _model) continue;
        ecuda_problem_desc d{};
        d.model = model;
        d.nphases = 1;
        d.nnodes[0] = static_cast<int32_t>(getNSteps() + 1);
        d.ncontrols = static_cast<int32_t>(nc);
        d.batch = 1;
        ecuda_dims dims{};
        if (ecuda_host_dims(&d, &dims) != ECUDA_OK || static_cast<size_t>(dims.nstates) != ns) continue;
        bool ok = true;
        for (int p = 0; p < npts && ok; ++p) {
            double f[ECUDA_MAX_STATES], cost = 0.0;
            ecuda_host_model_eval(model, pts[p].data(), pts[p].data() + ns, pts[p][ns + nc], f, &cost);
            ok = agree(vals[p][cost_id], cost);
            for (size_t i = 0; i < ns && ok; ++i) ok = agree(vals[p][f_ids[i]], f[i]);
        }
        if (ok) {
            found = model;
            break;
        }
    }
    if (found < 0) {
        // ---- no built-in model computes these callbacks: they become a user model. The recording of the
        // objective and the state derivatives (everything up to the last of those nodes) is handed to the
        // library, which differentiates it and compiles the kernels for it at setup().
        if (_model_set && _model < ECUDA_MODEL_USER_BASE)
            return no("objective / state derivatives do not agree with the device model selected by setModel()");
        int32_t id = -1;
        std::string msg;
        if (!registerTape(tape, f_ids, cost_id, {}, &id, &msg))
            return no("the callbacks match no built-in device model and cannot become a user model: " + msg);
        found = id;
    }

    // ---- constraint rows against the path constraints the VGP data generates (static rows, then tracks). Rows
    // that follow the recognised ones and are none of them become TRACED path rows of a user model: the reference
    // evaluates whatever _constraints holds at every node (ePSOPT.cpp:262-270).
    bool matched = row_ids.empty();
    bool use_obs = false, use_trk = false;
    size_t nbuiltin = 0;
    for (int combo = 3; combo >= 0 && !matched; --combo) {
        const bool obs = combo & 1, trk = combo & 2;
        if (trk && found != ECUDA_MODEL_SI2D && found < ECUDA_MODEL_USER_BASE) continue;
        if ((obs && getObstacles_Raw()->empty() && _cylinders.empty()) || (trk && getTracks()->empty())) continue;
        ecuda_problem_desc d;
        fillDesc(&d, found, obs, trk);
        d.batch = 1;
        ecuda_dims dims{};
        if (ecuda_host_dims(&d, &dims) != ECUDA_OK) continue;
        const size_t nb = static_cast<size_t>(d.nstatic[0] + d.ntracks);
        if (nb > row_ids.size() || (combo == 0 && nb != 0)) continue;
        std::vector<double> inst, rows(nb + 1);
        buildInstanceFor(&inst, found, obs, trk, d, dims.inst_stride);
        bool ok = true;
        for (int p = 0; p < npts && ok && nb > 0; ++p) {
            ecuda_host_path_eval(&d, inst.data(), pts[p][0], pts[p][1], pts[p][ns + nc], rows.data());
            for (size_t q = 0; q < nb && ok; ++q) ok = agree(vals[p][row_ids[q]], rows[q]);
        }
        if (ok) {
            matched = true;
            use_obs = obs;
            use_trk = trk;
            nbuiltin = nb;
        }
    }
    if (!matched)
        return no("the " + std::to_string(row_ids.size()) +
                  " constraint rows are not the exclusion-zone / moving-zone constraints of the loaded VGP");
    _nuser_rows = static_cast<int>(row_ids.size() - nbuiltin);
    if (_nuser_rows > 0) {
        // the model (built-in or recorded) is registered again together with the traced rows
        std::vector<int> extra(row_ids.begin() + static_cast<long>(nbuiltin), row_ids.end());
        int32_t id = -1;
        std::string msg;
        if (!registerTape(tape, f_ids, cost_id, extra, &id, &msg))
            return no(std::to_string(_nuser_rows) + " constraint rows are none of the zone constraints of the loaded VGP "
                      "and cannot become traced path rows: " + msg);
        found = id;
    }
    _model = found;
    _model_set = true;
    _obstacles_on = use_obs;
    _tracks_on = use_trk;
    return true;
}

// ---- setup ------------------------------------------------------------------------------------------------
void eCUDA::setup() {
    _nodes = 0;  // the first mesh: nsteps + 1 nodes (ePSOPT.cpp:44)
    _inst_user = false;  // instance data starts from the loaded VGP again
    transcribe();
    deviceSetup();
}

void eCUDA::deviceSetup() {
    if (_handle) {
        ecuda_destroy(_handle);
        _handle = nullptr;
    }
    if (ecuda_create(_algorithm.device, &_handle) != ECUDA_OK)
        fail(std::string("cannot create the device evaluator: ") + ecuda_last_error(nullptr));
    if (ecuda_set_problem(_handle, &_problem.desc) != ECUDA_OK) fail("ecuda_set_problem");
    if (ecuda_set_scaling(_handle, _problem.sz.data(), _problem.sg.data(), _problem.sf) != ECUDA_OK)
        fail("ecuda_set_scaling");
    if (ecuda_set_ipopt_jac_mode(_handle, _algorithm.derivatives == "numerical" ? ECUDA_JAC_FD_INDEXSET
                                                                                 : ECUDA_JAC_EXACT) != ECUDA_OK)
        fail("ecuda_set_ipopt_jac_mode");
    _is_setup = true;
    uploadInstances();
}

// host half of setup(): VGP -> NLP description (no device needed)
void eCUDA::transcribe() {
    // delayed states / controls: ePSOPT::dae appends x(t - i dt) for 1 <= i < xrhorizon and u(t - i dt) for
    // 1 <= i <= urhorizon to the callback arguments (src/ePSOPT/ePSOPT.cpp:231-248). The device kernels have no
    // delayed terms, so a VGP that asks for them must not be evaluated without them: fail, loudly.
    if (getXrhorizon() >= 2 || getUrhorizon() >= 1)
        fail("delayed states / controls (states rhorizon " + std::to_string(getXrhorizon()) + ", controls rhorizon " +
             std::to_string(getUrhorizon()) + ") are not supported by the eCUDA evaluator: ePSOPT would pass x(t - i dt), "
             "u(t - i dt) to the callbacks (ePSOPT.cpp:231-248)");
    if (_objective || !_gradient.empty() || !_constraints.empty()) {
        std::string why;
        if (!matchCallbacks(&why)) fail("the registered callbacks match no device model: " + why);
    }
    if (!_model_set) _model = getNStates() == 2 ? ECUDA_MODEL_SI2D : ECUDA_MODEL_PM3D;
    const size_t ns = getNStates(), nc = getNControls();
    ecuda_problem_desc& d = _problem.desc;
    fillDesc(&d, _model, _obstacles_on, _tracks_on);
    const size_t nstatic = static_cast<size_t>(d.nstatic[0]);

    if (ecuda_host_dims(&d, &_problem.dims) != ECUDA_OK)
        fail("the VGP does not fit the selected device model (states/controls/nodes)");
    if (static_cast<size_t>(_problem.dims.nstates) != ns)
        fail("device model expects " + std::to_string(_problem.dims.nstates) + " states, the VGP has " +
             std::to_string(ns));
    // npath = #parameters (ePSOPT.cpp:58): every path row must have its parameter
    const size_t npath = nstatic + static_cast<size_t>(d.ntracks) + static_cast<size_t>(_nuser_rows);
    if (npath != getParams()->size())
        fail("path rows (" + std::to_string(npath) + ") and registered parameters (" +
             std::to_string(getParams()->size()) + ") differ");

    const int nnz = _problem.dims.nnz, nv = _problem.dims.nvars;
    _problem.iRow.assign(nnz, 0);
    _problem.jCol.assign(nnz, 0);
    _problem.group_of_col.assign(nv, 0);
    ecuda_host_structure(&d, _problem.iRow.data(), _problem.jCol.data(), _problem.group_of_col.data());
    const int N = d.nnodes[0];
    _problem.tau.assign(N, 0.0);
    _problem.w.assign(N, 0.0);
    std::vector<double> D(static_cast<size_t>(N) * N);
    ecuda_host_collocation(d.collocation, N, _problem.tau.data(), _problem.w.data(), D.data());

    buildBounds();
    buildScaling();

    // guess. ePSOPT starts PSOPT/IPOPT from all-zero states and controls (ePSOPT.cpp:47-53); here the
    // default is the straight line from the initial to the terminal state with the constant control
    // that flies it (dynamics- and event-feasible for the integrator models), clipped into the box.
    // Like ePSOPT's guess it can be overwritten through getProblem()->guess before solve().
    _problem.guess.assign(nv, 0.0);
    {
        const double T = getNSteps() * getDt();
        for (int k = 0; k < N; ++k) {
            const double sfrac = 0.5 * (_problem.tau[k] + 1.0);
            for (size_t i = 0; i < ns; ++i) {
                const double a = i < getX0().size() ? getX0()[i] : 0.0, b = i < getXf().size() ? getXf()[i] : a;
                _problem.guess[nc * N + k * ns + i] = a + sfrac * (b - a);
            }
            if (_model == ECUDA_MODEL_SI2D)
                for (size_t j = 0; j < std::min<size_t>(nc, 2); ++j)
                    _problem.guess[k * nc + j] = (getXf()[j] - getX0()[j]) / T;
        }
        _problem.guess[(ns + nc) * N + 1] = T;
    }
    for (int c = 0; c < nv; ++c) _problem.guess[c] = std::min(std::max(_problem.guess[c], _problem.zl[c]), _problem.zu[c]);

    // per-instance data: every instance starts as a copy of the loaded VGP. The block does not depend on the mesh:
    // when solve() re-transcribes on a refined mesh, data edited through instanceData() is kept (setup() resets it).
    std::vector<double> fresh;
    buildInstance(&fresh);
    const bool keep = _inst_user && _inst.size() == _batch && !_inst.empty() && _inst[0].size() == fresh.size();
    if (!keep) {
        _inst.assign(_batch, fresh);
        _inst_user = false;
    }

}

// ePSOPT::addBounds (ePSOPT.cpp:125-155) in the NLP layout of include/ecuda.h
void eCUDA::buildBounds() {
    const size_t ns = getNStates(), nc = getNControls();
    const int N = _problem.desc.nnodes[0];
    const int nv = _problem.dims.nvars, ng = _problem.dims.ncons;
    const double inf = std::numeric_limits<double>::infinity();
    auto at = [](const state_t& v, size_t i, double dflt) { return i < v.size() ? v[i] : dflt; };
    _problem.zl.assign(nv, -inf);
    _problem.zu.assign(nv, inf);
    for (int k = 0; k < N; ++k) {
        for (size_t j = 0; j < nc; ++j) {
            _problem.zl[k * nc + j] = at(getUlower(), j, -inf);
            _problem.zu[k * nc + j] = at(getUupper(), j, inf);
        }
        for (size_t i = 0; i < ns; ++i) {
            _problem.zl[nc * N + k * ns + i] = at(getXlower(), i, -inf);
            _problem.zu[nc * N + k * ns + i] = at(getXupper(), i, inf);
        }
    }
    const double T = getNSteps() * getDt();
    _problem.zl[(ns + nc) * N] = _problem.zu[(ns + nc) * N] = 0.0;        // StartTime fixed at 0
    _problem.zl[(ns + nc) * N + 1] = _problem.zu[(ns + nc) * N + 1] = T;  // EndTime fixed at nsteps*dt

    _problem.gl.assign(ng, 0.0);  // defect rows: equality with 0
    _problem.gu.assign(ng, 0.0);
    const size_t e0 = ns * N;
    for (size_t i = 0; i < ns; ++i) {
        _problem.gl[e0 + i] = _problem.gu[e0 + i] = at(getX0(), i, 0.0);
        _problem.gl[e0 + ns + i] = at(getXf(), i, 0.0) - at(getXtol(), i, 0.0);
        _problem.gu[e0 + ns + i] = at(getXf(), i, 0.0) + at(getXtol(), i, 0.0);
    }
    // path rows. ePSOPT hands PSOPT the parameter bounds in std::map order while dae() fills the rows in
    // callback order (ePSOPT.cpp:147-150 vs :262-270) -- harmless there because every parameter of the
    // example has the same bounds. Here each row takes the bounds of the parameter it was registered
    // under: static rows side_i_j_0 in obstacle/edge order, then track rows ball_i_0_0.
    _problem.path_names.clear();
    if (_obstacles_on) {
        size_t i = 0;
        for (const border_t& b : *getObstacles_Raw()) {
            const size_t rows = usesEdges(_model) ? b.size() : 1;
            for (size_t j = 0; j < rows; ++j) _problem.path_names.push_back(paramName("side", i, j, 0));
            ++i;
        }
    }
    if (!usesEdges(_model))
        for (size_t c = 0; c < _cylinders.size(); ++c) _problem.path_names.push_back(paramName("cyl", c, 0, 0));
    for (int i = 0; i < _problem.desc.ntracks; ++i) _problem.path_names.push_back(paramName("ball", i, 0, 0));
    if (_nuser_rows > 0) {
        // traced rows: the parameters no zone row was registered under, in the std::map order ePSOPT::addBounds
        // walks (ePSOPT.cpp:147-150)
        std::vector<std::string> rest;
        for (const auto& kv : *getParams())
            if (std::find(_problem.path_names.begin(), _problem.path_names.end(), kv.first) == _problem.path_names.end())
                rest.push_back(kv.first);
        for (int r = 0; r < _nuser_rows; ++r)
            _problem.path_names.push_back(r < static_cast<int>(rest.size()) ? rest[r] : paramName("row", r, 0, 0));
    }
    const size_t np = _problem.path_names.size();
    const size_t p0 = e0 + 2 * ns;
    for (int k = 0; k < N; ++k)
        for (size_t q = 0; q < np; ++q) {
            auto it = getParams()->find(_problem.path_names[q]);
            const double lo = it != getParams()->end() ? it->second.lbnd : -1000.0;
            const double hi = it != getParams()->end() ? it->second.ubnd : 0.0;
            _problem.gl[p0 + k * np + q] = lo;
            _problem.gu[p0 + k * np + q] = hi;
        }
    _problem.gl[ng - 1] = 0.0;  // tf - t0 >= 0
    _problem.gu[ng - 1] = inf;
}

// PSOPT scaling = "automatic" (SURVEY.md appendix A.5): variable scale 1/max(|lb|,|ub|), defect and
// event rows scaled like their state, path rows and the objective left at 1.
void eCUDA::buildScaling() {
    const int nv = _problem.dims.nvars, ng = _problem.dims.ncons;
    const size_t ns = getNStates();
    const int N = _problem.desc.nnodes[0];
    _problem.sz.assign(nv, 1.0);
    _problem.sg.assign(ng, 1.0);
    _problem.sf = 1.0;
    if (_algorithm.scaling != "automatic") return;
    for (int c = 0; c < nv; ++c) {
        const double m = std::max(std::fabs(_problem.zl[c]), std::fabs(_problem.zu[c]));
        if (std::isfinite(m) && m > 0.0) _problem.sz[c] = 1.0 / m;
    }
    const size_t x0col = getNControls() * N;
    for (int k = 0; k < N; ++k)
        for (size_t i = 0; i < ns; ++i) _problem.sg[k * ns + i] = _problem.sz[x0col + i];
    for (size_t i = 0; i < ns; ++i) {
        _problem.sg[ns * N + i] = _problem.sz[x0col + i];
        _problem.sg[ns * N + ns + i] = _problem.sz[x0col + i];
    }
}

// obstacle / track records of the loaded VGP in the layout of ecuda_upload_instances
void eCUDA::buildInstance(std::vector<double>* out) const {
    buildInstanceFor(out, _model, _obstacles_on, _tracks_on, _problem.desc, _problem.dims.inst_stride);
}

void eCUDA::buildInstanceFor(std::vector<double>* out, int model, bool obstacles, bool tracks,
                             const ecuda_problem_desc& d, int inst_stride) const {
    (void)tracks;
    out->assign(inst_stride, 0.0);
    size_t o = 0;
    eCUDA* self = const_cast<eCUDA*>(this);
    if (obstacles) {
        for (const border_t& border : *self->getObstacles_Raw()) {
            std::vector<double> xy;
            for (const corner_t& c : border) {
                xy.push_back(c[0]);
                xy.push_back(c[1]);
            }
            const int n = static_cast<int>(border.size());
            if (usesEdges(model)) {  // one ellipse per polygon edge (etol_psopt_example1.cpp:164-179)
                ecuda_si2d_edge_records(xy.data(), n, out->data() + o);
                o += 6 * n;
            } else {  // circumscribed vertical cylinder: centroid + farthest corner
                double cx = 0, cy = 0, r2 = 0;
                for (int i = 0; i < n; ++i) {
                    cx += xy[2 * i];
                    cy += xy[2 * i + 1];
                }
                cx /= n;
                cy /= n;
                for (int i = 0; i < n; ++i)
                    r2 = std::max(r2, (xy[2 * i] - cx) * (xy[2 * i] - cx) + (xy[2 * i + 1] - cy) * (xy[2 * i + 1] - cy));
                (*out)[o] = cx;
                (*out)[o + 1] = cy;
                (*out)[o + 2] = r2;
                (*out)[o + 3] = 0.0;
                o += 4;
            }
        }
    }
    if (!usesEdges(model))
        for (const auto& c : _cylinders) {
            (*out)[o] = c[0];
            (*out)[o + 1] = c[1];
            (*out)[o + 2] = c[2] * c[2];
            (*out)[o + 3] = 0.0;
            o += 4;
        }
    if (d.ntracks > 0)
        for (const track_t& t : *self->getTracks()) {
            (*out)[o++] = t.radius;
            for (const traj_elem_t& wp : t.trajectory) {
                (*out)[o++] = wp.first;
                (*out)[o++] = wp.second.size() > 0 ? wp.second[0] : 0.0;
                (*out)[o++] = wp.second.size() > 1 ? wp.second[1] : 0.0;
            }
        }
}

std::vector<double>& eCUDA::instanceData(size_t b) {
    _inst_user = true;  // the caller may edit it: keep it when solve() re-meshes
    return _inst.at(b);
}

void eCUDA::uploadInstances() {
    if (!_is_setup) fail("uploadInstances() before setup()");
    const size_t stride = _problem.dims.inst_stride;
    std::vector<double> all(_batch * stride, 0.0);
    for (size_t b = 0; b < _batch; ++b) std::copy(_inst[b].begin(), _inst[b].end(), all.begin() + b * stride);
    if (ecuda_upload_instances(_handle, all.data(), ECUDA_MEM_HOST) != ECUDA_OK) fail("ecuda_upload_instances");
}

// ---- evaluation ---------------------------------------------------------------------------------------------
int eCUDA::evaluate(const double* z, double* f, double* g, double* jac) {
    if (!_is_setup) fail("evaluate() before setup()");
    const size_t nv = _problem.dims.nvars;
    _zscaled.resize(_batch * nv);
    for (size_t b = 0; b < _batch; ++b)
        for (size_t c = 0; c < nv; ++c) _zscaled[b * nv + c] = z[b * nv + c] * _problem.sz[c];
    const int mode = _algorithm.derivatives == "numerical" ? ECUDA_JAC_FD_INDEXSET : ECUDA_JAC_EXACT;
    return ecuda_eval(_handle, _zscaled.data(), f, g, jac, mode, ECUDA_MEM_HOST, nullptr);
}

int eCUDA::evaluateGradient(const double* z, double* grad) {
    if (!_is_setup) fail("evaluateGradient() before setup()");
    const size_t nv = _problem.dims.nvars;
    _zscaled.resize(_batch * nv);
    for (size_t b = 0; b < _batch; ++b)
        for (size_t c = 0; c < nv; ++c) _zscaled[b * nv + c] = z[b * nv + c] * _problem.sz[c];
    return ecuda_eval_grad_f(_handle, _zscaled.data(), grad, ECUDA_MEM_HOST, nullptr);
}

// ---- solve ----------------------------------------------------------------------------------------------------
// Automatic mesh refinement: what PSOPT does around its NLP solves when ePSOPT sets mesh_refinement =
// "automatic" (ePSOPT.cpp:69-71). PSOPT's own rule is not in the reference tree; this one is eCUDA's: the
// first refinement adds mr_initial_increment nodes; later ones extrapolate the straight line through
// (nodes, log10 error) of the last two solves to the tolerance, and add at least 2 nodes and at most
// mr_max_increment_factor * nodes.
int eCUDA::nextMeshSize(const std::vector<std::pair<int, double>>& hist, const ecuda_alg_t& alg) {
    const int N = hist.back().first;
    if (hist.size() < 2) return N + std::max(2, alg.mr_initial_increment);
    const int N0 = hist[hist.size() - 2].first;
    const double e0 = hist[hist.size() - 2].second, e1 = hist.back().second;
    const int cap = N + std::max(2, static_cast<int>(std::ceil(alg.mr_max_increment_factor * N)));
    if (!(e0 > 0.0) || !(e1 > 0.0) || e1 >= e0 || N == N0) return cap;  // no measurable decay: the full step
    const double slope = (std::log10(e1) - std::log10(e0)) / (N - N0);   // decades per node, negative
    const double need = N + (std::log10(alg.ode_tolerance) - std::log10(e1)) / slope;
    const int want = static_cast<int>(std::ceil(need));
    return std::min(cap, std::max(N + 2, want));
}

void eCUDA::solve() {
    if (!_is_setup) fail("solve() before setup()");
    if (_batch != 1) fail("solve() drives one instance; batches are evaluated with evaluate()");
    _solution.mesh_history.clear();
    ecuda_sol_t best;
    bool have_best = false;
    for (int it = 0;; ++it) {
        const int rc = solveOnce();
        if (rc != 0) {
            if (!have_best) return;  // first mesh failed: as ePSOPT::solve (:85-87)
            std::cout << "eCUDA: the solve on " << _problem.desc.nnodes[0]
                      << " nodes failed; keeping the solution of the previous mesh" << std::endl;
            const std::vector<std::pair<int, double>> hist = _solution.mesh_history;
            _solution = best;
            _solution.mesh_history = hist;
            _nodes = hist.back().first;  // back to the mesh that solution lives on
            transcribe();
            deviceSetup();
            break;
        }
        // relative local discretisation error of this solution, per mesh interval (device)
        const int N = _problem.desc.nnodes[0], nv = _problem.dims.nvars;
        std::vector<double> zs(nv), err(N - 1, 0.0);
        for (int c = 0; c < nv; ++c) zs[c] = _solution.z[c] * _problem.sz[c];
        if (ecuda_ode_error(_handle, zs.data(), err.data(), ECUDA_MEM_HOST, nullptr) != ECUDA_OK) fail("ecuda_ode_error");
        double emax = 0.0;
        for (double e : err) emax = std::max(emax, e);
        _solution.mesh_history.push_back({N, emax});
        if (_algorithm.print_level > 0)
            std::cout << "mesh " << it << ": " << N << " nodes, max relative local error " << emax << std::endl;
        if (_algorithm.mesh_refinement != "automatic" || emax <= _algorithm.ode_tolerance ||
            it >= _algorithm.mr_max_iterations)
            break;
        // next mesh: interpolate the solution onto it as the starting point, re-transcribe, re-create the evaluator
        best = _solution;
        have_best = true;
        const int32_t Nn = nextMeshSize(_solution.mesh_history, _algorithm);
        std::vector<double> znew((getNStates() + getNControls()) * static_cast<size_t>(Nn) + 2);
        if (ecuda_resample(_handle, zs.data(), &Nn, nullptr, znew.data(), ECUDA_MEM_HOST, nullptr) != ECUDA_OK)
            fail("ecuda_resample");
        _nodes = Nn;
        const std::vector<std::pair<int, double>> hist = _solution.mesh_history;
        transcribe();
        deviceSetup();
        _problem.guess = znew;
        for (int c = 0; c < _problem.dims.nvars; ++c)
            _problem.guess[c] = std::min(std::max(_problem.guess[c], _problem.zl[c]), _problem.zu[c]);
        _solution.mesh_history = hist;
    }
    if (_solution.error_flag == 0) {
        setScore(isMaximized() ? -_solution.cost : _solution.cost);
        extractTrajectories(_solution.z);
    }
}

int eCUDA::solveOnce() {
    const int nv = _problem.dims.nvars, ng = _problem.dims.ncons, nnz = _problem.dims.nnz;
    const int mode = _algorithm.derivatives == "numerical" ? ECUDA_JAC_FD_INDEXSET : ECUDA_JAC_EXACT;

    // the solver works in the scaled space the device evaluates in (like IPOPT under PSOPT)
    ecuda_nlp::Problem P;
    P.n = nv;
    P.m = ng;
    P.nnz = nnz;
    P.irow = _problem.iRow.data();
    P.jcol = _problem.jCol.data();
    P.zl.resize(nv);
    P.zu.resize(nv);
    P.gl.resize(ng);
    P.gu.resize(ng);
    for (int c = 0; c < nv; ++c) {
        P.zl[c] = _problem.zl[c] * _problem.sz[c];
        P.zu[c] = _problem.zu[c] * _problem.sz[c];
    }
    for (int r = 0; r < ng; ++r) {
        P.gl[r] = _problem.gl[r] * _problem.sg[r];
        P.gu[r] = _problem.gu[r] * _problem.sg[r];
    }
    ecuda_handle h = _handle;
    P.eval = [h, mode](const double* zs, double* f, double* g, double* jac, double* grad) -> bool {
        if ((f || g || jac) && ecuda_eval(h, zs, f, g, jac, mode, ECUDA_MEM_HOST, nullptr) != ECUDA_OK) return false;
        if (grad && ecuda_eval_grad_f(h, zs, grad, ECUDA_MEM_HOST, nullptr) != ECUDA_OK) return false;
        return true;
    };
    ecuda_nlp::Options opt;
    opt.max_iter = _algorithm.nlp_iter_max;
    opt.tol = _algorithm.nlp_tolerance;
    opt.print_level = _algorithm.print_level;
    std::vector<double> z(nv);
    for (int c = 0; c < nv; ++c) z[c] = _problem.guess[c] * _problem.sz[c];
    const bool want_ipopt = _algorithm.nlp_method == "IPOPT";
    const bool use_ipopt = want_ipopt && ecuda_nlp::have_ipopt();
    std::vector<int32_t> hrow, hcol;
    if (_algorithm.hessian == "exact") {
        // The exact Lagrangian Hessian (ecuda_eval_hess) is what IPOPT gets through eval_h. The built-in driver does
        // not use it (it keeps a damped-BFGS model): say so instead of silently ignoring the setting. Should the
        // device Hessian be unavailable for a model, IPOPT runs with its limited-memory approximation.
        int32_t hn = 0;
        if (ecuda_get_hess_structure(h, &hn, nullptr, nullptr) != ECUDA_OK) fail("ecuda_get_hess_structure");
        std::vector<double> probe(static_cast<size_t>(hn)), lam0(static_cast<size_t>(ng), 0.0);
        const bool available = ecuda_eval_hess(h, z.data(), nullptr, 1.0, lam0.data(), probe.data(), ECUDA_MEM_HOST, nullptr) == ECUDA_OK;
        if (!use_ipopt) {
            if (_algorithm.print_level > 0)
                std::cout << "eCUDA: hessian = \"exact\" is used by the IPOPT adapter only; the built-in NLP driver keeps its "
                             "quasi-Newton model" << std::endl;
        } else if (!available) {
            std::cout << "eCUDA: no exact Hessian for this model (" << ecuda_last_error(h)
                      << "); IPOPT runs with hessian_approximation = limited-memory" << std::endl;
        } else {
            hrow.resize(hn);
            hcol.resize(hn);
            ecuda_get_hess_structure(h, nullptr, hrow.data(), hcol.data());
            P.hnnz = hn;
            P.hrow = hrow.data();
            P.hcol = hcol.data();
            P.eval_h = [h](const double* zs, double sigma, const double* lambda, double* hv) -> bool {
                return ecuda_eval_hess(h, zs, nullptr, sigma, lambda, hv, ECUDA_MEM_HOST, nullptr) == ECUDA_OK;
            };
        }
    }
    ecuda_nlp::Result R;
    int rc;
    if (use_ipopt)
        rc = ecuda_nlp::solve_ipopt(P, opt, &z, &R);
    else
        rc = ecuda_nlp::solve_builtin(P, opt, &z, &R);

    _solution.error_flag = rc;
    _solution.error_msg = R.message;
    _solution.nlp_iterations = R.iterations;
    _solution.max_violation = R.max_violation;
    if (rc != 0) {  // as ePSOPT::solve (:85-87): report, leave score and trajectories untouched
        std::cout << "!!!!!Problem failed!!!!!" << std::endl << _solution.error_msg << std::endl;
        return rc;
    }
    _solution.z.resize(nv);
    for (int c = 0; c < nv; ++c) _solution.z[c] = z[c] / _problem.sz[c];
    _solution.cost = R.objective / _problem.sf;
    return 0;
}

void eCUDA::extractTrajectories(const std::vector<double>& z) {
    const size_t ns = getNStates(), nc = getNControls();
    const int N = _problem.desc.nnodes[0];
    const double t0 = z[(ns + nc) * N], tf = z[(ns + nc) * N + 1];
    traj_t* xt = getXtraj();
    traj_t* ut = getUtraj();
    xt->clear();
    ut->clear();
    for (int k = 0; k < N; ++k) {
        const double t = 0.5 * (tf - t0) * _problem.tau[k] + 0.5 * (tf + t0);
        state_t x(z.begin() + nc * N + k * ns, z.begin() + nc * N + (k + 1) * ns);
        state_t u(z.begin() + k * nc, z.begin() + (k + 1) * nc);
        xt->push_back(traj_elem_t(t, x));
        ut->push_back(traj_elem_t(t, u));
    }
}

void eCUDA::debug() { _algorithm.print_level = 5; }

void eCUDA::close() {
    if (_handle) {
        ecuda_destroy(_handle);
        _handle = nullptr;
    }
    _is_setup = false;
}

}  // namespace ETOL
