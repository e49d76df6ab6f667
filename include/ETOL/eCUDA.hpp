// eCUDA.hpp -- the eCUDA eSolver: ETOL's TrajectoryOptimizer interface in front of the B200
// collocation-NLP evaluator (include/ecuda.h).
//
// It takes the place ePSOPT has in the reference (include/ETOL/ePSOPT.hpp, src/ePSOPT/ePSOPT.cpp):
// the same XML, the same call sequence
//     loadConfigs -> setMaximize -> [model + constraint registration] -> setup -> debug -> solve
//     -> getScore / getXtraj / getUtraj / save -> close
// (src/Examples/PSOPT/etol_psopt_example1.cpp:41-81), the same NLP dimensions and bounds
// (ePSOPT::setup / addBounds, src/ePSOPT/ePSOPT.cpp:40-81,125-155). What differs is where the VGP
// callbacks run. ePSOPT calls host std::function lambdas on adouble once per node
// (ePSOPT.cpp:186-276); a GPU cannot call those, so -- like every eSolver, which fixes the scalar
// type its callbacks are written in (src/docs/source/tutorials/vgp.rst:155) -- eCUDA takes its
// callbacks as *device models*: setModel() selects dynamics + running cost, and
// addObstacleConstraints() / addTrackConstraints() register the path constraints that the
// reference example builds in obsConstraint() / saaConstraint()
// (etol_psopt_example1.cpp:140-258), including the same "side_i_j_0" / "ball_i_0_0" parameters.
#ifndef INCLUDE_ETOL_ECUDA_HPP_
#define INCLUDE_ETOL_ECUDA_HPP_

#include <array>
#include <string>
#include <vector>

#include <ETOL/TrajectoryOptimizer.hpp>
#include <ETOL/eCUDA_Types.hpp>
#include <ETOL/eCUDA_var.hpp>

namespace ETOL {

class eCUDA : public TrajectoryOptimizer {
 public:
    eCUDA();
    virtual ~eCUDA();

    // ---- the eSolver interface (TrajectoryOptimizer.hpp:39-54) -----------------------------------
    void setup();  // transcribe() + creates the device evaluator and uploads the problem to the GPU
    void solve();  // runs the NLP solver with GPU callbacks; on success setScore() + trajectories
    void debug();  // print_level = 5 (call after setup, before solve -- as ePSOPT::debug)
    void close();  // releases the device handle

    // host half of setup(): VGP -> NLP dimensions, bounds, scaling, guess, Jacobian structure and
    // instance data in getProblem(); needs no device
    void transcribe();

    // ---- solver structs, as ePSOPT::getAlgorithm/getProblem/getSolution ---------------------------
    ecuda_alg_t* getAlgorithm();
    ecuda_prob_t* getProblem();
    ecuda_sol_t* getSolution();

    // ---- VGP callbacks written for eCUDA -----------------------------------------------------------
    // setObjective / setGradient / setConstraints work as for ePSOPT, with ecuda::var in place of
    // adouble: x, u hold `ecuda::var*`, k is `ecuda::var*` (time), dt a double; objective and state
    // derivatives return ecuda::var, constraint callbacks return fout_ecuda_t (eCUDA_var.hpp). At
    // setup() every callback is run once on symbolic inputs; the recorded expressions are evaluated at
    // sample points of the state/control box and compared with the device models and with the path
    // constraints the VGP data generates. The match selects what the GPU runs -- the callbacks
    // themselves never run on the hot path. When objective and state derivatives are none of the
    // built-in device models, their recording is registered as a user model
    // (ecuda_register_user_model): the library differentiates it and compiles the evaluation kernels
    // for it at setup(), so any autonomous dynamics / running cost written with ecuda::var runs on the
    // GPU. The constraint callbacks must still be the exclusion-zone / moving-zone constraints of the
    // VGP data. False (with a reason) when nothing matches; setup() then fails like ePSOPT does on a
    // bad callback. Called by transcribe() when callbacks are registered.
    bool matchCallbacks(std::string* why = nullptr);
    bool registerTape(const ecuda::Tape& tape, const std::vector<int>& f_ids, int cost_id, const std::vector<int>& row_ids,
                      int32_t* id, std::string* msg);
    // true when setup() turned the callbacks into a user model (kernels compiled for them at run time)
    bool isUserModel() const;

    // ---- device-side VGP callbacks, selected directly ------------------------------------------------
    // dynamics + running cost: ECUDA_MODEL_SI2D (the reference example's x'=u0, y'=u1, u0^2+u1^2),
    // ECUDA_MODEL_PM3D, ECUDA_MODEL_FW6
    void setModel(int model);
    // one ellipse constraint per polygon edge of every exclusion zone (si2d) or one circumscribed
    // vertical cylinder per exclusion zone (pm3d, fw6); adds parameters side_i_j_0 = {C,-1000,0,0,T}
    void addObstacleConstraints();
    // one keep-out circle per moving exclusion zone; adds parameters ball_i_0_0 = {C,-1000,0,0,T}
    void addTrackConstraints();
    // explicit vertical cylinder (pm3d / fw6), in addition to the XML exclusion zones
    void addCylinder(double cx, double cy, double radius);

    // ---- batches of independent instances (evaluation-only use, BASELINE configs 2 and 5) ---------
    // B copies of the loaded VGP; instance b can then be given its own obstacle data
    void setBatch(size_t nInstances);
    size_t getBatch() const;
    // raw instance-data block of instance b (layout: ecuda_upload_instances); valid after setup()
    // per-instance data block (layout of ecuda_upload_instances). Edits survive the mesh refinement of solve();
    // setup() starts again from the loaded VGP. Edits of getProblem()'s node-dependent arrays (zl, zu, gl, gu, sz,
    // guess) apply to the current mesh only: per-state / per-control bounds belong in the VGP (setXlower, ...).
    std::vector<double>& instanceData(size_t b);
    void uploadInstances();  // push edited instance data to the device
    // test hooks: transcription as solve() re-runs it on a mesh of `nodes` nodes (from_setup: as setup() does)
    void retranscribeForTest(int nodes, bool from_setup) {
        if (from_setup) _inst_user = false;
        _nodes = nodes;
        transcribe();
    }
    const std::vector<std::vector<double>>& instanceBlocksForTest() const { return _inst; }

    // next node count of the automatic mesh refinement, from the (nodes, error) history of the solves so far
    static int nextMeshSize(const std::vector<std::pair<int, double>>& history, const ecuda_alg_t& alg);

    // ---- evaluation (the hot path), host buffers --------------------------------------------------
    // z: [B][nvars] unscaled decision vectors; outputs may be null. Values are returned in the
    // solver's (scaled) space exactly as IPOPT would see them.
    int evaluate(const double* z, double* f, double* g, double* jac);
    int evaluateGradient(const double* z, double* grad);
    ecuda_handle handle();  // the C-ABI handle, for callers that manage device buffers themselves

 private:
    void fillDesc(ecuda_problem_desc* d, int model, bool obstacles, bool tracks);
    bool usesEdges(int model) const;
    void buildBounds();
    void buildScaling();
    void buildInstance(std::vector<double>* out) const;
    void buildInstanceFor(std::vector<double>* out, int model, bool obstacles, bool tracks, const ecuda_problem_desc& d,
                          int inst_stride) const;
    void extractTrajectories(const std::vector<double>& z);
    int solveOnce();     // one NLP solve on the current mesh; fills _solution, returns its error flag
    void deviceSetup();  // (re)creates the device evaluator for the current transcription
    void fail(const std::string& what);

    ecuda_alg_t _algorithm;
    ecuda_prob_t _problem;
    ecuda_sol_t _solution;
    ecuda_handle _handle;
    int _model;
    bool _model_set, _obstacles_on, _tracks_on, _is_setup;
    bool _user_edges;  // user model: static path rows are edge ellipses (no explicit cylinders registered)
    int _nuser_rows = 0;  // constraint rows that are none of the built-in zone rows: traced path rows of the user model
    size_t _batch;
    int _nodes;  // collocation nodes of the current mesh (0: nsteps + 1, the first mesh)
    std::vector<std::array<double, 3>> _cylinders;
    std::vector<std::vector<double>> _inst;  // per-instance data blocks
    bool _inst_user = false;                 // handed out through instanceData(): kept across re-meshing in solve()
    std::vector<double> _zscaled;            // scratch: z * sz
};

}  // namespace ETOL
#endif  // INCLUDE_ETOL_ECUDA_HPP_
