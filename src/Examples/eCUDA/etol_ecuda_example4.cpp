// etol_ecuda_example4.cpp -- callbacks that read the node time, and a constraint that is none of the VGP's zones.
//
// The vehicle of example 2 flies in a gust that swings with time (vgp_si2d::gustXdot / gustYdot read `k`, the node
// time ePSOPT::dae hands to every callback, src/ePSOPT/ePSOPT.cpp:218-260 of the reference) and must also stay out
// of a disc whose radius grows with time (vgp_si2d::growingDisc) -- a third entry of setConstraints that the loaded
// VGP's exclusion zones and moving zones do not explain. Call sequence as src/Examples/PSOPT/etol_psopt_example1.cpp
// :41-81. At setup() eCUDA records all callbacks: objective and dynamics become a time-dependent user model, the
// zone constraints are recognised from the VGP data, and the remaining row is traced into the model
// (ecuda_register_user_model_rows, include/ecuda.h); NVRTC compiles the evaluation kernels for it.
#include <cstdio>
#include <cstdlib>

#include "vgp_si2d_callbacks.hpp"

int main(int argc, char** argv) {
    if (argc != 2) {
        printf("Usage: %s <ETOL configuration xml filepath>\n", argv[0]);
        return EXIT_FAILURE;
    }
    ETOL::eCUDA solver;
    ETOL::TrajectoryOptimizer* t = &solver;
    t->loadConfigs(argv[1]);
    t->setMaximize(false);

    ETOL::f_t cost = &vgp_si2d::effort, fx = &vgp_si2d::gustXdot, fy = &vgp_si2d::gustYdot;
    t->setObjective(&cost);
    t->setGradient({&fx, &fy});
    ETOL::f_t zones = vgp_si2d::exclusionZones(t), movers = vgp_si2d::movingZones(t);
    ETOL::f_t disc = vgp_si2d::growingDisc(t, 3.0, 3.5, 0.2, 0.01);  // centre, radius at t = 0, growth per second
    t->setConstraints({&zones, &movers, &disc});

    solver.getAlgorithm()->mesh_refinement = "manual";
    t->setup();
    printf("device model: %s\n", solver.isUserModel() ? "user model compiled from the callbacks" : "built-in");
    const ETOL::ecuda_prob_t* prob = solver.getProblem();
    const int N = prob->desc.nnodes[0], ns = prob->dims.nstates;
    printf("path rows per node: %d (zones and moving zones of the VGP, plus the traced disc)\n",
           (prob->dims.ncons - 1 - 2 * ns - ns * N) / N);
    t->solve();
    printf("\nMinimization Score:\t%f\n", t->getScore());
    printf("State variables saved in %s\n", ETOL::TrajectoryOptimizer::save(t->getXtraj(), "state_ecuda4.csv").c_str());
    printf("Control variables saved in %s\n",
           ETOL::TrajectoryOptimizer::save(t->getUtraj(), "control_ecuda4.csv").c_str());
    t->close();
    return EXIT_SUCCESS;
}
