// etol_ecuda_example3.cpp -- a VGP whose dynamics no built-in device model implements, through eCUDA.
//
// The vehicle of example 2 flies through a wind field that depends on its position
// (vgp_si2d::windyXdot / windyYdot). Call sequence as src/Examples/PSOPT/etol_psopt_example1.cpp:41-81 of
// the reference. At setup() eCUDA records the callbacks, finds that objective and state derivatives
// are none of its built-in models, registers the recording as a user model
// (ecuda_register_user_model, include/ecuda.h) and compiles the evaluation kernels for it with
// NVRTC; the exclusion-zone and moving-zone constraints are still recognised from the VGP data.
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

    ETOL::f_t cost = &vgp_si2d::effort, fx = &vgp_si2d::windyXdot, fy = &vgp_si2d::windyYdot;
    t->setObjective(&cost);
    t->setGradient({&fx, &fy});
    ETOL::f_t zones = vgp_si2d::exclusionZones(t), movers = vgp_si2d::movingZones(t);
    t->setConstraints({&zones, &movers});

    solver.getAlgorithm()->ode_tolerance = 1.e-6;  // tighter than ePSOPT's 1e-4 so that the refinement shows
    t->setup();
    printf("device model: %s\n", solver.isUserModel() ? "user model compiled from the callbacks" : "built-in");
    t->solve();

    // the dynamics are nonlinear, so the collocation solution has a discretisation error between the nodes:
    // eCUDA estimates it on the device and refines the mesh like PSOPT does for ePSOPT (ePSOPT.cpp:69-71)
    for (const auto& m : solver.getSolution()->mesh_history)
        printf("mesh: %d nodes, max relative local error %.3e\n", m.first, m.second);
    printf("\nMinimization Score:\t%f\n", t->getScore());
    printf("State variables saved in %s\n", ETOL::TrajectoryOptimizer::save(t->getXtraj(), "state_ecuda3.csv").c_str());
    printf("Control variables saved in %s\n",
           ETOL::TrajectoryOptimizer::save(t->getUtraj(), "control_ecuda3.csv").c_str());
    t->close();
    return EXIT_SUCCESS;
}
