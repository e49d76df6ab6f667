// etol_ecuda_example2.cpp -- the reference's PSOPT example VGP with user callbacks, through eCUDA.
//
// Same call sequence as src/Examples/PSOPT/etol_psopt_example1.cpp:41-81 of the reference, including
// setObjective / setGradient / setConstraints; the callbacks (vgp_si2d_callbacks.hpp) are written with
// ecuda::var where the reference uses adouble. eCUDA records them at setup(), recognises the
// single-integrator model and the exclusion-zone / moving-zone constraints of the loaded VGP, and
// evaluates those on the GPU.
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

    ETOL::f_t cost = &vgp_si2d::effort, fx = &vgp_si2d::xdot, fy = &vgp_si2d::ydot;
    t->setObjective(&cost);
    t->setGradient({&fx, &fy});
    ETOL::f_t zones = vgp_si2d::exclusionZones(t), movers = vgp_si2d::movingZones(t);
    t->setConstraints({&zones, &movers});

    t->setup();
    t->debug();
    t->solve();

    printf("\nMinimization Score:\t%f\n", t->getScore());
    printf("State variables saved in %s\n", ETOL::TrajectoryOptimizer::save(t->getXtraj(), "state_ecuda2.csv").c_str());
    printf("Control variables saved in %s\n",
           ETOL::TrajectoryOptimizer::save(t->getUtraj(), "control_ecuda2.csv").c_str());
    t->close();
    return EXIT_SUCCESS;
}
