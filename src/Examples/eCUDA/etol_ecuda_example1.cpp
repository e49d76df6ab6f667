// etol_ecuda_example1.cpp -- the VGP of the reference's PSOPT example, solved through eCUDA.
//
// Same program shape as src/Examples/PSOPT/etol_psopt_example1.cpp:36-84 of the reference: load an
// ETOL XML configuration (resource/configs/ocp_2d_ex1.xml runs unchanged), register the objective,
// dynamics and obstacle constraints, setup, solve, save the trajectories as CSV. The three host
// lambdas of the reference (objFunction :101-114, dxdt/dydt :116-138, obsConstraint :140-197,
// saaConstraint :199-258) are replaced by the si2d device model and the two constraint
// registrations, which add the same side_i_j_0 / ball_i_0_0 parameters.
#include <ETOL/eCUDA.hpp>

#include <cstdio>
#include <cstdlib>
#include <iostream>

static void editAlgo(ETOL::TrajectoryOptimizer* t) {
    ETOL::eCUDA* ptr = dynamic_cast<ETOL::eCUDA*>(t);
    if (!ptr) {
        std::cout << "EditAlgo only works for eCUDA!" << std::endl;
        exit(EXIT_FAILURE);
    }
    ETOL::ecuda_alg_t* algo = ptr->getAlgorithm();
    algo->nlp_method = "IPOPT";       // falls back to the built-in driver when IPOPT is not linked
    algo->nlp_iter_max = 400;
    algo->nlp_tolerance = 1.e-6;
}

int main(int argc, char** argv) {
    if (argc != 2) {
        printf("Usage: %s <ETOL configuration xml filepath>\n", argv[0]);
        exit(EXIT_FAILURE);
    }
    ETOL::TrajectoryOptimizer* t;
    ETOL::eCUDA tp = ETOL::eCUDA();
    t = &tp;

    t->loadConfigs(argv[1]);
    t->printConfigs();

    t->setMaximize(false);
    tp.setModel(ECUDA_MODEL_SI2D);  // objective u0^2 + u1^2, dynamics x' = u0, y' = u1

    // Obstacle constraints
    tp.addObstacleConstraints();
    tp.addTrackConstraints();

    // Setup
    editAlgo(t);
    t->setup();
    t->debug();
    // Solve
    t->solve();

    // Results
    printf("\n!!!!!!!!!!!!!!!!!Results!!!!!!!!!!!!!!!!!\n");
    printf("Minimization Score:\t%f\n", t->getScore());
    printf("State variables saved in %s\n", ETOL::TrajectoryOptimizer::save(t->getXtraj(), "state_ecuda1.csv").c_str());
    printf("Control variables saved in %s\n",
           ETOL::TrajectoryOptimizer::save(t->getUtraj(), "control_ecuda1.csv").c_str());

    // Gracefully release resources e.g. memory, file handles, etc
    t->close();
    printf("\n!!!!!!!!!!!!!!Graceful Exit!!!!!!!!!!!!!!\n");
    return EXIT_SUCCESS;
}
