// vgp_si2d_callbacks.hpp -- the 2-D single-integrator VGP of the reference's PSOPT example written as
// eCUDA callbacks (ecuda::var scalars). Same mathematics as src/Examples/PSOPT/etol_psopt_example1.cpp
// of the reference (objective :101-114, dynamics :116-138, exclusion zones :140-197, moving zones
// :199-258), organised differently: geometry is precomputed into small structs when the constraint is
// built, the callbacks only do the per-node arithmetic.
#ifndef SRC_EXAMPLES_ECUDA_VGP_SI2D_CALLBACKS_HPP_
#define SRC_EXAMPLES_ECUDA_VGP_SI2D_CALLBACKS_HPP_

#include <ETOL/eCUDA.hpp>

#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace vgp_si2d {

using ecuda::var;

inline var& at(ETOL::vector_t& v, size_t i) { return *std::any_cast<var*>(v.at(i)); }

// running cost: control effort
inline ETOL::scalar_t effort(F_ARGS) {
    var a = at(u, 0), b = at(u, 1);
    return a * a + b * b;
}
// single integrator: the velocity is the control
inline ETOL::scalar_t xdot(F_ARGS) { return at(u, 0); }
inline ETOL::scalar_t ydot(F_ARGS) { return at(u, 1); }

// a model no built-in device model implements (eCUDA then compiles the kernels for the recorded
// callbacks): the same vehicle in a sheared wind field that depends on where it is
inline ETOL::scalar_t windyXdot(F_ARGS) { return at(u, 0) + 0.05 * at(x, 1); }
inline ETOL::scalar_t windyYdot(F_ARGS) { return at(u, 1) - 0.02 * (at(x, 0) * at(x, 0)); }

// dynamics that read the node time (`k` holds the time scalar, as in ePSOPT::dae, src/ePSOPT/ePSOPT.cpp:218-260):
// a gust that swings with a period of about half a minute
inline var& timeOf(std::any& k) { return *std::any_cast<var*>(k); }
inline ETOL::scalar_t gustXdot(F_ARGS) { return at(u, 0) + 0.3 * sin(0.2 * timeOf(k)); }
inline ETOL::scalar_t gustYdot(F_ARGS) { return at(u, 1) - 0.01 * timeOf(k); }

inline std::string rowName(const char* kind, size_t i, size_t j) {
    return std::string(kind) + "_" + std::to_string(i) + "_" + std::to_string(j) + "_0";
}

// one keep-out ellipse per polygon edge: centred on the edge midpoint, major axis along the edge
struct EdgeEllipse {
    double xc, yc, ct, st, asq, bsq;
};

// registers one path parameter per edge (bounds -1000 .. 0 over the whole horizon) and returns the
// constraint callback: value <= 0 outside the ellipse
inline ETOL::f_t exclusionZones(ETOL::TrajectoryOptimizer* t) {
    auto shapes = std::make_shared<std::vector<EdgeEllipse>>();
    const double horizon = t->getDt() * t->getNSteps();
    size_t zone = 0;
    for (const ETOL::border_t& border : *t->getObstacles_Raw()) {
        std::vector<ETOL::corner_t> c(border.begin(), border.end());
        for (size_t e = 0; e < c.size(); ++e) {
            const ETOL::corner_t &p = c[e], &q = c[(e + 1) % c.size()];
            EdgeEllipse s;
            s.xc = (q[0] + p[0]) / 2.;
            const double slope = (q[1] - p[1]) / (q[0] - p[0]);
            s.yc = p[1] + slope * (s.xc - p[0]);
            const double r2 = std::pow(s.xc - p[0], 2.0) + std::pow(s.yc - p[1], 2.0);
            const double tilt = -1.0 * std::atan2(s.yc - p[1], s.xc - p[0]);
            s.ct = std::cos(tilt);
            s.st = std::sin(tilt);
            s.asq = r2;
            s.bsq = .2 * r2;
            shapes->push_back(s);
            t->addParams({ETOL::param_t(rowName("side", zone, e), {ETOL::var_t::CONTINUOUS, -1000., 0., 0., horizon})});
        }
        ++zone;
    }
    return [shapes](F_ARGS) -> ETOL::scalar_t {
        ETOL::fout_ecuda_t rows;
        var px = at(x, 0), py = at(x, 1);
        for (const EdgeEllipse& s : *shapes) {
            var dx = px - s.xc, dy = py - s.yc;
            var along = s.ct * dx - s.st * dy;
            var across = s.st * dx + s.ct * dy;
            rows.push_back(s.asq * s.bsq - (s.bsq * (along * along) + s.asq * (across * across)));
        }
        return rows;
    };
}

// keep-out circle around every moving zone; the zone centre moves on a straight line between its two
// waypoints (branching on the time value cannot be recorded, so only two-waypoint tracks are
// written this way -- eCUDA::addTrackConstraints() handles the general table)
struct MovingZone {
    double r, t0, x0, y0, t1, x1, y1;
};
inline ETOL::f_t movingZones(ETOL::TrajectoryOptimizer* t) {
    auto zones = std::make_shared<std::vector<MovingZone>>();
    const double horizon = t->getDt() * t->getNSteps();
    size_t i = 0;
    for (const ETOL::track_t& track : *t->getTracks()) {
        const ETOL::traj_elem_t &a = track.trajectory.front(), &b = track.trajectory.back();
        zones->push_back({track.radius, a.first, a.second.at(0), a.second.at(1), b.first, b.second.at(0), b.second.at(1)});
        t->addParams({ETOL::param_t(rowName("ball", i++, 0), {ETOL::var_t::CONTINUOUS, -1000., 0., 0., horizon})});
    }
    return [zones](F_ARGS) -> ETOL::scalar_t {
        ETOL::fout_ecuda_t rows;
        var px = at(x, 0), py = at(x, 1);
        var now = *std::any_cast<var*>(k);
        for (const MovingZone& z : *zones) {
            var cx = (now - z.t0) * (z.x1 - z.x0) / (z.t1 - z.t0) + z.x0;
            var cy = (now - z.t0) * (z.y1 - z.y0) / (z.t1 - z.t0) + z.y0;
            var dx = px - cx, dy = py - cy;
            rows.push_back((dx * dx + dy * dy) * (-1.) + z.r * z.r);
        }
        return rows;
    };
}

// a path constraint that is none of the VGP's zone constraints: a keep-out disc around (cx, cy) whose radius grows
// with time, r(t) = r0 + rate * t. eCUDA records it and evaluates it on the GPU as a traced path row (it may read the
// two position states and the time, like the moving zones above). Registers its parameter (bounds -1e6 .. 0).
inline ETOL::f_t growingDisc(ETOL::TrajectoryOptimizer* t, double cx, double cy, double r0, double rate) {
    const double horizon = t->getDt() * t->getNSteps();
    t->addParams({ETOL::param_t(rowName("disc", 0, 0), {ETOL::var_t::CONTINUOUS, -1.e6, 0., 0., horizon})});
    return [cx, cy, r0, rate](F_ARGS) -> ETOL::scalar_t {
        var px = at(x, 0), py = at(x, 1);
        var r = r0 + rate * *std::any_cast<var*>(k);
        var dx = px - cx, dy = py - cy;
        ETOL::fout_ecuda_t rows;
        rows.push_back(r * r - (dx * dx + dy * dy));
        return rows;
    };
}

}  // namespace vgp_si2d
#endif  // SRC_EXAMPLES_ECUDA_VGP_SI2D_CALLBACKS_HPP_
