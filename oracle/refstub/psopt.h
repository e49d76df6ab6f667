// psopt.h -- STUB of the PSOPT 5.0.0 public header, test infrastructure only (oracle/_ref build).
//
// The reference's ePSOPT plugin (/root/reference/src/ePSOPT/ePSOPT.cpp) and its example
// (/root/reference/src/Examples/PSOPT/etol_psopt_example1.cpp) include <psopt.h>; PSOPT, ADOL-C, Eigen and IPOPT
// are not in this image. This file declares just enough of that interface -- adouble, MatrixXd, RowVectorXi,
// Prob / Alg / Sol / Workspace, zeros / linspace / linear_interpolation, psopt_level1_setup / psopt_level2_setup /
// psopt -- for those two reference files to compile UNMODIFIED, so that the callbacks the reference hands to PSOPT
// (dae, integrand_cost, events, endpoint_cost, the obs / saa lambdas, addBounds) can be executed at chosen
// node inputs and compared with oracle/ (tests/test_oracle_vs_reference.py). Nothing here is PSOPT code:
//   * adouble is a forward-mode dual number (value + ADOUBLE_NDIR tangents) instead of an ADOL-C tape scalar;
//   * psopt() does not solve anything (it reports an error through Sol, as PSOPT does on failure);
//   * linear_interpolation follows the interval rule of ETOL's own TrajectoryOptimizer::linear_interpolation
//     (include/ETOL/TrajectoryOptimizer.hpp:239-257 of the reference); PSOPT's routine is not in the tree, so the
//     moving-zone rows are pinned up to that routine only.
// Never included by anything under etol_b200/, src/ or include/.
#ifndef ORACLE_REFSTUB_PSOPT_H_
#define ORACLE_REFSTUB_PSOPT_H_

#include <cmath>
#include <cstddef>
#include <iostream>
#include <list>
#include <string>
#include <utility>
#include <vector>

using namespace std;  // PSOPT's header does this; the reference example relies on it (cout, endl, string)

#ifndef ADOUBLE_NDIR
#define ADOUBLE_NDIR 8
#endif

// ---- adouble: value + tangents ------------------------------------------------------------------------------------
class adouble {
 public:
    double v;
    double d[ADOUBLE_NDIR];
    adouble() : v(0.0) { clear(); }
    adouble(double x) : v(x) { clear(); }  // NOLINT (implicit, as ADOL-C's)
    double value() const { return v; }
    void clear() { for (int i = 0; i < ADOUBLE_NDIR; ++i) d[i] = 0.0; }
    adouble& operator=(double x) { v = x; clear(); return *this; }
};
inline adouble operator+(const adouble& a, const adouble& b) {
    adouble r(a.v + b.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
inline adouble operator-(const adouble& a, const adouble& b) {
    adouble r(a.v - b.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
inline adouble operator*(const adouble& a, const adouble& b) {
    adouble r(a.v * b.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
inline adouble operator/(const adouble& a, const adouble& b) {
    adouble r(a.v / b.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
inline adouble operator-(const adouble& a) {
    adouble r(-a.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = -a.d[i];
    return r;
}
inline adouble operator+(const adouble& a, double b) { adouble r(a); r.v = a.v + b; return r; }
inline adouble operator+(double a, const adouble& b) { adouble r(b); r.v = a + b.v; return r; }
inline adouble operator-(const adouble& a, double b) { adouble r(a); r.v = a.v - b; return r; }
inline adouble operator-(double a, const adouble& b) { adouble r = -b; r.v = a - b.v; return r; }
inline adouble operator*(const adouble& a, double b) {
    adouble r(a.v * b);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a.d[i] * b;
    return r;
}
inline adouble operator*(double a, const adouble& b) {
    adouble r(a * b.v);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a * b.d[i];
    return r;
}
inline adouble operator/(const adouble& a, double b) {
    adouble r(a.v / b);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = a.d[i] / b;
    return r;
}
inline adouble pow(const adouble& a, double n) {
    adouble r(std::pow(a.v, n));
    const double dr = n * std::pow(a.v, n - 1.0);
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = dr * a.d[i];
    return r;
}
inline adouble sin(const adouble& a) {
    adouble r(std::sin(a.v));
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = std::cos(a.v) * a.d[i];
    return r;
}
inline adouble cos(const adouble& a) {
    adouble r(std::cos(a.v));
    for (int i = 0; i < ADOUBLE_NDIR; ++i) r.d[i] = -std::sin(a.v) * a.d[i];
    return r;
}
inline bool operator<(const adouble& a, const adouble& b) { return a.v < b.v; }
inline bool operator>(const adouble& a, const adouble& b) { return a.v > b.v; }
inline bool operator<=(const adouble& a, const adouble& b) { return a.v <= b.v; }
inline bool operator>=(const adouble& a, const adouble& b) { return a.v >= b.v; }

// ---- the few Eigen types the reference touches ------------------------------------------------------------------------
class MatrixXd {
 public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(long r, long c) : r_(r), c_(c), a_(static_cast<size_t>(r * c), 0.0) {}
    long rows() const { return r_; }
    long cols() const { return c_; }
    double& operator()(long i, long j) { return a_[static_cast<size_t>(i * c_ + j)]; }
    double operator()(long i, long j) const { return a_[static_cast<size_t>(i * c_ + j)]; }
    MatrixXd col(long j) const {
        MatrixXd m(r_, 1);
        for (long i = 0; i < r_; ++i) m(i, 0) = (*this)(i, j);
        return m;
    }
 private:
    long r_, c_;
    std::vector<double> a_;
};
typedef MatrixXd DMatrix;

class RowVectorXi {
 public:
    RowVectorXi() {}
    explicit RowVectorXi(int n) : a_(static_cast<size_t>(n), 0) {}
    struct Comma {
        RowVectorXi* m;
        size_t at;
        Comma& operator,(int v) { m->a_[at++] = v; return *this; }
        RowVectorXi finished() { return *m; }
    };
    Comma operator<<(int v) { a_[0] = v; return Comma{this, 1}; }
    int operator()(int i) const { return a_[static_cast<size_t>(i)]; }
    long cols() const { return static_cast<long>(a_.size()); }
    std::vector<int> a_;
};

inline MatrixXd zeros(long r, long c) { return MatrixXd(r, c); }
inline MatrixXd linspace(double a, double b, long n) {
    MatrixXd m(1, n);
    for (long i = 0; i < n; ++i) m(0, i) = n > 1 ? a + (b - a) * static_cast<double>(i) / static_cast<double>(n - 1) : a;
    return m;
}

// ---- problem / algorithm / solution structures (fields the reference reads or writes) -----------------------------------
struct PsoptVec {  // bounds vectors: 0-based call syntax
    std::vector<double> a;
    double& operator()(size_t i) {
        if (i >= a.size()) a.resize(i + 1, 0.0);
        return a[i];
    }
};
struct PsoptBoundSide {
    PsoptVec states, controls, events, path, parameters;
    double StartTime = 0.0, EndTime = 0.0;
};
struct PsoptBounds { PsoptBoundSide lower, upper; };
struct PsoptGuess { MatrixXd controls, states, time, parameters; };
struct Phase {
    int nstates = 0, ncontrols = 0, nevents = 0, npath = 0, nparameters = 0;
    RowVectorXi nodes;
    PsoptGuess guess;
    PsoptBounds bounds;
};
struct Workspace;
struct Prob {
    std::string name, outfilename;
    int nphases = 0, nlinkages = 0;
    void* user_data = nullptr;
    std::vector<Phase> phase_;
    Phase& phases(int i) { return phase_.at(static_cast<size_t>(i - 1)); }  // 1-based, as in PSOPT
    adouble (*integrand_cost)(adouble*, adouble*, adouble*, adouble&, adouble*, int, Workspace*) = nullptr;
    adouble (*endpoint_cost)(adouble*, adouble*, adouble*, adouble&, adouble&, adouble*, int, Workspace*) = nullptr;
    void (*dae)(adouble*, adouble*, adouble*, adouble*, adouble*, adouble&, adouble*, int, Workspace*) = nullptr;
    void (*events)(adouble*, adouble*, adouble*, adouble*, adouble&, adouble&, adouble*, int, Workspace*) = nullptr;
    void (*linkages)(adouble*, adouble*, Workspace*) = nullptr;
};
struct Alg {
    std::string nlp_method, scaling, derivatives, hessian, collocation_method, mesh_refinement, defect_scaling;
    int nlp_iter_max = 0, mr_max_iterations = 0, print_level = 0;
    double nlp_tolerance = 0.0, ode_tolerance = 0.0, ipopt_max_cpu_time = 0.0;
};
struct Sol {
    bool error_flag = false;
    std::string error_msg;
    double cost = 0.0;
    MatrixXd states_, controls_, time_;
    MatrixXd get_states_in_phase(int) { return states_; }
    MatrixXd get_controls_in_phase(int) { return controls_; }
    MatrixXd get_time_in_phase(int) { return time_; }
};
struct Workspace { Prob* problem = nullptr; };

inline void psopt_level1_setup(Prob& p) { p.phase_.assign(static_cast<size_t>(p.nphases), Phase()); }
inline void psopt_level2_setup(Prob& p, Alg&) {
    for (Phase& ph : p.phase_) {
        for (PsoptBoundSide* s : {&ph.bounds.lower, &ph.bounds.upper}) {
            s->states.a.assign(static_cast<size_t>(ph.nstates), 0.0);
            s->controls.a.assign(static_cast<size_t>(ph.ncontrols), 0.0);
            s->events.a.assign(static_cast<size_t>(ph.nevents), 0.0);
            s->path.a.assign(static_cast<size_t>(ph.npath), 0.0);
        }
    }
}
inline void psopt(Sol& s, Prob&, Alg&) {
    s.error_flag = true;
    s.error_msg = "oracle/refstub: PSOPT is not part of this build; only the callbacks are exercised";
}
// delayed states / controls need PSOPT's interpolation of the whole trajectory: not available in the stub
inline void get_delayed_state(adouble* out, int, int, adouble&, double, adouble*, Workspace*) { *out = 0.0; }
inline void get_delayed_control(adouble* out, int, int, adouble&, double, adouble*, Workspace*) { *out = 0.0; }

// y = table lookup at x; interval rule of ETOL's own linear_interpolation (see the header comment)
inline void linear_interpolation(adouble* y, adouble& x, MatrixXd& X, MatrixXd& Y, int n) {
    long j = 0;
    if (x.value() > X(n - 1, 0)) {
        j = n - 2;
    } else if (x.value() >= X(0, 0)) {
        for (long c = 0; c + 1 < n; ++c)
            if (x.value() >= X(c, 0) && x.value() <= X(c + 1, 0)) j = c;
    }
    *y = (x - X(j, 0)) * (Y(j + 1, 0) - Y(j, 0)) / (X(j + 1, 0) - X(j, 0)) + Y(j, 0);
}

#endif  // ORACLE_REFSTUB_PSOPT_H_
