// models.hpp (oracle) -- the VGP callbacks, written the way an ETOL user writes them for ePSOPT:
// std::function<std::any(F_ARGS)> lambdas that any_cast scalar pointers out of x/u/k.
// TEST INFRASTRUCTURE (see oracle.hpp header).
//
// si2d restates src/Examples/PSOPT/etol_psopt_example1.cpp:101-258 (objective, dxdt/dydt,
// obsConstraint, saaConstraint). pm3d and fw6 are the build-defined extensions of SURVEY.md
// section 8(d) (configs C1-C5), written in the same callback style.
// The scalar type T is `double` for values and `Dual` for exact derivatives (the role ADOL-C's
// adouble plays in the reference, ePSOPT.cpp:64).
#ifndef ORACLE_MODELS_HPP_
#define ORACLE_MODELS_HPP_

#include <cmath>
#include <iterator>
#include <list>
#include <stdexcept>
#include <vector>

#include "../include/ecuda_detmath.h"
#include "oracle.hpp"

namespace oracle {

// ---- forward-mode dual number (exact first derivatives to rounding) -----------------------------
constexpr int MAXD = 17;  // ns + nc + 1 <= 8 + 8 + 1
struct Dual {
    double v = 0.0;
    double d[MAXD] = {0.0};
    Dual() {}
    Dual(double val) : v(val) {}  // NOLINT: implicit by design (constants)
};
inline Dual operator+(const Dual& a, const Dual& b) {
    Dual r(a.v + b.v);
    for (int i = 0; i < MAXD; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
inline Dual operator-(const Dual& a, const Dual& b) {
    Dual r(a.v - b.v);
    for (int i = 0; i < MAXD; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
inline Dual operator-(const Dual& a) {
    Dual r(-a.v);
    for (int i = 0; i < MAXD; ++i) r.d[i] = -a.d[i];
    return r;
}
inline Dual operator*(const Dual& a, const Dual& b) {
    Dual r(a.v * b.v);
    for (int i = 0; i < MAXD; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
inline Dual operator/(const Dual& a, const Dual& b) {
    Dual r(a.v / b.v);
    for (int i = 0; i < MAXD; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
inline void sincos_s(const double& a, double* s, double* c) { ecuda_sincos(a, s, c); }
inline void sincos_s(const Dual& a, Dual* s, Dual* c) {
    double sv, cv;
    ecuda_sincos(a.v, &sv, &cv);
    s->v = sv;
    c->v = cv;
    for (int i = 0; i < MAXD; ++i) {
        s->d[i] = cv * a.d[i];
        c->d[i] = -sv * a.d[i];
    }
}
inline double sqrt_s(const double& a) { return std::sqrt(a); }
inline Dual sqrt_s(const Dual& a) {
    Dual r(std::sqrt(a.v));
    for (int i = 0; i < MAXD; ++i) r.d[i] = a.d[i] / (2.0 * r.v);
    return r;
}
inline double exp_s(const double& a) { return std::exp(a); }
inline Dual exp_s(const Dual& a) {
    Dual r(std::exp(a.v));
    for (int i = 0; i < MAXD; ++i) r.d[i] = r.v * a.d[i];
    return r;
}
// a^e: integral |e| <= 8 is the left-to-right product (include/ecuda.h), anything else libm pow
inline double pow_s(const double& a, double e) {
    if (std::fabs(e) <= 8.0 && e == std::floor(e)) {
        const int n = static_cast<int>(std::fabs(e));
        double p = n == 0 ? 1.0 : a;
        for (int r = 1; r < n; ++r) p = p * a;
        return e < 0 ? 1.0 / p : p;
    }
    return std::pow(a, e);
}
inline Dual pow_s(const Dual& a, double e) {
    Dual r(pow_s(a.v, e));
    const double slope = e == 0.0 ? 0.0 : e * pow_s(a.v, e - 1.0);
    for (int i = 0; i < MAXD; ++i) r.d[i] = slope * a.d[i];
    return r;
}
// ---- second-order forward mode: value, gradient and full Hessian over up to MAXD2 directions ----------
// (exact second derivatives for the Lagrangian Hessian; the role ADOL-C's sparse_hess plays under PSOPT)
constexpr int MAXD2 = 18;  // ns + nc + t0 + tf
struct Dual2 {
    double v = 0.0;
    double d[MAXD2] = {0.0};
    double h[MAXD2][MAXD2] = {{0.0}};
    Dual2() {}
    Dual2(double val) : v(val) {}  // NOLINT
};
// r = phi(a) with phi' = p1, phi'' = p2 at a.v
inline Dual2 chain2(const Dual2& a, double val, double p1, double p2) {
    Dual2 r(val);
    for (int i = 0; i < MAXD2; ++i) {
        r.d[i] = p1 * a.d[i];
        for (int j = 0; j < MAXD2; ++j) r.h[i][j] = p1 * a.h[i][j] + p2 * (a.d[i] * a.d[j]);
    }
    return r;
}
inline Dual2 operator+(const Dual2& a, const Dual2& b) {
    Dual2 r(a.v + b.v);
    for (int i = 0; i < MAXD2; ++i) {
        r.d[i] = a.d[i] + b.d[i];
        for (int j = 0; j < MAXD2; ++j) r.h[i][j] = a.h[i][j] + b.h[i][j];
    }
    return r;
}
inline Dual2 operator-(const Dual2& a) { return chain2(a, -a.v, -1.0, 0.0); }
inline Dual2 operator-(const Dual2& a, const Dual2& b) { return a + (-b); }
inline Dual2 operator*(const Dual2& a, const Dual2& b) {
    Dual2 r(a.v * b.v);
    for (int i = 0; i < MAXD2; ++i) {
        r.d[i] = a.d[i] * b.v + a.v * b.d[i];
        for (int j = 0; j < MAXD2; ++j)
            r.h[i][j] = a.h[i][j] * b.v + a.d[i] * b.d[j] + a.d[j] * b.d[i] + a.v * b.h[i][j];
    }
    return r;
}
inline Dual2 operator/(const Dual2& a, const Dual2& b) {
    return a * chain2(b, 1.0 / b.v, -1.0 / (b.v * b.v), 2.0 / (b.v * b.v * b.v));
}
inline void sincos_s(const Dual2& a, Dual2* s, Dual2* c) {
    double sv, cv;
    ecuda_sincos(a.v, &sv, &cv);
    *s = chain2(a, sv, cv, -sv);
    *c = chain2(a, cv, -sv, -cv);
}
inline Dual2 sqrt_s(const Dual2& a) {
    const double r = std::sqrt(a.v);
    return chain2(a, r, 0.5 / r, -0.25 / (r * a.v));
}
inline Dual2 exp_s(const Dual2& a) {
    const double e = std::exp(a.v);
    return chain2(a, e, e, e);
}
inline Dual2 pow_s(const Dual2& a, double e) {
    const double p1 = e == 0.0 ? 0.0 : e * pow_s(a.v, e - 1.0);
    const double p2 = (e == 0.0 || e == 1.0) ? 0.0 : e * (e - 1.0) * pow_s(a.v, e - 2.0);
    return chain2(a, pow_s(a.v, e), p1, p2);
}
inline double value_of(const double& a) { return a; }
inline double value_of(const Dual& a) { return a.v; }
inline double value_of(const Dual2& a) { return a.v; }

// ETOL's own interpolation rule, include/ETOL/TrajectoryOptimizer.hpp:239-257 (interval choice on
// the value of t, the formula in T so that d/dt flows through).
template <class T>
T linear_interpolation(const T& tval, const std::vector<double>& tvec, const std::vector<double>& ref) {
    size_t j = 0;
    double tv = value_of(tval);
    if (tv > tvec.back()) {
        j = tvec.size() - 2;
    } else if (tv >= tvec.front()) {
        for (size_t c = 0; c + 1 < tvec.size(); ++c)
            if (tv >= tvec[c] && tv <= tvec[c + 1]) j = c;
    }
    return (tval - T(tvec.at(j))) * T(ref.at(j + 1) - ref.at(j)) / T(tvec.at(j + 1) - tvec.at(j)) +
           T(ref.at(j));
}

template <class T>
struct Callbacks {
    f_t objective;
    std::vector<f_t> gradient;     // one per state, like setGradient({&xdot,&ydot})
    std::vector<f_t> constraints;  // like setConstraints({&obs,&saa})
};

constexpr double kG0 = 9.80665;

// replay of a recorded user tape at one node, in scalar type T; returns the value of node `want`
template <class T>
T replay(const UserTape& ut, const vector_t& x, const vector_t& u, const std::any& k, int want) {
    std::vector<T> v(static_cast<size_t>(want) + 1);
    for (int i = 0; i <= want; ++i) {
        const TapeNode& n = ut.nodes[i];
        switch (n.op) {
            case T_INPUT:
                if (n.a < ut.ns) v[i] = *std::any_cast<T*>(x.at(n.a));
                else if (n.a < ut.ns + ut.nc) v[i] = *std::any_cast<T*>(u.at(n.a - ut.ns));
                else v[i] = *std::any_cast<T*>(k);
                break;
            case T_CONST: v[i] = T(n.imm); break;
            case T_ADD: v[i] = v[n.a] + v[n.b]; break;
            case T_SUB: v[i] = v[n.a] - v[n.b]; break;
            case T_MUL: v[i] = v[n.a] * v[n.b]; break;
            case T_DIV: v[i] = v[n.a] / v[n.b]; break;
            case T_NEG: v[i] = -v[n.a]; break;
            case T_POW: v[i] = pow_s(v[n.a], n.imm); break;
            case T_SQRT: v[i] = sqrt_s(v[n.a]); break;
            case T_EXP: v[i] = exp_s(v[n.a]); break;
            case T_SIN:
            case T_COS: {
                T s, c;
                sincos_s(v[n.a], &s, &c);
                v[i] = n.op == T_SIN ? s : c;
                break;
            }
            default: throw std::invalid_argument("bad tape op");
        }
    }
    return v[want];
}

template <class T>
Callbacks<T> make_callbacks(int model, const PhaseData* pd, const std::vector<Track>* tracks,
                            const UserTape* ut = nullptr) {
    Callbacks<T> cb;
    auto X = [](const vector_t& v, size_t i) -> const T& { return *std::any_cast<T*>(v.at(i)); };
    if (model == USER) {  // dynamics / cost from the tape; path rows of the declared kinds
        Callbacks<T> path = make_callbacks<T>(ut->edges ? SI2D : PM3D, pd, tracks);
        cb.objective = [ut](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any k,
                            std::any) -> scalar_t { return replay<T>(*ut, x, u, k, ut->cost_out); };
        for (int i = 0; i < ut->ns; ++i)
            cb.gradient.push_back([ut, i](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any k,
                                          std::any) -> scalar_t { return replay<T>(*ut, x, u, k, ut->f_out[i]); });
        cb.constraints.push_back(path.constraints.at(0));       // ellipses per edge, or cylinders
        if (ut->edges) cb.constraints.push_back(path.constraints.at(1));  // moving circles
        else if (!tracks->empty()) cb.constraints.push_back(make_callbacks<T>(SI2D, pd, tracks).constraints.at(1));
        if (!ut->row_out.empty())  // traced path rows: a third constraint callback, replayed like the dynamics
            cb.constraints.push_back([ut](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any k,
                                          std::any) -> scalar_t {
                std::vector<T> fout;
                for (int id : ut->row_out) fout.push_back(replay<T>(*ut, x, u, k, id));
                return fout;
            });
        return cb;
    }
    if (model == SI2D) {
        // objFunction, etol_psopt_example1.cpp:101-114
        cb.objective = [X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                           std::any) -> scalar_t {
            T u0 = X(u, 0), u1 = X(u, 1);
            T obj = u0 * u0 + u1 * u1;
            return obj;
        };
        // dxdt / dydt, etol_psopt_example1.cpp:116-138
        for (size_t i = 0; i < 2; ++i)
            cb.gradient.push_back([X, i](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                         std::any, std::any) -> scalar_t { return X(u, i); });
        // obsConstraint lambda, etol_psopt_example1.cpp:153-195: one ellipse per polygon edge
        cb.constraints.push_back([X, pd](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                         std::any, std::any) -> scalar_t {
            std::vector<T> fout;
            T xk = X(x, 0), yk = X(x, 1);
            for (auto bd : pd->borders) {  // by value, as the reference iterates
                size_t n = bd.size();
                for (size_t i = 0; i < n; ++i) {
                    const Corner& a = bd[i];
                    const Corner& b = bd[(i + 1) % n];  // last edge wraps, :185-186
                    EdgeGeom g = edge_geometry(a, b);
                    T dx = xk - T(g.xc);
                    T dy = yk - T(g.yc);
                    T delx = T(std::cos(g.tt)) * dx - T(std::sin(g.tt)) * dy;
                    T dely = T(std::sin(g.tt)) * dx + T(std::cos(g.tt)) * dy;
                    // pow(delx, 2.) of the reference is taken as the exact square delx*delx
                    T out = T(g.asq * g.bsq) - (T(g.bsq) * (delx * delx) + T(g.asq) * (dely * dely));
                    fout.push_back(out);
                }
            }
            return fout;
        });
        // saaConstraint lambda, etol_psopt_example1.cpp:226-255: one circle per moving track
        cb.constraints.push_back([X, tracks](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                             std::any k, std::any) -> scalar_t {
            std::vector<T> fout;
            T xk = X(x, 0), yk = X(x, 1);
            T tval = *std::any_cast<T*>(k);
            for (auto track : *tracks) {  // by value, as the reference iterates
                T xc = linear_interpolation(tval, track.t, track.x);
                T yc = linear_interpolation(tval, track.t, track.y);
                T dx = xk - xc;
                T dy = yk - yc;
                T dist = dx * dx + dy * dy;
                T circ = dist * T(-1.) + T(track.radius * track.radius);
                fout.push_back(circ);
            }
            return fout;
        });
        return cb;
    }
    // cylinders: r^2 - ((x-cx)^2 + (y-cy)^2) <= 0, same form as the moving circles above
    auto cyl = [X, pd](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                       std::any) -> scalar_t {
        std::vector<T> fout;
        T xk = X(x, 0), yk = X(x, 1);
        for (auto c : pd->cylinders) {
            T dx = xk - T(c.cx);
            T dy = yk - T(c.cy);
            T out = T(c.r * c.r) - (dx * dx + dy * dy);
            fout.push_back(out);
        }
        return fout;
    };
    if (model == PM3D) {
        cb.objective = [X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                           std::any) -> scalar_t {
            T a0 = X(u, 0), a1 = X(u, 1), a2 = X(u, 2);
            T obj = (a0 * a0 + a1 * a1) + a2 * a2;
            return obj;
        };
        for (size_t i = 0; i < 3; ++i)
            cb.gradient.push_back([X, i](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                         std::any, std::any) -> scalar_t { return X(x, 3 + i); });
        for (size_t i = 0; i < 3; ++i)
            cb.gradient.push_back([X, i](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                         std::any, std::any) -> scalar_t { return X(u, i); });
        cb.constraints.push_back(cyl);
        return cb;
    }
    // FW6: states x,y,z,V,gamma,psi ; controls aT, gamma_dot, psi_dot
    cb.objective = [X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                       std::any) -> scalar_t {
        T a0 = X(u, 0), a1 = X(u, 1), a2 = X(u, 2);
        T obj = (a0 * a0 + a1 * a1) + a2 * a2;
        return obj;
    };
    cb.gradient.push_back([X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                              std::any) -> scalar_t {
        T sg, cg, sp, cp;
        sincos_s(X(x, 4), &sg, &cg);
        sincos_s(X(x, 5), &sp, &cp);
        return (X(x, 3) * cg) * cp;
    });
    cb.gradient.push_back([X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                              std::any) -> scalar_t {
        T sg, cg, sp, cp;
        sincos_s(X(x, 4), &sg, &cg);
        sincos_s(X(x, 5), &sp, &cp);
        return (X(x, 3) * cg) * sp;
    });
    cb.gradient.push_back([X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                              std::any) -> scalar_t {
        T sg, cg;
        sincos_s(X(x, 4), &sg, &cg);
        return X(x, 3) * sg;
    });
    cb.gradient.push_back([X](vector_t x, vector_t u, vector_t, std::vector<std::string>, std::any,
                              std::any) -> scalar_t {
        T sg, cg;
        sincos_s(X(x, 4), &sg, &cg);
        return X(u, 0) - T(kG0) * sg;
    });
    for (size_t i = 1; i < 3; ++i)
        cb.gradient.push_back([X, i](vector_t x, vector_t u, vector_t, std::vector<std::string>,
                                     std::any, std::any) -> scalar_t { return X(u, i); });
    cb.constraints.push_back(cyl);
    return cb;
}

}  // namespace oracle
#endif
