// ecuda_models.cuh -- device models of the VGP callbacks: state derivatives, running cost, path
// (obstacle) constraint rows and their analytic partials.
//
// These replace, on the device, the user lambdas that ePSOPT::dae / integrand_cost call per node
// (src/ePSOPT/ePSOPT.cpp:186-276):
//   si2d  : objFunction, dxdt, dydt, obsConstraint, saaConstraint of
//           src/Examples/PSOPT/etol_psopt_example1.cpp:101-258 (ellipse per polygon edge :174-181,
//           circle around a linearly interpolated track point :238-246; interpolation rule of
//           include/ETOL/TrajectoryOptimizer.hpp:239-257)
//   pm3d, fw6 : the build-defined UAS models of SURVEY.md section 8(d)
// Analytic partials follow the forms in src/Examples/Dymos/etol_dymos_example1.cpp:239-240,296-297
// (without that example's exp() wrapping).
//
// Every function is __host__ __device__ so that the identical source can be stepped through on a
// CPU by the kernel-logic emulator under tests/emu (a debugging aid; the product never runs it).
// Operation order is normative: no implicit fma (-fmad=false), explicit
// fma() only where written.
#ifndef ECUDA_MODELS_CUH_
#define ECUDA_MODELS_CUH_

#include "../../include/ecuda_detmath.h"
#include "ecuda_internal.hpp"

namespace ecuda {

template <int M>
struct Model;

// does f_i of model M read state j / control c? (FX / FU: 8 bits per i)
template <int M>
ECUDA_HD bool reads_state(int i, int j) { return (Model<M>::FX >> (8 * i + j)) & 1ull; }
template <int M>
ECUDA_HD bool reads_control(int i, int c) { return (Model<M>::FU >> (8 * i + c)) & 1ull; }

// ------------------------------------------------------------------------------------------------------
// path rows shared by the models
// cylinder record: cx, cy, r^2, 0
ECUDA_HD double cylinder_row(const double* rec, double x, double y) {
    double dx = x - rec[0];
    double dy = y - rec[1];
    return rec[2] - (dx * dx + dy * dy);
}
ECUDA_HD void cylinder_row_dxy(const double* rec, double x, double y, double* ddx, double* ddy) {
    *ddx = -2.0 * (x - rec[0]);
    *ddy = -2.0 * (y - rec[1]);
}

// second derivatives of the cylinder row: d2/dx2 = d2/dy2 = -2, d2/dxdy = 0
ECUDA_HD void cylinder_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {
    (void)rec;
    *hxx = -2.0;
    *hxy = 0.0;
    *hyy = -2.0;
}

// si2d edge record: xc, yc, cos(tt), sin(tt), asq, bsq
ECUDA_HD double edge_row(const double* rec, double x, double y) {
    double dx = x - rec[0];
    double dy = y - rec[1];
    double delx = rec[2] * dx - rec[3] * dy;
    double dely = rec[3] * dx + rec[2] * dy;
    return rec[4] * rec[5] - (rec[5] * (delx * delx) + rec[4] * (dely * dely));
}
ECUDA_HD void edge_row_dxy(const double* rec, double x, double y, double* ddx, double* ddy) {
    double dx = x - rec[0];
    double dy = y - rec[1];
    double ct = rec[2], st = rec[3], asq = rec[4], bsq = rec[5];
    double delx = ct * dx - st * dy;
    double dely = st * dx + ct * dy;
    *ddx = -2.0 * (bsq * delx * ct + asq * dely * st);
    *ddy = -2.0 * (-bsq * delx * st + asq * dely * ct);
}

ECUDA_HD void edge_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {
    const double ct = rec[2], st = rec[3], asq = rec[4], bsq = rec[5];
    *hxx = -2.0 * (bsq * (ct * ct) + asq * (st * st));
    *hxy = -2.0 * ((asq - bsq) * (ct * st));
    *hyy = -2.0 * (bsq * (st * st) + asq * (ct * ct));
}

// track record: radius, then nway x (t, x, y). Interval choice and formula as in ETOL's
// TrajectoryOptimizer::linear_interpolation.
ECUDA_HD int track_interval(const double* trk, int nway, double t) {
    int j = 0;
    if (t > trk[1 + 3 * (nway - 1)]) {
        j = nway - 2;
    } else if (t >= trk[1]) {
        for (int c = 0; c + 1 < nway; ++c)
            if (t >= trk[1 + 3 * c] && t <= trk[1 + 3 * (c + 1)]) j = c;
    }
    return j;
}
// centre of the moving zone at time t (the two divisions and the waypoint search of a track row), and the row value
// for a given centre: track_row is their composition, so callers that evaluate one row at several positions and one
// time (value and the x / y finite differences) may compute the centre once and get the same bits
ECUDA_HD void track_center(const double* trk, int nway, double t, double* xc, double* yc) {
    int j = track_interval(trk, nway, t);
    const double* a = trk + 1 + 3 * j;
    const double* b = a + 3;
    *xc = (t - a[0]) * (b[1] - a[1]) / (b[0] - a[0]) + a[1];
    *yc = (t - a[0]) * (b[2] - a[2]) / (b[0] - a[0]) + a[2];
}
ECUDA_HD double track_row_at(const double* trk, double xc, double yc, double x, double y) {
    double dx = x - xc;
    double dy = y - yc;
    double dist = dx * dx + dy * dy;
    return dist * (-1.) + trk[0] * trk[0];
}
ECUDA_HD double track_row(const double* trk, int nway, double x, double y, double t) {
    double xc, yc;
    track_center(trk, nway, t, &xc, &yc);
    return track_row_at(trk, xc, yc, x, y);
}
ECUDA_HD void track_row_partials(const double* trk, int nway, double x, double y, double t, double* ddx,
                                 double* ddy, double* ddt) {
    int j = track_interval(trk, nway, t);
    const double* a = trk + 1 + 3 * j;
    const double* b = a + 3;
    double sx = (b[1] - a[1]) / (b[0] - a[0]);
    double sy = (b[2] - a[2]) / (b[0] - a[0]);
    // the centre from the slopes (two divisions instead of four; the value row keeps the reference's formula, the
    // analytic partials are compared to 1e-9 and differ from it in the last bit at most)
    double xc = (t - a[0]) * sx + a[1];
    double yc = (t - a[0]) * sy + a[2];
    double dx = x - xc;
    double dy = y - yc;
    *ddx = -2.0 * dx;
    *ddy = -2.0 * dy;
    *ddt = 2.0 * (dx * sx + dy * sy);
}

// second derivatives of the moving-circle row inside the waypoint interval of t:
// d2/dx2 = d2/dy2 = -2, d2/dxdt = 2 sx, d2/dydt = 2 sy, d2/dt2 = -2 (sx^2 + sy^2)
ECUDA_HD void track_row_hess(const double* trk, int nway, double t, double* hxt, double* hyt, double* htt) {
    int j = track_interval(trk, nway, t);
    const double* a = trk + 1 + 3 * j;
    const double* b = a + 3;
    double sx = (b[1] - a[1]) / (b[0] - a[0]);
    double sy = (b[2] - a[2]) / (b[0] - a[0]);
    *hxt = 2.0 * sx;
    *hyt = 2.0 * sy;
    *htt = -2.0 * (sx * sx + sy * sy);
}

// ------------------------------------------------------------------------------------------------------
// Every model also provides hess(x, u, lam, lamL, H): H[a][b] = sum_i lam[i] d2 f_i + lamL d2 L over the
// node variables ordered [x_0..x_NS-1, u_0..u_NCU-1] (full symmetric matrix), for the Lagrangian Hessian.
template <>
struct Model<ECUDA_MODEL_SI2D> {
    static constexpr int NS = 2, NCU = 2, REC = 6;
    // f_i never reads x_i: the defect row (k,i) depends on its own state column only through D, so the
    // diagonal triplet of that column needs no second evaluation of the dynamics
    static constexpr bool DIAG_FREE = true;
    static constexpr bool TDEP = false;  // neither the dynamics nor the running cost read t
    static constexpr int NUSER = 0;  // traced path rows exist for user models only
    ECUDA_HD static void dtime(const double*, const double*, double, double*, double* dLdt) { *dLdt = 0.0; }
    // which states / controls f_i reads, 8 bits per i (same data as model_info() on the host): a
    // finite-difference triplet of a variable f_i does not read is exactly +0.0 and is stored as such
    static constexpr unsigned long long FX = 0x0000ull, FU = 0x0201ull;
    ECUDA_HD static void f(const double* x, const double* u, double t, double* out) {
        out[0] = u[0];
        out[1] = u[1];
    }
    ECUDA_HD static double cost(const double* x, const double* u, double t) { return u[0] * u[0] + u[1] * u[1]; }
    ECUDA_HD static void dcost(const double* x, const double* u, double t, double* dx, double* du) {
        dx[0] = 0.0; dx[1] = 0.0;
        du[0] = 2.0 * u[0]; du[1] = 2.0 * u[1];
    }
    // dfdx[i][j] = d f_i / d x_j ; dfdu[i][j] = d f_i / d u_j (only the NCU controls the model reads)
    ECUDA_HD static void jac(const double* x, const double* u, double t, double (*dfdx)[NS], double (*dfdu)[NCU]) {
        dfdx[0][0] = 0.0; dfdx[0][1] = 0.0; dfdx[1][0] = 0.0; dfdx[1][1] = 0.0;
        dfdu[0][0] = 1.0; dfdu[0][1] = 0.0; dfdu[1][0] = 0.0; dfdu[1][1] = 1.0;
    }
    ECUDA_HD static void hess(const double* x, const double* u, double t, const double* lam, double lamL,
                              double (*H)[NS + NCU]) {
        for (int a = 0; a < NS + NCU; ++a)
            for (int b = 0; b < NS + NCU; ++b) H[a][b] = (a == b && a >= NS) ? 2.0 * lamL : 0.0;
    }
    ECUDA_HD static double static_row(const double* rec, double x, double y) { return edge_row(rec, x, y); }
    ECUDA_HD static void static_row_dxy(const double* rec, double x, double y, double* a, double* b) {
        edge_row_dxy(rec, x, y, a, b);
    }
    ECUDA_HD static void static_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {
        edge_row_hess(rec, hxx, hxy, hyy);
    }
};

template <>
struct Model<ECUDA_MODEL_PM3D> {
    static constexpr int NS = 6, NCU = 3, REC = 4;
    static constexpr bool DIAG_FREE = true;  // f_i never reads x_i
    static constexpr bool TDEP = false;
    static constexpr int NUSER = 0;  // traced path rows exist for user models only
    ECUDA_HD static void dtime(const double*, const double*, double, double*, double* dLdt) { *dLdt = 0.0; }
    // 8 bits per i: f_0..f_2 read x_3..x_5, f_3..f_5 read u_0..u_2
    static constexpr unsigned long long FX = 0x000000201008ull, FU = 0x040201000000ull;
    ECUDA_HD static void f(const double* x, const double* u, double t, double* out) {
        out[0] = x[3]; out[1] = x[4]; out[2] = x[5];
        out[3] = u[0]; out[4] = u[1]; out[5] = u[2];
    }
    ECUDA_HD static double cost(const double* x, const double* u, double t) {
        return (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2];
    }
    ECUDA_HD static void dcost(const double* x, const double* u, double t, double* dx, double* du) {
        for (int i = 0; i < NS; ++i) dx[i] = 0.0;
        for (int j = 0; j < NCU; ++j) du[j] = 2.0 * u[j];
    }
    ECUDA_HD static void jac(const double* x, const double* u, double t, double (*dfdx)[NS], double (*dfdu)[NCU]) {
        for (int i = 0; i < NS; ++i) {
            for (int j = 0; j < NS; ++j) dfdx[i][j] = (i < 3 && j == i + 3) ? 1.0 : 0.0;
            for (int j = 0; j < NCU; ++j) dfdu[i][j] = (i >= 3 && j == i - 3) ? 1.0 : 0.0;
        }
    }
    ECUDA_HD static void hess(const double* x, const double* u, double t, const double* lam, double lamL,
                              double (*H)[NS + NCU]) {
        for (int a = 0; a < NS + NCU; ++a)
            for (int b = 0; b < NS + NCU; ++b) H[a][b] = (a == b && a >= NS) ? 2.0 * lamL : 0.0;
    }
    ECUDA_HD static void static_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {
        cylinder_row_hess(rec, hxx, hxy, hyy);
    }
    ECUDA_HD static double static_row(const double* rec, double x, double y) { return cylinder_row(rec, x, y); }
    ECUDA_HD static void static_row_dxy(const double* rec, double x, double y, double* a, double* b) {
        cylinder_row_dxy(rec, x, y, a, b);
    }
};

template <>
struct Model<ECUDA_MODEL_FW6> {
    static constexpr int NS = 6, NCU = 3, REC = 4;
    static constexpr bool DIAG_FREE = true;  // f_i never reads x_i (x,y,z,V,gamma,psi derivatives)
    static constexpr bool TDEP = false;
    static constexpr int NUSER = 0;  // traced path rows exist for user models only
    ECUDA_HD static void dtime(const double*, const double*, double, double*, double* dLdt) { *dLdt = 0.0; }
    // 8 bits per i: f_0, f_1 read V, gamma, psi; f_2 reads V, gamma; f_3 reads gamma and u_0; f_4, f_5 read u_1, u_2
    static constexpr unsigned long long FX = 0x000010183838ull, FU = 0x040201000000ull;
    static constexpr double G0 = 9.80665;
    ECUDA_HD static void f(const double* x, const double* u, double t, double* out) {
        double sg, cg, sp, cp;
        ecuda_sincos(x[4], &sg, &cg);
        ecuda_sincos(x[5], &sp, &cp);
        out[0] = (x[3] * cg) * cp;
        out[1] = (x[3] * cg) * sp;
        out[2] = x[3] * sg;
        out[3] = u[0] - G0 * sg;
        out[4] = u[1];
        out[5] = u[2];
    }
    ECUDA_HD static double cost(const double* x, const double* u, double t) {
        return (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2];
    }
    ECUDA_HD static void dcost(const double* x, const double* u, double t, double* dx, double* du) {
        for (int i = 0; i < NS; ++i) dx[i] = 0.0;
        for (int j = 0; j < NCU; ++j) du[j] = 2.0 * u[j];
    }
    ECUDA_HD static void jac(const double* x, const double* u, double t, double (*dfdx)[NS], double (*dfdu)[NCU]) {
        double sg, cg, sp, cp;
        ecuda_sincos(x[4], &sg, &cg);
        ecuda_sincos(x[5], &sp, &cp);
        const double V = x[3];
        for (int i = 0; i < NS; ++i) {
            for (int j = 0; j < NS; ++j) dfdx[i][j] = 0.0;
            for (int j = 0; j < NCU; ++j) dfdu[i][j] = (i >= 3 && j == i - 3) ? 1.0 : 0.0;
        }
        dfdx[0][3] = cg * cp;  dfdx[0][4] = -(V * sg) * cp;  dfdx[0][5] = -(V * cg) * sp;
        dfdx[1][3] = cg * sp;  dfdx[1][4] = -(V * sg) * sp;  dfdx[1][5] = (V * cg) * cp;
        dfdx[2][3] = sg;       dfdx[2][4] = V * cg;
        dfdx[3][4] = -G0 * cg;
    }
    ECUDA_HD static void hess(const double* x, const double* u, double t, const double* lam, double lamL,
                              double (*H)[NS + NCU]) {
        double sg, cg, sp, cp;
        ecuda_sincos(x[4], &sg, &cg);
        ecuda_sincos(x[5], &sp, &cp);
        const double V = x[3];
        for (int a = 0; a < NS + NCU; ++a)
            for (int b = 0; b < NS + NCU; ++b) H[a][b] = (a == b && a >= NS) ? 2.0 * lamL : 0.0;
        // f0 = V cg cp, f1 = V cg sp, f2 = V sg, f3 = u0 - G0 sg   over (V, gamma, psi) = x[3..5]
        const double vg = lam[0] * (-(sg * cp)) + lam[1] * (-(sg * sp)) + lam[2] * cg;
        const double vp = lam[0] * (-(cg * sp)) + lam[1] * (cg * cp);
        const double gg = lam[0] * (-(V * cg) * cp) + lam[1] * (-(V * cg) * sp) + lam[2] * (-(V * sg)) + lam[3] * (G0 * sg);
        const double gp = lam[0] * ((V * sg) * sp) + lam[1] * (-(V * sg) * cp);
        const double pp = lam[0] * (-(V * cg) * cp) + lam[1] * (-(V * cg) * sp);
        H[3][4] = H[4][3] = vg;
        H[3][5] = H[5][3] = vp;
        H[4][4] = gg;
        H[4][5] = H[5][4] = gp;
        H[5][5] = pp;
    }
    ECUDA_HD static void static_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {
        cylinder_row_hess(rec, hxx, hxy, hyy);
    }
    ECUDA_HD static double static_row(const double* rec, double x, double y) { return cylinder_row(rec, x, y); }
    ECUDA_HD static void static_row_dxy(const double* rec, double x, double y, double* a, double* b) {
        cylinder_row_dxy(rec, x, y, a, b);
    }
};

}  // namespace ecuda

// Model<ECUDA_MODEL_USER>: generated from a callback tape (ecuda_usermodel.cpp) and compiled with the
// kernels at run time (NVRTC), or on the host by the kernel-logic emulator of the test-suite
#ifdef ECUDA_USER_MODEL_HEADER
#include ECUDA_USER_MODEL_HEADER
#endif
#endif
