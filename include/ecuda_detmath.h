/*
 * ecuda_detmath.h -- deterministic sin/cos shared by host and device.
 *
 * Why this exists: the finite-difference Jacobian amplifies a 1-ulp difference in a constraint
 * value by ~1/(2*2^-26) ~ 3e7, so GPU/CPU parity at 1e-9 needs bit-identical constraint values
 * (SURVEY.md section 7.4-1). CUDA's sin()/cos() and glibc's differ in the last bit for some
 * arguments, so models with trigonometric dynamics (fw6) call this routine on BOTH sides. Every
 * operation below is either an explicit fma() or a single IEEE operation (the build disables
 * implicit contraction: nvcc -fmad=false, g++ -ffp-contract=off), so the host and the device
 * execute the same correctly-rounded operation sequence.
 *
 * Method: Cody-Waite reduction by pi/2 with a two-term (hi/lo) constant applied through fma, then
 * degree-13 / degree-14 minimax kernels on [-pi/4, pi/4] (the classic fdlibm coefficient sets).
 * Accuracy: < 2 ulp on [-7, 7], the range the fw6 model can feed it (measured maxima over a 2.4-million-point sweep
 * against long-double libm: sin 1.56 ulp, cos 1.38 ulp, both at the edge of a reduction interval); < 1.5 ulp on
 * random samples of |x| < 100 (tests/test_detmath.py). Not intended for huge arguments.
 */
#ifndef ECUDA_DETMATH_H_
#define ECUDA_DETMATH_H_

#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#if defined(__CUDACC__)
#define ECUDA_DETMATH_FN __host__ __device__ __forceinline__
#else
#define ECUDA_DETMATH_FN static inline
#endif

ECUDA_DETMATH_FN void ecuda_sincos(double x, double* s_out, double* c_out) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_hi = 1.57079632679489655800e+00; /* nearest double to pi/2 */
    const double pio2_lo = 6.12323399573676603587e-17; /* pi/2 - pio2_hi */
    double n = rint(x * two_over_pi);
    double r = fma(-n, pio2_hi, x);
    r = fma(-n, pio2_lo, r);
    double r2 = r * r;
    /* sin kernel: r + r^3 * S(r^2) */
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, r2, -2.50507602534068634195e-08);
    ps = fma(ps, r2, 2.75573137070700676789e-06);
    ps = fma(ps, r2, -1.98412698298579493134e-04);
    ps = fma(ps, r2, 8.33333333332248946124e-03);
    ps = fma(ps, r2, -1.66666666666666324348e-01);
    double sr = fma(r * r2, ps, r);
    /* cos kernel: 1 - r^2/2 + r^4 * C(r^2) */
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, r2, 2.08757232129817482790e-09);
    pc = fma(pc, r2, -2.75573143513906633035e-07);
    pc = fma(pc, r2, 2.48015872894767294178e-05);
    pc = fma(pc, r2, -1.38888888888741095749e-03);
    pc = fma(pc, r2, 4.16666666666666019037e-02);
    double cr = fma(r2 * r2, pc, fma(-0.5, r2, 1.0));
    int q = ((int)n) & 3;
    double s, c;
    if (q == 0) {
        s = sr;
        c = cr;
    } else if (q == 1) {
        s = cr;
        c = -sr;
    } else if (q == 2) {
        s = -sr;
        c = -cr;
    } else {
        s = -cr;
        c = sr;
    }
    *s_out = s;
    *c_out = c;
}

#endif /* ECUDA_DETMATH_H_ */
