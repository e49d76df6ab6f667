// collocation.cpp (oracle) -- Legendre/Chebyshev Gauss-Lobatto nodes, quadrature weights and
// differentiation matrix.  TEST INFRASTRUCTURE (see oracle.hpp header).
//
// Follows the published method PSOPT 5.0.0 implements for collocation_method="Legendre"
// (selected at src/ePSOPT/ePSOPT.cpp:68) as restated in SURVEY.md Appendix A.1; PSOPT itself is
// not vendored in the reference tree (pin: container/singularity/ETOL-examples.def:150-152).
#include <cmath>
#include <stdexcept>

#include "oracle.hpp"

namespace oracle {

namespace {

const double kPi = 3.14159265358979323846;

// P_{n-1}(x) and P_n(x) by the three-term recurrence (n >= 1)
void legendre_pair(int n, double x, double* pnm1, double* pn) {
    double a = 1.0, b = x;  // P0, P1
    for (int m = 1; m < n; ++m) {
        double c = ((2.0 * m + 1.0) * x * b - m * a) / (m + 1.0);
        a = b;
        b = c;
    }
    *pnm1 = a;
    *pn = b;
}

Collocation legendre(int N) {
    Collocation c;
    c.N = N;
    const int No = N - 1;
    c.tau.assign(N, 0.0);
    c.w.assign(N, 0.0);
    c.D.assign(static_cast<size_t>(N) * N, 0.0);
    c.tau[0] = -1.0;
    c.tau[No] = 1.0;
    for (int k = 1; 2 * k < No; ++k) {
        double x = -std::cos(kPi * k / No);
        for (int it = 0; it < 100; ++it) {
            double pm, p;
            legendre_pair(No, x, &pm, &p);
            double dx = (x * p - pm) / (N * p);
            x = x - dx;
            if (std::fabs(dx) <= 1e-16) break;
        }
        c.tau[k] = x;
        c.tau[No - k] = -x;
    }
    if (No % 2 == 0) c.tau[No / 2] = 0.0;
    std::vector<double> P(N);
    for (int k = 0; k < N; ++k) {
        double pm, p;
        legendre_pair(No, c.tau[k], &pm, &p);
        P[k] = p;
        c.w[k] = 2.0 / (No * (No + 1.0) * (p * p));
    }
    for (int k = 0; k < N; ++k)
        for (int j = 0; j < N; ++j)
            if (k != j) c.D[static_cast<size_t>(k) * N + j] = (P[k] / P[j]) / (c.tau[k] - c.tau[j]);
    c.D[0] = -(No * (No + 1.0)) / 4.0;
    c.D[static_cast<size_t>(No) * N + No] = (No * (No + 1.0)) / 4.0;
    return c;
}

Collocation chebyshev(int N) {
    Collocation c;
    c.N = N;
    const int No = N - 1;
    c.tau.assign(N, 0.0);
    c.w.assign(N, 0.0);
    c.D.assign(static_cast<size_t>(N) * N, 0.0);
    c.tau[0] = -1.0;
    c.tau[No] = 1.0;
    for (int k = 1; 2 * k < No; ++k) {
        double x = -std::cos(kPi * k / No);
        c.tau[k] = x;
        c.tau[No - k] = -x;
    }
    if (No % 2 == 0) c.tau[No / 2] = 0.0;
    for (int k = 0; k < N; ++k) {
        double ck = (k == 0 || k == No) ? 2.0 : 1.0;
        for (int j = 0; j < N; ++j) {
            if (k == j) continue;
            double cj = (j == 0 || j == No) ? 2.0 : 1.0;
            double sgn = ((k + j) % 2 == 0) ? 1.0 : -1.0;
            c.D[static_cast<size_t>(k) * N + j] = ((ck / cj) * sgn) / (c.tau[k] - c.tau[j]);
        }
        if (k > 0 && k < No)
            c.D[static_cast<size_t>(k) * N + k] = -c.tau[k] / (2.0 * (1.0 - c.tau[k] * c.tau[k]));
    }
    c.D[0] = -(2.0 * No * No + 1.0) / 6.0;
    c.D[static_cast<size_t>(No) * N + No] = (2.0 * No * No + 1.0) / 6.0;
    // Clenshaw-Curtis weights
    if (No % 2 == 0) {
        c.w[0] = c.w[No] = 1.0 / (static_cast<double>(No) * No - 1.0);
    } else {
        c.w[0] = c.w[No] = 1.0 / (static_cast<double>(No) * No);
    }
    for (int i = 1; i < No; ++i) {
        double th = kPi * i / No;
        double v = 1.0;
        if (No % 2 == 0) {
            for (int k = 1; k <= No / 2 - 1; ++k) v = v - 2.0 * std::cos(2.0 * k * th) / (4.0 * k * k - 1.0);
            v = v - std::cos(No * th) / (static_cast<double>(No) * No - 1.0);
        } else {
            for (int k = 1; k <= (No - 1) / 2; ++k) v = v - 2.0 * std::cos(2.0 * k * th) / (4.0 * k * k - 1.0);
        }
        c.w[i] = 2.0 * v / No;
    }
    return c;
}

}  // namespace

Collocation make_collocation(int kind, int N) {
    if (N < 2) throw std::invalid_argument("collocation needs at least 2 nodes");
    if (kind == LEGENDRE) return legendre(N);
    if (kind == CHEBYSHEV) return chebyshev(N);
    throw std::invalid_argument("unknown collocation kind");
}

}  // namespace oracle
