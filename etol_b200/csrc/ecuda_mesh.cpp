// ecuda_mesh.cpp -- host half of the mesh-refinement support: the
// interpolation matrices behind ecuda_ode_error and ecuda_resample.
//
// ePSOPT runs PSOPT with mesh_refinement = "automatic" and ode_tolerance = 1e-4
// (src/ePSOPT/ePSOPT.cpp:69-71; etol_psopt_example1.cpp:93-94): after every NLP solve PSOPT estimates
// the discretisation error of the collocation solution and re-solves on more nodes, starting from the
// interpolated previous solution. PSOPT's source is not in the reference tree; what is restated here
// is the published estimate it documents (Betts' relative local error): on every mesh interval
//     eta_{i,k} = integral over [t_k, t_k+1] of | d/dt x~_i(t) - f_i(x~(t), u~(t)) | dt ,
//     eps_k     = max_i eta_{i,k} / (w_i + 1),   w_i = max_k max(|x~_{i,k}|, |d/dt x~_{i,k}|),
// with x~, u~ the interpolating polynomials through the node values. The integral uses 4-point
// Gauss-Legendre quadrature per interval. Product code; nothing here touches oracle/.
#include <cmath>

#include "ecuda_internal.hpp"

namespace ecuda {

namespace {

const double kGaussX[ECUDA_MESH_Q] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};
const double kGaussW[ECUDA_MESH_Q] = {0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538};

// barycentric weights 1 / prod_{m != l} (tau_l - tau_m)
void bary_weights(const std::vector<double>& tau, std::vector<double>* bw) {
    const int N = static_cast<int>(tau.size());
    bw->assign(N, 1.0);
    for (int l = 0; l < N; ++l) {
        double p = 1.0;
        for (int m = 0; m < N; ++m)
            if (m != l) p = p * (tau[l] - tau[m]);
        (*bw)[l] = 1.0 / p;
    }
}

// Lagrange basis L_l(t) and d/dt L_l(t), l = 0..N-1 (t on a node: unit row, row of the derivative there)
void lagrange_rows(const std::vector<double>& tau, const std::vector<double>& bw, double t, double* L, double* dL) {
    const int N = static_cast<int>(tau.size());
    int hit = -1;
    for (int l = 0; l < N; ++l)
        if (t == tau[l]) hit = l;
    if (hit >= 0) {
        for (int l = 0; l < N; ++l) {
            L[l] = l == hit ? 1.0 : 0.0;
            if (dL) dL[l] = l == hit ? 0.0 : (bw[l] / bw[hit]) / (tau[hit] - tau[l]);
        }
        if (dL) {
            double s = 0.0;
            for (int l = 0; l < N; ++l)
                if (l != hit) s = s + dL[l];
            dL[hit] = -s;
        }
        return;
    }
    double s = 0.0, s2 = 0.0;
    for (int l = 0; l < N; ++l) {
        const double r = bw[l] / (t - tau[l]);
        s = s + r;
        s2 = s2 + r / (t - tau[l]);
    }
    for (int l = 0; l < N; ++l) {
        L[l] = (bw[l] / (t - tau[l])) / s;
        if (dL) dL[l] = L[l] * (s2 / s - 1.0 / (t - tau[l]));
    }
}

}  // namespace

void build_error_mesh(const Collocation& c, MeshHost* out) {
    const int N = c.N, Q = ECUDA_MESH_Q;
    std::vector<double> bw;
    bary_weights(c.tau, &bw);
    out->E.assign(static_cast<size_t>(N - 1) * Q * N, 0.0);
    out->dE.assign(static_cast<size_t>(N - 1) * Q * N, 0.0);
    out->wq.assign(static_cast<size_t>(N - 1) * Q, 0.0);
    out->tq.assign(static_cast<size_t>(N - 1) * Q, 0.0);
    for (int k = 0; k + 1 < N; ++k) {
        const double half = 0.5 * (c.tau[k + 1] - c.tau[k]), mid = 0.5 * (c.tau[k + 1] + c.tau[k]);
        for (int q = 0; q < Q; ++q) {
            const size_t r = static_cast<size_t>(k) * Q + q;
            const double t = mid + half * kGaussX[q];
            out->tq[r] = t;
            out->wq[r] = half * kGaussW[q];
            lagrange_rows(c.tau, bw, t, &out->E[r * N], &out->dE[r * N]);
        }
    }
}

void build_resample(const Collocation& from, const Collocation& to, std::vector<double>* R) {
    std::vector<double> bw;
    bary_weights(from.tau, &bw);
    R->assign(static_cast<size_t>(to.N) * from.N, 0.0);
    for (int k = 0; k < to.N; ++k) lagrange_rows(from.tau, bw, to.tau[k], &(*R)[static_cast<size_t>(k) * from.N], nullptr);
}

}  // namespace ecuda

extern "C" {

int ecuda_host_error_mesh(int kind, int nnodes, double* tq, double* wq, double* E, double* dE) {
    ecuda::Collocation c;
    std::string err;
    if (!ecuda::build_collocation(kind, nnodes, &c, &err)) return ECUDA_ERR_ARG;
    ecuda::MeshHost m;
    ecuda::build_error_mesh(c, &m);
    auto put = [](double* dst, const std::vector<double>& v) {
        if (dst)
            for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i];
    };
    put(tq, m.tq);
    put(wq, m.wq);
    put(E, m.E);
    put(dE, m.dE);
    return ECUDA_OK;
}

int ecuda_host_resample_matrix(int kind, int nnodes_from, int nnodes_to, double* R) {
    ecuda::Collocation a, b;
    std::string err;
    if (!R || !ecuda::build_collocation(kind, nnodes_from, &a, &err) || !ecuda::build_collocation(kind, nnodes_to, &b, &err))
        return ECUDA_ERR_ARG;
    std::vector<double> m;
    ecuda::build_resample(a, b, &m);
    for (size_t i = 0; i < m.size(); ++i) R[i] = m[i];
    return ECUDA_OK;
}

}  // extern "C"
