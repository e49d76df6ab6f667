// ecuda_usermodel.cpp -- user models: a recorded callback tape becomes a device model.
//
// ePSOPT evaluates whatever lambdas the caller registered by running them on ADOL-C adoubles at every
// node (src/ePSOPT/ePSOPT.cpp:186-276). eCUDA cannot call host lambdas from a kernel; instead the
// plugin runs them once on ecuda::var (include/ETOL/eCUDA_var.hpp), and the resulting tape is handed
// to ecuda_register_user_model. This file
//   1. validates the tape and finds which states / controls every f_i reads (the masks behind
//      ECUDA_PATTERN_MODEL_DEPS and the "+0.0 triplet" shortcut of the kernels),
//   2. differentiates it symbolically (forward rules with zero propagation) for the exact Jacobian
//      and the objective gradient,
//   3. prints Model<ECUDA_MODEL_USER> as CUDA source with one statement per tape node, so the
//      canonical operation order of DESIGN.md section 3 carries over, and
//   4. compiles ecuda_kernels.cuh with that model through NVRTC (loaded with dlopen, so that
//      libecuda.so itself has no link-time dependency on it) into an sm_100a image.
// Product code; nothing here touches oracle/.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>

#include "../../include/ecuda_detmath.h"
#include "ecuda_internal.hpp"

namespace ecuda {

namespace {

std::mutex g_mu;
std::vector<std::unique_ptr<UserModel>> g_models;

bool integral_pow(double e, int* n) {
    if (!(std::fabs(e) <= 8.0) || e != std::floor(e)) return false;
    *n = static_cast<int>(e);
    return true;
}

// ---- expression pool with the few simplifications that keep derivative code small -------------------
struct Pool {
    std::vector<ecuda_tape_node>& n;
    std::vector<unsigned>& deps;  // input slots every node depends on, kept in step with n
    int push(int op, int a, int b = -1, double imm = 0.0) {
        n.push_back(ecuda_tape_node{op, a, b, 0, imm});
        deps.push_back((a >= 0 ? deps[a] : 0u) | (b >= 0 ? deps[b] : 0u));
        return static_cast<int>(n.size()) - 1;
    }
    bool is_const(int id, double v) const { return id >= 0 && n[id].op == ECUDA_OP_CONST && n[id].imm == v; }
    int cst(double v) { return push(ECUDA_OP_CONST, -1, -1, v); }
    // -1 stands for "identically zero"
    int add(int a, int b) { return a < 0 ? b : b < 0 ? a : push(ECUDA_OP_ADD, a, b); }
    int neg(int a) { return a < 0 ? -1 : push(ECUDA_OP_NEG, a); }
    int sub(int a, int b) { return b < 0 ? a : a < 0 ? neg(b) : push(ECUDA_OP_SUB, a, b); }
    int mul(int a, int b) {
        if (a < 0 || b < 0) return -1;
        if (is_const(a, 1.0)) return b;
        if (is_const(b, 1.0)) return a;
        return push(ECUDA_OP_MUL, a, b);
    }
    int div(int a, int b) { return a < 0 ? -1 : push(ECUDA_OP_DIV, a, b); }
    int un(int op, int a, double imm = 0.0) { return push(op, a, -1, imm); }
};

// d node[out] / d input[slot] as a node id in the (growing) pool, -1 when identically zero
int differentiate(Pool& P, int out, int slot) {
    if (out < 0 || !((P.deps[out] >> slot) & 1u)) return -1;
    std::vector<int> d(out + 1, -1);
    std::vector<char> live(out + 1, 0);  // only what `out` is computed from
    live[out] = 1;
    for (int k = out; k >= 0; --k) {
        if (!live[k] || P.n[k].op == ECUDA_OP_INPUT) continue;
        if (P.n[k].a >= 0) live[P.n[k].a] = 1;
        if (P.n[k].b >= 0) live[P.n[k].b] = 1;
    }
    for (int k = 0; k <= out; ++k) {
        if (!live[k] || !((P.deps[k] >> slot) & 1u)) continue;
        const ecuda_tape_node nd = P.n[k];  // by value: the pool grows below
        const int a = nd.a, b = nd.b;
        switch (nd.op) {
            case ECUDA_OP_INPUT: d[k] = P.cst(1.0); break;
            case ECUDA_OP_ADD: d[k] = P.add(d[a], d[b]); break;
            case ECUDA_OP_SUB: d[k] = P.sub(d[a], d[b]); break;
            case ECUDA_OP_MUL: d[k] = P.add(P.mul(d[a], b), P.mul(a, d[b])); break;
            case ECUDA_OP_DIV:  // (a/b)' = (a' - (a/b) b') / b
                d[k] = P.div(P.sub(d[a], P.mul(k, d[b])), b);
                break;
            case ECUDA_OP_NEG: d[k] = P.neg(d[a]); break;
            case ECUDA_OP_POW: {
                const double e = nd.imm;
                int inner;
                if (e == 1.0) inner = P.cst(1.0);
                else if (e == 2.0) inner = P.mul(P.cst(2.0), a);
                else inner = P.mul(P.cst(e), P.un(ECUDA_OP_POW, a, e - 1.0));
                d[k] = P.mul(inner, d[a]);
                break;
            }
            case ECUDA_OP_SQRT: d[k] = P.div(d[a], P.mul(P.cst(2.0), k)); break;
            case ECUDA_OP_SIN: d[k] = P.mul(P.un(ECUDA_OP_COS, a), d[a]); break;
            case ECUDA_OP_COS: d[k] = P.neg(P.mul(P.un(ECUDA_OP_SIN, a), d[a])); break;
            case ECUDA_OP_EXP: d[k] = P.mul(k, d[a]); break;
            default: break;
        }
    }
    return d[out];
}

// ---- source printer -----------------------------------------------------------------------------------
std::string literal(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".en") == std::string::npos) s += ".0";  // 'n' also catches inf / nan (rejected earlier)
    return v < 0 || std::signbit(v) ? "(" + s + ")" : s;
}

// statements computing the nodes `outs` (ids, -1 allowed) need
void print_body(const UserModel& m, const std::vector<int>& outs, std::ostringstream& o) {
    const std::vector<ecuda_tape_node>& n = m.nodes;
    std::vector<char> need(n.size(), 0);
    for (int id : outs)
        if (id >= 0) need[id] = 1;
    for (int k = static_cast<int>(n.size()) - 1; k >= 0; --k) {
        if (!need[k]) continue;
        if (n[k].a >= 0 && n[k].op != ECUDA_OP_INPUT) need[n[k].a] = 1;
        if (n[k].b >= 0) need[n[k].b] = 1;
    }
    for (size_t k = 0; k < n.size(); ++k) {
        if (!need[k]) continue;
        const ecuda_tape_node& nd = n[k];
        const std::string A = "v" + std::to_string(nd.a), B = "v" + std::to_string(nd.b);
        o << "        ";
        switch (nd.op) {
            case ECUDA_OP_INPUT:
                o << "const double v" << k << " = ";
                if (nd.a < m.ns) o << "x[" << nd.a << "]";
                else if (nd.a < m.ns + m.nc) o << "u[" << nd.a - m.ns << "]";
                else o << "t";
                break;
            case ECUDA_OP_CONST: o << "const double v" << k << " = " << literal(nd.imm); break;
            case ECUDA_OP_ADD: o << "const double v" << k << " = " << A << " + " << B; break;
            case ECUDA_OP_SUB: o << "const double v" << k << " = " << A << " - " << B; break;
            case ECUDA_OP_MUL: o << "const double v" << k << " = " << A << " * " << B; break;
            case ECUDA_OP_DIV: o << "const double v" << k << " = " << A << " / " << B; break;
            case ECUDA_OP_NEG: o << "const double v" << k << " = -" << A; break;
            case ECUDA_OP_SQRT: o << "const double v" << k << " = sqrt(" << A << ")"; break;
            case ECUDA_OP_EXP: o << "const double v" << k << " = exp(" << A << ")"; break;
            case ECUDA_OP_SIN:
            case ECUDA_OP_COS:
                o << "double s" << k << ", c" << k << "; ecuda_sincos(" << A << ", &s" << k << ", &c" << k
                  << "); const double v" << k << " = " << (nd.op == ECUDA_OP_SIN ? "s" : "c") << k;
                break;
            case ECUDA_OP_POW: {
                int e;
                if (integral_pow(nd.imm, &e)) {
                    const int ae = e < 0 ? -e : e;
                    std::string prod = ae == 0 ? "1.0" : A;
                    for (int r = 1; r < ae; ++r) prod = "(" + prod + ") * " + A;
                    o << "const double v" << k << " = " << (e < 0 ? "1.0 / (" + prod + ")" : prod);
                } else {
                    o << "const double v" << k << " = pow(" << A << ", " << literal(nd.imm) << ")";
                }
                break;
            }
        }
        o << ";\n";
    }
}

std::string generate_source(const UserModel& m) {
    std::ostringstream o;
    unsigned long long FX = 0, FU = 0;
    bool diag_free = true;
    for (int i = 0; i < m.ns; ++i) {
        FX |= static_cast<unsigned long long>(m.fx[i] & 0xffu) << (8 * i);
        FU |= static_cast<unsigned long long>(m.fu[i] & 0xffu) << (8 * i);
        if ((m.fx[i] >> i) & 1u) diag_free = false;
    }
    char masks[96];
    std::snprintf(masks, sizeof masks, "FX = 0x%016llxull, FU = 0x%016llxull", FX, FU);
    o << "// generated by libecuda (ecuda_register_user_model) from a recorded callback tape; one statement per\n"
         "// tape node, in tape order\n"
         "namespace ecuda {\n"
         "template <>\nstruct Model<ECUDA_MODEL_USER> {\n"
      << "    static constexpr int NS = " << m.ns << ", NCU = " << m.nc
      << ", REC = " << (m.static_kind == ECUDA_STATIC_EDGE ? 6 : 4) << ";\n"
      << "    static constexpr bool DIAG_FREE = " << (diag_free ? "true" : "false") << ";\n"
      << "    static constexpr bool TDEP = " << (m.tdep ? "true" : "false") << ";\n"
      << "    static constexpr unsigned long long " << masks << ";\n";
    // f
    o << "    ECUDA_HD static void f(const double* x, const double* u, double t, double* out) {\n";
    std::vector<int> outs(m.f_out, m.f_out + m.ns);
    print_body(m, outs, o);
    for (int i = 0; i < m.ns; ++i) o << "        out[" << i << "] = v" << m.f_out[i] << ";\n";
    o << "    }\n";
    // cost
    o << "    ECUDA_HD static double cost(const double* x, const double* u, double t) {\n";
    print_body(m, {m.cost_out}, o);
    o << "        return v" << m.cost_out << ";\n    }\n";
    // dcost
    // dtime: d f_i / d t and d L / d t (time-dependent models; exact Jacobian and objective gradient)
    o << "    ECUDA_HD static void dtime(const double* x, const double* u, double t, double* dfdt, double* dLdt) {\n";
    outs.clear();
    for (int i = 0; i < m.ns; ++i) outs.push_back(m.dfdt[i]);
    outs.push_back(m.dcdt);
    print_body(m, outs, o);
    for (int i = 0; i < m.ns; ++i)
        o << "        dfdt[" << i << "] = " << (m.dfdt[i] < 0 ? std::string("0.0") : "v" + std::to_string(m.dfdt[i])) << ";\n";
    o << "        *dLdt = " << (m.dcdt < 0 ? std::string("0.0") : "v" + std::to_string(m.dcdt)) << ";\n    }\n";
    o << "    ECUDA_HD static void dcost(const double* x, const double* u, double t, double* dx, double* du) {\n";
    outs.clear();
    for (int i = 0; i < m.ns; ++i) outs.push_back(m.dcdx[i]);
    for (int j = 0; j < m.nc; ++j) outs.push_back(m.dcdu[j]);
    print_body(m, outs, o);
    for (int i = 0; i < m.ns; ++i)
        o << "        dx[" << i << "] = " << (m.dcdx[i] < 0 ? std::string("0.0") : "v" + std::to_string(m.dcdx[i])) << ";\n";
    for (int j = 0; j < m.nc; ++j)
        o << "        du[" << j << "] = " << (m.dcdu[j] < 0 ? std::string("0.0") : "v" + std::to_string(m.dcdu[j])) << ";\n";
    o << "    }\n";
    // jac
    o << "    ECUDA_HD static void jac(const double* x, const double* u, double t, double (*dfdx)[NS], double (*dfdu)[NCU]) {\n";
    outs.clear();
    for (int i = 0; i < m.ns; ++i) {
        for (int j = 0; j < m.ns; ++j) outs.push_back(m.dfdx[i][j]);
        for (int j = 0; j < m.nc; ++j) outs.push_back(m.dfdu[i][j]);
    }
    print_body(m, outs, o);
    for (int i = 0; i < m.ns; ++i) {
        for (int j = 0; j < m.ns; ++j)
            o << "        dfdx[" << i << "][" << j << "] = "
              << (m.dfdx[i][j] < 0 ? std::string("0.0") : "v" + std::to_string(m.dfdx[i][j])) << ";\n";
        for (int j = 0; j < m.nc; ++j)
            o << "        dfdu[" << i << "][" << j << "] = "
              << (m.dfdu[i][j] < 0 ? std::string("0.0") : "v" + std::to_string(m.dfdu[i][j])) << ";\n";
    }
    o << "    }\n";
    // hess: H = sum_i lam[i] d2 f_i + lamL d2 L over [x | u]
    o << "    ECUDA_HD static void hess(const double* x, const double* u, double t, const double* lam, double lamL,\n"
         "                              double (*H)[NS + NCU]) {\n"
         "        (void)t;\n";
    {
        const int nvn = m.ns + m.nc;
        outs.assign(m.d2.begin(), m.d2.end());
        print_body(m, outs, o);
        for (int a = 0; a < nvn; ++a)
            for (int b = a; b < nvn; ++b) {
                std::string sum;
                for (int k = 0; k <= m.ns; ++k) {
                    const int id = m.d2[(static_cast<size_t>(k) * nvn + a) * nvn + b];
                    if (id < 0) continue;
                    const std::string term =
                        (k < m.ns ? "lam[" + std::to_string(k) + "]" : std::string("lamL")) + " * v" + std::to_string(id);
                    sum = sum.empty() ? term : "(" + sum + ") + " + term;
                }
                o << "        H[" << a << "][" << b << "] = " << (sum.empty() ? std::string("0.0") : sum) << ";\n";
                if (b != a) o << "        H[" << b << "][" << a << "] = H[" << a << "][" << b << "];\n";
            }
    }
    o << "    }\n";
    // tdir: time part of the weighted second derivatives, gt[v] = sum_i lam[i] d2f_i/dv dt + lamL d2L/dv dt,
    // st = sum_i lam[i] df_i/dt + lamL dL/dt, stt = sum_i lam[i] d2f_i/dt2 + lamL d2L/dt2
    o << "    ECUDA_HD static void tdir(const double* x, const double* u, double t, const double* lam, double lamL,\n"
         "                              double* gt, double* st, double* stt) {\n"
         "        (void)x; (void)u; (void)t;\n";
    {
        const int nvn = m.ns + m.nc;
        outs.assign(m.dvt.begin(), m.dvt.end());
        outs.insert(outs.end(), m.dtt.begin(), m.dtt.end());
        for (int i = 0; i < m.ns; ++i) outs.push_back(m.dfdt[i]);
        outs.push_back(m.dcdt);
        print_body(m, outs, o);
        auto wsum = [&](const std::function<int(int)>& id) {
            std::string sum;
            for (int k = 0; k <= m.ns; ++k) {
                const int n = id(k);
                if (n < 0) continue;
                const std::string term =
                    (k < m.ns ? "lam[" + std::to_string(k) + "]" : std::string("lamL")) + " * v" + std::to_string(n);
                sum = sum.empty() ? term : "(" + sum + ") + " + term;
            }
            return sum.empty() ? std::string("0.0") : sum;
        };
        for (int a = 0; a < nvn; ++a)
            o << "        gt[" << a << "] = " << wsum([&](int k) { return m.dvt[static_cast<size_t>(k) * nvn + a]; }) << ";\n";
        o << "        *st = " << wsum([&](int k) { return k < m.ns ? m.dfdt[k] : m.dcdt; }) << ";\n";
        o << "        *stt = " << wsum([&](int k) { return m.dtt[k]; }) << ";\n    }\n";
    }
    // second derivatives of the traced path rows in (x_0, x_1, t)
    o << "    ECUDA_HD static void user_row_hess(int r, double x0, double x1, double t, double* h) {\n"
         "        const double x[2] = {x0, x1};\n        const double* u = nullptr;\n        (void)x; (void)u; (void)t;\n"
         "        for (int e = 0; e < 6; ++e) h[e] = 0.0;\n";
    for (size_t r = 0; r < m.row_out.size(); ++r) {
        o << "        if (r == " << r << ") {\n";
        std::vector<int> six(m.rhess.begin() + static_cast<long>(6 * r), m.rhess.begin() + static_cast<long>(6 * r + 6));
        print_body(m, six, o);
        for (int e = 0; e < 6; ++e)
            if (six[e] >= 0) o << "        h[" << e << "] = v" << six[e] << ";\n";
        o << "        }\n";
    }
    o << "    }\n";
    // traced path rows
    o << "    static constexpr int NUSER = " << m.row_out.size() << ";\n"
      << "    ECUDA_HD static double user_row(int r, double x0, double x1, double t) {\n"
         "        const double x[2] = {x0, x1};\n        const double* u = nullptr;\n        (void)x; (void)u; (void)t;\n";
    for (size_t r = 0; r < m.row_out.size(); ++r) {
        o << "        if (r == " << r << ") {\n";
        print_body(m, {m.row_out[r]}, o);
        o << "        return v" << m.row_out[r] << ";\n        }\n";
    }
    o << "        return 0.0;\n    }\n"
      << "    ECUDA_HD static void user_row_partials(int r, double x0, double x1, double t, double* ddx, double* ddy, double* ddt) {\n"
         "        const double x[2] = {x0, x1};\n        const double* u = nullptr;\n        (void)x; (void)u; (void)t;\n"
         "        *ddx = 0.0; *ddy = 0.0; *ddt = 0.0;\n";
    for (size_t r = 0; r < m.row_out.size(); ++r) {
        o << "        if (r == " << r << ") {\n";
        print_body(m, {m.drdx[r], m.drdy[r], m.drdt[r]}, o);
        if (m.drdx[r] >= 0) o << "        *ddx = v" << m.drdx[r] << ";\n";
        if (m.drdy[r] >= 0) o << "        *ddy = v" << m.drdy[r] << ";\n";
        if (m.drdt[r] >= 0) o << "        *ddt = v" << m.drdt[r] << ";\n";
        o << "        }\n";
    }
    o << "    }\n";
    const char* row = m.static_kind == ECUDA_STATIC_EDGE ? "edge_row" : "cylinder_row";
    o << "    ECUDA_HD static double static_row(const double* rec, double x, double y) { return " << row
      << "(rec, x, y); }\n"
      << "    ECUDA_HD static void static_row_dxy(const double* rec, double x, double y, double* a, double* b) {\n"
      << "        " << row << "_dxy(rec, x, y, a, b);\n    }\n"
      << "    ECUDA_HD static void static_row_hess(const double* rec, double* hxx, double* hxy, double* hyy) {\n"
      << "        " << row << "_hess(rec, hxx, hxy, hyy);\n    }\n};\n}  // namespace ecuda\n";
    return o.str();
}

double powi(double a, int e) {
    const int ae = e < 0 ? -e : e;
    double p = ae == 0 ? 1.0 : a;
    for (int r = 1; r < ae; ++r) p = p * a;
    return e < 0 ? 1.0 / p : p;
}

}  // namespace

// values of nodes [0, upto] at one point
static void tape_values(const UserModel& m, const double* x, const double* u, double t, std::vector<double>* val) {
    const std::vector<ecuda_tape_node>& n = m.nodes;
    val->resize(n.size());
    double* v = val->data();
    for (size_t k = 0; k < n.size(); ++k) {
        const ecuda_tape_node& nd = n[k];
        switch (nd.op) {
            case ECUDA_OP_INPUT: v[k] = nd.a < m.ns ? x[nd.a] : nd.a < m.ns + m.nc ? u[nd.a - m.ns] : t; break;
            case ECUDA_OP_CONST: v[k] = nd.imm; break;
            case ECUDA_OP_ADD: v[k] = v[nd.a] + v[nd.b]; break;
            case ECUDA_OP_SUB: v[k] = v[nd.a] - v[nd.b]; break;
            case ECUDA_OP_MUL: v[k] = v[nd.a] * v[nd.b]; break;
            case ECUDA_OP_DIV: v[k] = v[nd.a] / v[nd.b]; break;
            case ECUDA_OP_NEG: v[k] = -v[nd.a]; break;
            case ECUDA_OP_SQRT: v[k] = std::sqrt(v[nd.a]); break;
            case ECUDA_OP_EXP: v[k] = std::exp(v[nd.a]); break;
            case ECUDA_OP_SIN:
            case ECUDA_OP_COS: {
                double s, c;
                ecuda_sincos(v[nd.a], &s, &c);
                v[k] = nd.op == ECUDA_OP_SIN ? s : c;
                break;
            }
            case ECUDA_OP_POW: {
                int e;
                v[k] = integral_pow(nd.imm, &e) ? powi(v[nd.a], e) : std::pow(v[nd.a], nd.imm);
                break;
            }
        }
    }
}

void user_model_eval(const UserModel& m, const double* x, const double* u, double t, double* f_out, double* cost_out) {
    std::vector<double> v;
    tape_values(m, x, u, t, &v);
    for (int i = 0; i < m.ns; ++i) f_out[i] = v[m.f_out[i]];
    if (cost_out) *cost_out = v[m.cost_out];
}

void user_model_rows(const UserModel& m, double x0, double x1, double t, double* rows) {
    std::vector<double> v;
    const double x[ECUDA_MAX_STATES] = {x0, x1}, u[ECUDA_MAX_CONTROLS] = {0};
    tape_values(m, x, u, t, &v);
    for (size_t r = 0; r < m.row_out.size(); ++r) rows[r] = v[m.row_out[r]];
}

void user_model_partials(const UserModel& m, const double* x, const double* u, double t, double* dfdx, double* dfdu,
                         double* dcdx, double* dcdu) {
    std::vector<double> v;
    tape_values(m, x, u, t, &v);
    auto at = [&](int id) { return id < 0 ? 0.0 : v[id]; };
    for (int i = 0; i < m.ns; ++i) {
        for (int j = 0; j < m.ns; ++j) dfdx[i * m.ns + j] = at(m.dfdx[i][j]);
        for (int j = 0; j < m.nc; ++j) dfdu[i * m.nc + j] = at(m.dfdu[i][j]);
        dcdx[i] = at(m.dcdx[i]);
    }
    for (int j = 0; j < m.nc; ++j) dcdu[j] = at(m.dcdu[j]);
}

const UserModel* user_model(int model_id) {
    std::lock_guard<std::mutex> lock(g_mu);
    const int idx = model_id - ECUDA_MODEL_USER_BASE;
    if (idx < 0 || idx >= static_cast<int>(g_models.size())) return nullptr;
    return g_models[idx].get();
}

int register_user_model(const ecuda_user_model* um, int nrows, const int32_t* row_out, int32_t* model_id, std::string* err) {
    auto bad = [&](const std::string& msg) {
        *err = msg;
        return ECUDA_ERR_ARG;
    };
    if (!um || !model_id || !um->nodes) return bad("null argument");
    if (um->nstates < 2 || um->nstates > ECUDA_MAX_STATES) return bad("nstates must be 2..8 (path rows read states 0 and 1)");
    if (um->ncontrols < 1 || um->ncontrols > ECUDA_MAX_CONTROLS) return bad("ncontrols must be 1..8");
    if (um->static_kind != ECUDA_STATIC_CYLINDER && um->static_kind != ECUDA_STATIC_EDGE) return bad("bad static_kind");
    if (um->nnodes < 1 || um->nnodes > (1 << 16)) return bad("tape length must be 1..65536");
    if (nrows < 0 || nrows > ECUDA_MAX_USER_ROWS || (nrows > 0 && !row_out)) return bad("traced path rows: 0..16 per node");
    {  // the same model registered again (e.g. a plugin re-transcribing on a refined mesh): the same id
        std::lock_guard<std::mutex> lock(g_mu);
        for (const auto& e : g_models) {
            if (e->ns != um->nstates || e->nc != um->ncontrols || e->static_kind != um->static_kind ||
                e->nregistered != um->nnodes || e->cost_out != um->cost_out || static_cast<int>(e->row_out.size()) != nrows)
                continue;
            bool same = true;
            for (int r = 0; r < nrows && same; ++r) same = e->row_out[r] == row_out[r];
            for (int i = 0; i < um->nstates && same; ++i) same = e->f_out[i] == um->f_out[i];
            for (int k = 0; k < um->nnodes && same; ++k) {
                const ecuda_tape_node &a = e->nodes[k], &b = um->nodes[k];
                same = a.op == b.op && a.imm == b.imm &&
                       (a.op == ECUDA_OP_CONST || (a.a == b.a && (a.b == b.b || a.b < 0)));
            }
            if (same) {
                *model_id = e->id;
                return ECUDA_OK;
            }
        }
    }
    std::unique_ptr<UserModel> m(new UserModel);
    m->nregistered = um->nnodes;
    m->ns = um->nstates;
    m->nc = um->ncontrols;
    m->static_kind = um->static_kind;
    const int nin = m->ns + m->nc + 1;
    m->nodes.assign(um->nodes, um->nodes + um->nnodes);
    std::vector<unsigned> deps(um->nnodes, 0u);
    for (int k = 0; k < um->nnodes; ++k) {
        ecuda_tape_node& nd = m->nodes[k];
        nd.reserved = 0;
        const std::string at = "tape node " + std::to_string(k) + ": ";
        switch (nd.op) {
            case ECUDA_OP_INPUT:
                if (nd.a < 0 || nd.a >= nin) return bad(at + "input slot out of range");
                deps[k] = 1u << nd.a;
                nd.b = -1;
                break;
            case ECUDA_OP_CONST:
                if (!std::isfinite(nd.imm)) return bad(at + "constant is not finite");
                nd.a = nd.b = -1;
                break;
            case ECUDA_OP_ADD: case ECUDA_OP_SUB: case ECUDA_OP_MUL: case ECUDA_OP_DIV:
                if (nd.a < 0 || nd.a >= k || nd.b < 0 || nd.b >= k) return bad(at + "operand does not precede the node");
                deps[k] = deps[nd.a] | deps[nd.b];
                break;
            case ECUDA_OP_POW:
                if (!std::isfinite(nd.imm)) return bad(at + "exponent is not finite");
                // fallthrough
            case ECUDA_OP_NEG: case ECUDA_OP_SQRT: case ECUDA_OP_SIN: case ECUDA_OP_COS: case ECUDA_OP_EXP:
                if (nd.a < 0 || nd.a >= k) return bad(at + "operand does not precede the node");
                nd.b = -1;
                deps[k] = deps[nd.a];
                break;
            default: return bad(at + "unknown operation");
        }
    }
    auto out_ok = [&](int id) { return id >= 0 && id < um->nnodes; };
    if (!out_ok(um->cost_out)) return bad("cost_out is not a tape node");
    const unsigned tbit = 1u << (m->ns + m->nc);
    m->tdep = (deps[um->cost_out] & tbit) != 0;
    m->cost_out = um->cost_out;
    for (int i = 0; i < ECUDA_MAX_STATES; ++i) {
        m->f_out[i] = -1;
        m->fx[i] = m->fu[i] = 0;
    }
    for (int i = 0; i < m->ns; ++i) {
        if (!out_ok(um->f_out[i])) return bad("f_out[" + std::to_string(i) + "] is not a tape node");
        if (deps[um->f_out[i]] & tbit) m->tdep = true;
        m->f_out[i] = um->f_out[i];
        m->fx[i] = deps[m->f_out[i]] & ((1u << m->ns) - 1u);
        m->fu[i] = (deps[m->f_out[i]] >> m->ns) & ((1u << m->nc) - 1u);
    }
    for (int r = 0; r < nrows; ++r) {
        if (!out_ok(row_out[r])) return bad("row_out[" + std::to_string(r) + "] is not a tape node");
        if (deps[row_out[r]] & ~(3u | tbit))
            return bad("traced path row " + std::to_string(r) + " reads more than states 0, 1 and t (the read set of a "
                       "moving-zone row, whose sparsity traced rows share)");
        m->row_out.push_back(row_out[r]);
    }
    // derivatives, appended to the same node list
    Pool P{m->nodes, deps};
    for (int i = 0; i < ECUDA_MAX_STATES; ++i) {
        for (int j = 0; j < ECUDA_MAX_STATES; ++j) m->dfdx[i][j] = -1;
        for (int j = 0; j < ECUDA_MAX_CONTROLS; ++j) m->dfdu[i][j] = -1;
        m->dcdx[i] = -1;
    }
    for (int j = 0; j < ECUDA_MAX_CONTROLS; ++j) m->dcdu[j] = -1;
    for (int i = 0; i < m->ns; ++i) {
        for (int j = 0; j < m->ns; ++j) m->dfdx[i][j] = differentiate(P, m->f_out[i], j);
        for (int j = 0; j < m->nc; ++j) m->dfdu[i][j] = differentiate(P, m->f_out[i], m->ns + j);
        m->dcdx[i] = differentiate(P, m->cost_out, i);
    }
    for (int j = 0; j < m->nc; ++j) m->dcdu[j] = differentiate(P, m->cost_out, m->ns + j);
    for (int id : m->row_out) {
        m->drdx.push_back(differentiate(P, id, 0));
        m->drdy.push_back(differentiate(P, id, 1));
        m->drdt.push_back(differentiate(P, id, m->ns + m->nc));
    }
    for (int i = 0; i < ECUDA_MAX_STATES; ++i) m->dfdt[i] = -1;
    for (int i = 0; i < m->ns; ++i) m->dfdt[i] = differentiate(P, m->f_out[i], m->ns + m->nc);
    m->dcdt = differentiate(P, m->cost_out, m->ns + m->nc);
    // second derivatives over the node variables [x | u] (upper triangle a <= b), for the Lagrangian Hessian
    const int nvn = m->ns + m->nc;
    m->d2.assign(static_cast<size_t>(m->ns + 1) * nvn * nvn, -1);
    for (int o = 0; o <= m->ns; ++o)  // o < ns: f_o, o == ns: the running cost
        for (int a = 0; a < nvn; ++a) {
            const int first = o < m->ns ? (a < m->ns ? m->dfdx[o][a] : m->dfdu[o][a - m->ns])
                                        : (a < m->ns ? m->dcdx[a] : m->dcdu[a - m->ns]);
            for (int b = a; b < nvn; ++b)
                m->d2[(static_cast<size_t>(o) * nvn + a) * nvn + b] = differentiate(P, first, b);
        }
    {  // time part of the second derivatives, and the second derivatives of the traced rows
        const int tslot = m->ns + m->nc;
        m->dvt.assign(static_cast<size_t>(m->ns + 1) * nvn, -1);
        m->dtt.assign(static_cast<size_t>(m->ns + 1), -1);
        for (int o = 0; o <= m->ns; ++o) {
            for (int a = 0; a < nvn; ++a) {
                const int first = o < m->ns ? (a < m->ns ? m->dfdx[o][a] : m->dfdu[o][a - m->ns])
                                            : (a < m->ns ? m->dcdx[a] : m->dcdu[a - m->ns]);
                m->dvt[static_cast<size_t>(o) * nvn + a] = differentiate(P, first, tslot);
            }
            m->dtt[o] = differentiate(P, o < m->ns ? m->dfdt[o] : m->dcdt, tslot);
        }
        for (size_t r = 0; r < m->row_out.size(); ++r) {
            m->rhess.push_back(differentiate(P, m->drdx[r], 0));
            m->rhess.push_back(differentiate(P, m->drdx[r], 1));
            m->rhess.push_back(differentiate(P, m->drdy[r], 1));
            m->rhess.push_back(differentiate(P, m->drdx[r], tslot));
            m->rhess.push_back(differentiate(P, m->drdy[r], tslot));
            m->rhess.push_back(differentiate(P, m->drdt[r], tslot));
        }
    }
    m->source = generate_source(*m);
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_models.size() >= ECUDA_MAX_USER_MODELS) return bad("too many user models (64 per process)");
    m->id = ECUDA_MODEL_USER_BASE + static_cast<int>(g_models.size());
    *model_id = m->id;
    g_models.push_back(std::move(m));
    return ECUDA_OK;
}

// ---- NVRTC ------------------------------------------------------------------------------------------------
namespace {

struct Nvrtc {
    void* lib = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*AddNameExpression)(void*, const char*) = nullptr;
    int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

bool load_nvrtc(Nvrtc* nv, std::string* err) {
    static Nvrtc cached;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!cached.lib) {
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13"};
        for (const char* n : names)
            if ((cached.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!cached.lib) {
            *err = "user models need NVRTC, and libnvrtc.so.12 could not be loaded";
            return false;
        }
#define ECUDA_NVRTC_SYM(field, sym)                                                     \
    cached.field = reinterpret_cast<decltype(cached.field)>(dlsym(cached.lib, sym));    \
    if (!cached.field) {                                                                \
        *err = std::string("libnvrtc lacks ") + sym;                                    \
        cached.lib = nullptr;                                                           \
        return false;                                                                   \
    }
        ECUDA_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram")
        ECUDA_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
        ECUDA_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram")
        ECUDA_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        ECUDA_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
        ECUDA_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        ECUDA_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
        ECUDA_NVRTC_SYM(AddNameExpression, "nvrtcAddNameExpression")
        ECUDA_NVRTC_SYM(GetLoweredName, "nvrtcGetLoweredName")
        ECUDA_NVRTC_SYM(GetErrorString, "nvrtcGetErrorString")
#undef ECUDA_NVRTC_SYM
    }
    *nv = cached;
    return true;
}

// the kernel sources as NVRTC in-memory headers, under the names the #include directives use
const char* const kHeaderNames[] = {"ecuda_kernels.cuh", "ecuda_stream.cuh", "ecuda_rowsn.cuh", "ecuda_rows.cuh", "ecuda_fast.cuh", "ecuda_phases.cuh",
                                    "ecuda_models.cuh", "ecuda_internal.hpp", "../../include/ecuda.h",
                                    "../../include/ecuda_detmath.h"};
constexpr int kNumHeaders = sizeof(kHeaderNames) / sizeof(kHeaderNames[0]);

bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

// ECUDA_KERNEL_SOURCE_DIR (flat directory with the ten files) or the in-tree layout next to libecuda.so
bool load_kernel_sources(std::vector<std::string>* texts, std::string* err) {
    static std::vector<std::string> cached;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (cached.empty()) {
        std::string flat, tree;
        if (const char* e = std::getenv("ECUDA_KERNEL_SOURCE_DIR")) flat = e;
        Dl_info info;
        if (dladdr(reinterpret_cast<const void*>(&load_kernel_sources), &info) && info.dli_fname) {
            tree = info.dli_fname;
            const size_t s = tree.rfind('/');
            tree = s == std::string::npos ? "." : tree.substr(0, s);
        }
        std::vector<std::string> got(kNumHeaders);
        for (int i = 0; i < kNumHeaders; ++i) {
            const std::string name = kHeaderNames[i];
            const std::string base = name.substr(name.rfind('/') == std::string::npos ? 0 : name.rfind('/') + 1);
            if (!(!flat.empty() && read_file(flat + "/" + base, &got[i])) && !read_file(tree + "/" + name, &got[i]) &&
                !read_file(tree + "/" + base, &got[i])) {
                *err = "kernel source " + base + " not found (looked in ECUDA_KERNEL_SOURCE_DIR and next to libecuda.so: " +
                       tree + ")";
                return false;
            }
        }
        cached.swap(got);
    }
    *texts = cached;
    return true;
}

}  // namespace

bool user_model_compile(const UserModel& m, int nb, bool rows, int rowsn_N, bool trk, UserImage* out, std::string* err) {
    Nvrtc nv;
    if (!load_nvrtc(&nv, err)) return false;
    std::vector<std::string> texts;
    if (!load_kernel_sources(&texts, err)) return false;
    std::vector<const char*> hdr_text, hdr_name;
    for (int i = 0; i < kNumHeaders; ++i) {
        hdr_text.push_back(texts[i].c_str());
        hdr_name.push_back(kHeaderNames[i]);
    }
    hdr_text.push_back(m.source.c_str());
    hdr_name.push_back("ecuda_user_model.cuh");
    const char* main_src =
        "#define ECUDA_USER_MODEL_HEADER \"ecuda_user_model.cuh\"\n"
        "#include \"ecuda_kernels.cuh\"\n";
    void* prog = nullptr;
    int rc = nv.CreateProgram(&prog, main_src, "ecuda_user_model.cu", static_cast<int>(hdr_text.size()), hdr_text.data(),
                              hdr_name.data());
    if (rc) {
        *err = std::string("nvrtcCreateProgram: ") + nv.GetErrorString(rc);
        return false;
    }
    const std::string M = std::to_string(ECUDA_MODEL_USER), NB = std::to_string(nb);
    std::vector<std::string> exprs(UserImage::NKERNELS);
    exprs[UserImage::GENERIC] = "ecuda::k_eval<" + M + ", " + NB + ">";
    exprs[UserImage::GRAD] = "ecuda::k_grad<" + M + ">";
    exprs[UserImage::ODE_ERROR] = "ecuda::k_ode_error<" + M + ">";
    exprs[UserImage::HESS] = "ecuda::k_hess<" + M + ">";
    if (rows) {
        exprs[UserImage::ROWS_FD] = "ecuda::k_eval_rows<" + M + ", " + NB + ", true>";
        exprs[UserImage::ROWS_EXACT] = "ecuda::k_eval_rows<" + M + ", " + NB + ", false>";
        if (rowsn_N > 0)  // <M, N, FD, TRK, SUM, RING>
            exprs[UserImage::ROWSN_FD] = "ecuda::k_rows_n<" + M + ", " + std::to_string(rowsn_N) + ", true, " +
                                         (trk ? "true" : "false") + ", false, false>";
    }
    for (const std::string& e : exprs)
        if (!e.empty() && (rc = nv.AddNameExpression(prog, e.c_str()))) {
            *err = std::string("nvrtcAddNameExpression: ") + nv.GetErrorString(rc);
            nv.DestroyProgram(&prog);
            return false;
        }
    // same code generation rules as the nvcc build of the built-in models (etol_b200/_build.py)
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-fmad=false", "-lineinfo"};
    rc = nv.CompileProgram(prog, 4, opts);
    size_t n = 0;
    nv.GetProgramLogSize(prog, &n);
    out->log.assign(n, '\0');
    if (n) nv.GetProgramLog(prog, &out->log[0]);
    if (rc) {
        *err = std::string("NVRTC could not compile the user model: ") + nv.GetErrorString(rc) + "\n" + out->log;
        nv.DestroyProgram(&prog);
        return false;
    }
    nv.GetCUBINSize(prog, &n);
    out->cubin.resize(n);
    nv.GetCUBIN(prog, out->cubin.data());
    for (int k = 0; k < UserImage::NKERNELS; ++k) {
        out->name[k].clear();
        if (exprs[k].empty()) continue;
        const char* low = nullptr;
        if ((rc = nv.GetLoweredName(prog, exprs[k].c_str(), &low)) || !low) {
            *err = "nvrtcGetLoweredName failed for " + exprs[k];
            nv.DestroyProgram(&prog);
            return false;
        }
        out->name[k] = low;
    }
    nv.DestroyProgram(&prog);
    return true;
}

}  // namespace ecuda

// ---- C ABI (no device needed) --------------------------------------------------------------------------------
extern "C" {

int ecuda_register_user_model(const ecuda_user_model* m, int32_t* model_id, char* err, size_t errlen) {
    std::string e;
    const int rc = ecuda::register_user_model(m, 0, nullptr, model_id, &e);
    if (rc && err && errlen) std::snprintf(err, errlen, "%s", e.c_str());
    return rc;
}

int ecuda_register_user_model_rows(const ecuda_user_model* m, int32_t nrows, const int32_t* row_out, int32_t* model_id,
                                   char* err, size_t errlen) {
    std::string e;
    const int rc = ecuda::register_user_model(m, nrows, row_out, model_id, &e);
    if (rc && err && errlen) std::snprintf(err, errlen, "%s", e.c_str());
    return rc;
}

int ecuda_user_model_source(int32_t model_id, char* buf, size_t buflen, size_t* needed) {
    const ecuda::UserModel* m = ecuda::user_model(model_id);
    if (!m) return ECUDA_ERR_ARG;
    if (needed) *needed = m->source.size() + 1;
    if (buf && buflen) std::snprintf(buf, buflen, "%s", m->source.c_str());
    return ECUDA_OK;
}

int ecuda_user_model_compile_check(int32_t model_id, int nnodes, size_t* image_bytes, char* log, size_t loglen) {
    const ecuda::UserModel* m = ecuda::user_model(model_id);
    if (image_bytes) *image_bytes = 0;
    if (!m || nnodes < 2) return ECUDA_ERR_ARG;
    const int nb = (nnodes + ECUDA_DOT_BLOCK - 1) / ECUDA_DOT_BLOCK;
    const bool rows = nb >= 3 && nb <= 5;
    ecuda::UserImage img;
    std::string err;
    // the N-specialised kernel under the same rule as the handle applies (ecuda_api.cu: user_rowsn_N)
    const int rn = rows && m->ns * nnodes <= 256 ? nnodes : 0;
    const bool ok = ecuda::user_model_compile(*m, rows ? nb : 0, rows, rn, true, &img, &err);
    if (log && loglen) std::snprintf(log, loglen, "%s", ok ? img.log.c_str() : err.c_str());
    if (!ok) return ECUDA_ERR_CUDA;
    if (image_bytes) *image_bytes = img.cubin.size();
    return ECUDA_OK;
}

}  // extern "C"
