// eCUDA_var.hpp -- the scalar type VGP callbacks are written in for the eCUDA eSolver.
//
// Every ETOL eSolver fixes the scalar its callbacks use (src/docs/source/tutorials/vgp.rst:155 of the
// reference): ePSOPT passes `adouble*` inside the std::any arguments and expects `adouble` /
// `std::vector<adouble>` back (src/ePSOPT/ePSOPT.cpp:225-270, include/ETOL/ePSOPT_Types.hpp:20).
// eCUDA passes `ecuda::var*` and expects `ecuda::var` / `ETOL::fout_ecuda_t`. A var is a handle into
// a recording of the arithmetic done on it (like an ADOL-C tape): calling a callback once with
// symbolic inputs yields an expression program that can be re-evaluated for any (x, u, t). eCUDA uses
// the recordings to find and verify the device model that implements the callbacks (eCUDA.hpp).
#ifndef INCLUDE_ETOL_ECUDA_VAR_HPP_
#define INCLUDE_ETOL_ECUDA_VAR_HPP_

#include <cmath>
#include <cstdint>
#include <vector>

namespace ecuda {

enum class Op : uint8_t { INPUT, CONST, ADD, SUB, MUL, DIV, NEG, POW, SQRT, SIN, COS, EXP };

struct Node {
    Op op;
    int a, b;     // operand node ids (INPUT: a = input slot)
    double imm;   // CONST value / POW exponent
};

// one recording; callbacks executed while a Tape is active append to it
class Tape {
 public:
    std::vector<Node> nodes;
    int push(Op op, int a = -1, int b = -1, double imm = 0.0) {
        nodes.push_back(Node{op, a, b, imm});
        return static_cast<int>(nodes.size()) - 1;
    }
    // values of all nodes for the given input slots
    void eval(const double* inputs, std::vector<double>* val) const {
        val->resize(nodes.size());
        double* v = val->data();
        for (size_t i = 0; i < nodes.size(); ++i) {
            const Node& n = nodes[i];
            switch (n.op) {
                case Op::INPUT: v[i] = inputs[n.a]; break;
                case Op::CONST: v[i] = n.imm; break;
                case Op::ADD: v[i] = v[n.a] + v[n.b]; break;
                case Op::SUB: v[i] = v[n.a] - v[n.b]; break;
                case Op::MUL: v[i] = v[n.a] * v[n.b]; break;
                case Op::DIV: v[i] = v[n.a] / v[n.b]; break;
                case Op::NEG: v[i] = -v[n.a]; break;
                case Op::POW: v[i] = std::pow(v[n.a], n.imm); break;
                case Op::SQRT: v[i] = std::sqrt(v[n.a]); break;
                case Op::SIN: v[i] = std::sin(v[n.a]); break;
                case Op::COS: v[i] = std::cos(v[n.a]); break;
                case Op::EXP: v[i] = std::exp(v[n.a]); break;
            }
        }
    }
    static Tape*& active() {
        static thread_local Tape* t = nullptr;
        return t;
    }
};

class var {
 public:
    var() : id_(-1), value_(0.0) {}
    var(double c) : id_(-1), value_(c) {}  // NOLINT: constants convert implicitly, like adouble
    static var input(int slot) {
        var v;
        v.id_ = Tape::active()->push(Op::INPUT, slot);
        return v;
    }
    int id() const {  // node id, materialising a constant on first use
        if (id_ < 0) id_ = Tape::active()->push(Op::CONST, -1, -1, value_);
        return id_;
    }
    static var make(Op op, const var& a, const var& b) {
        var r;
        r.id_ = Tape::active()->push(op, a.id(), b.id());
        return r;
    }
    static var make1(Op op, const var& a, double imm = 0.0) {
        var r;
        r.id_ = Tape::active()->push(op, a.id(), -1, imm);
        return r;
    }

 private:
    mutable int id_;
    double value_;
};

inline var operator+(const var& a, const var& b) { return var::make(Op::ADD, a, b); }
inline var operator-(const var& a, const var& b) { return var::make(Op::SUB, a, b); }
inline var operator*(const var& a, const var& b) { return var::make(Op::MUL, a, b); }
inline var operator/(const var& a, const var& b) { return var::make(Op::DIV, a, b); }
inline var operator-(const var& a) { return var::make1(Op::NEG, a); }
inline var pow(const var& a, double e) { return var::make1(Op::POW, a, e); }
inline var sqrt(const var& a) { return var::make1(Op::SQRT, a); }
inline var sin(const var& a) { return var::make1(Op::SIN, a); }
inline var cos(const var& a) { return var::make1(Op::COS, a); }
inline var exp(const var& a) { return var::make1(Op::EXP, a); }

}  // namespace ecuda

namespace ETOL {
using fout_ecuda_t = std::vector<ecuda::var>;  // what a constraint callback returns (cf. fout_psopt_t)
}

#endif  // INCLUDE_ETOL_ECUDA_VAR_HPP_
