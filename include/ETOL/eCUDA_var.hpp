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

#include <atomic>
#include <cmath>
#include <cstdint>
#include <stdexcept>
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
    // every recording has its own serial number: a var remembers which recording its node id belongs to, so a
    // constant kept across recordings (captured by a callback, static, global) is materialised again on the new tape
    // instead of pointing at an unrelated node of it
    const uint64_t serial;
    Tape() : serial(next_serial()) {}
    Tape(const Tape& o) : nodes(o.nodes), serial(next_serial()) {}
    Tape& operator=(const Tape& o) {
        nodes = o.nodes;
        return *this;
    }
    static uint64_t next_serial() {
        static std::atomic<uint64_t> n{1};
        return n.fetch_add(1);
    }
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
    var() : id_(-1), tape_(0), constant_(true), value_(0.0) {}
    var(double c) : id_(-1), tape_(0), constant_(true), value_(c) {}  // NOLINT: constants convert implicitly, like adouble
    static var input(int slot) {
        var v;
        v.constant_ = false;
        v.id_ = Tape::active()->push(Op::INPUT, slot);
        v.tape_ = Tape::active()->serial;
        return v;
    }
    // node id on the ACTIVE tape. A constant is materialised on first use per recording; a recorded value that
    // belongs to another recording cannot be used (its id would name an unrelated node).
    int id() const {
        Tape* t = Tape::active();
        if (!t) throw std::logic_error("ecuda::var used outside a recording");
        if (constant_) {
            if (id_ < 0 || tape_ != t->serial) {
                id_ = t->push(Op::CONST, -1, -1, value_);
                tape_ = t->serial;
            }
            return id_;
        }
        if (tape_ != t->serial) throw std::logic_error("ecuda::var recorded on another tape (kept across transcriptions?)");
        return id_;
    }
    static var make(Op op, const var& a, const var& b) {
        var r;
        r.constant_ = false;
        const int ia = a.id(), ib = b.id();
        r.id_ = Tape::active()->push(op, ia, ib);
        r.tape_ = Tape::active()->serial;
        return r;
    }
    static var make1(Op op, const var& a, double imm = 0.0) {
        var r;
        r.constant_ = false;
        const int ia = a.id();
        r.id_ = Tape::active()->push(op, ia, -1, imm);
        r.tape_ = Tape::active()->serial;
        return r;
    }

 private:
    mutable int id_;
    mutable uint64_t tape_;  // serial of the recording id_ belongs to
    bool constant_;
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
