// Expression bytecode shared by host and device.
//
// Replaces deal.II FunctionParser / muParser as used by Diffusion_B
// (reference: source/equation_data/eqn_coeff_B.cc:72-91) and RightHandSideParsed
// (source/equation_data/eqn_rhs.cc:58-107): the expression strings of the .prm are
// compiled once on the host into a small stack program that both the host and the
// coefficient-sampling kernel interpret.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <map>
#include <vector>

#if defined(__CUDACC__)
#define MSFEC_HD __host__ __device__ __forceinline__
#else
#define MSFEC_HD inline
#endif

namespace msfec {

// malformed expression / constants string of a .prm (-> MSFEC_EPARSE at the C ABI)
struct ParseError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum ExprOp : int {
  OP_CONST = 0, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG,
  OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH,
  OP_EXP, OP_LOG, OP_LOG10, OP_LOG2, OP_SQRT, OP_ABS, OP_SIGN, OP_FLOOR, OP_CEIL,
  OP_MIN, OP_MAX
};

struct ExprInstr {
  int op;
  int arg;      // OP_VAR: variable index
  double val;   // OP_CONST
};

constexpr int kExprMaxStack = 24;

// prog[0..len) evaluated at (x, y, z).
MSFEC_HD double expr_eval(const ExprInstr *prog, int len, double x, double y, double z) {
  double st[kExprMaxStack];
  int sp = 0;
  for (int pc = 0; pc < len; ++pc) {
    const ExprInstr in = prog[pc];
    switch (in.op) {
      case OP_CONST: st[sp++] = in.val; break;
      case OP_VAR: st[sp++] = in.arg == 0 ? x : (in.arg == 1 ? y : z); break;
      case OP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
      case OP_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
      case OP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
      case OP_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
      case OP_POW: {   // small integer exponents (x^2, x^3 in the reference's .prm right-hand sides) as products
        --sp;
        const double b = st[sp - 1], e = st[sp];
        st[sp - 1] = e == 2.0 ? b * b : (e == 3.0 ? b * b * b : (e == 1.0 ? b : pow(b, e)));
        break;
      }
      case OP_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
      case OP_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
      case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
      case OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
      case OP_TAN: st[sp - 1] = tan(st[sp - 1]); break;
      case OP_ASIN: st[sp - 1] = asin(st[sp - 1]); break;
      case OP_ACOS: st[sp - 1] = acos(st[sp - 1]); break;
      case OP_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
      case OP_SINH: st[sp - 1] = sinh(st[sp - 1]); break;
      case OP_COSH: st[sp - 1] = cosh(st[sp - 1]); break;
      case OP_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
      case OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
      case OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
      case OP_LOG10: st[sp - 1] = log10(st[sp - 1]); break;
      case OP_LOG2: st[sp - 1] = log2(st[sp - 1]); break;
      case OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
      case OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      case OP_SIGN: st[sp - 1] = (st[sp - 1] > 0) - (st[sp - 1] < 0); break;
      case OP_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
      case OP_CEIL: st[sp - 1] = ceil(st[sp - 1]); break;
      default: break;
    }
  }
  return st[0];
}

#ifndef __CUDA_ARCH__
// Host-side compiler.  Throws std::runtime_error on syntax errors.
std::vector<ExprInstr> expr_compile(const std::string &text,
                                    const std::map<std::string, double> &constants);
// "a=1, b=2" -> map (deal.II "Function constants" syntax)
std::map<std::string, double> expr_parse_constants(const std::string &s);
// split a ';'-separated list of component expressions
std::vector<std::string> expr_split_components(const std::string &s);
#endif

}  // namespace msfec
