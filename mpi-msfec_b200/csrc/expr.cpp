// Recursive-descent compiler for the muParser subset used by the reference's .prm
// files (see expr.h).  Precedence, lowest first: + -, * /, unary +-, ^ (right
// associative, binds tighter than unary minus, as in muParser).
#include "expr.h"

#include <cctype>
#include <cstdlib>
#include <stdexcept>

namespace msfec {
namespace {

struct Parser {
  const std::string &s;
  const std::map<std::string, double> &constants;
  size_t pos = 0;
  std::vector<ExprInstr> prog;
  int depth = 0, max_depth = 0;

  Parser(const std::string &s_, const std::map<std::string, double> &c) : s(s_), constants(c) {}

  [[noreturn]] void fail(const std::string &why) const {
    throw ParseError("expression '" + s + "': " + why + " at offset " + std::to_string(pos));
  }
  void skip() { while (pos < s.size() && std::isspace((unsigned char)s[pos])) ++pos; }
  bool sym(char c) {
    skip();
    if (pos < s.size() && s[pos] == c) { ++pos; return true; }
    return false;
  }
  void emit(int op, int arg = 0, double val = 0.0, int delta = 0) {
    prog.push_back({op, arg, val});
    depth += delta;
    if (depth > max_depth) max_depth = depth;
  }
  void sum() {
    prod();
    for (;;) {
      if (sym('+')) { prod(); emit(OP_ADD, 0, 0, -1); }
      else if (sym('-')) { prod(); emit(OP_SUB, 0, 0, -1); }
      else return;
    }
  }
  void prod() {
    unary();
    for (;;) {
      if (sym('*')) { unary(); emit(OP_MUL, 0, 0, -1); }
      else if (sym('/')) { unary(); emit(OP_DIV, 0, 0, -1); }
      else return;
    }
  }
  void unary() {
    if (sym('-')) { unary(); emit(OP_NEG); }
    else if (sym('+')) { unary(); }
    else power();
  }
  void power() {
    atom();
    if (sym('^')) { unary_pow(); emit(OP_POW, 0, 0, -1); }
  }
  void unary_pow() {
    if (sym('-')) { unary_pow(); emit(OP_NEG); }
    else if (sym('+')) { unary_pow(); }
    else power();
  }
  void atom() {
    skip();
    if (pos >= s.size()) fail("unexpected end");
    const char c = s[pos];
    if (std::isdigit((unsigned char)c) || c == '.') {
      char *end = nullptr;
      const double v = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) fail("bad number");
      pos = size_t(end - s.c_str());
      emit(OP_CONST, 0, v, +1);
      return;
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      size_t b = pos;
      while (pos < s.size() && (std::isalnum((unsigned char)s[pos]) || s[pos] == '_')) ++pos;
      const std::string name = s.substr(b, pos - b);
      if (sym('(')) {
        int nargs = 1;
        sum();
        while (sym(',')) { sum(); ++nargs; }
        if (!sym(')')) fail("missing ')'");
        static const std::map<std::string, int> f1 = {
            {"sin", OP_SIN}, {"cos", OP_COS}, {"tan", OP_TAN}, {"asin", OP_ASIN}, {"acos", OP_ACOS},
            {"atan", OP_ATAN}, {"sinh", OP_SINH}, {"cosh", OP_COSH}, {"tanh", OP_TANH}, {"exp", OP_EXP},
            {"log", OP_LOG}, {"ln", OP_LOG}, {"log10", OP_LOG10}, {"log2", OP_LOG2}, {"sqrt", OP_SQRT},
            {"abs", OP_ABS}, {"sign", OP_SIGN}, {"floor", OP_FLOOR}, {"ceil", OP_CEIL}};
        static const std::map<std::string, int> f2 = {{"pow", OP_POW}, {"min", OP_MIN}, {"max", OP_MAX}};
        if (nargs == 1 && f1.count(name)) emit(f1.at(name));
        else if (nargs == 2 && f2.count(name)) emit(f2.at(name), 0, 0, -1);
        else fail("unknown function " + name);
        return;
      }
      if (name == "x") { emit(OP_VAR, 0, 0, +1); return; }
      if (name == "y") { emit(OP_VAR, 1, 0, +1); return; }
      if (name == "z") { emit(OP_VAR, 2, 0, +1); return; }
      auto it = constants.find(name);
      if (it != constants.end()) { emit(OP_CONST, 0, it->second, +1); return; }
      if (name == "pi" || name == "Pi" || name == "_pi") { emit(OP_CONST, 0, M_PI, +1); return; }
      if (name == "_e") { emit(OP_CONST, 0, M_E, +1); return; }
      fail("unknown identifier " + name);
    }
    if (c == '(') {
      ++pos;
      sum();
      if (!sym(')')) fail("missing ')'");
      return;
    }
    fail(std::string("unexpected character '") + c + "'");
  }
};

std::string trim(const std::string &s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace((unsigned char)s[b])) ++b;
  while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

}  // namespace

std::vector<ExprInstr> expr_compile(const std::string &text, const std::map<std::string, double> &constants) {
  Parser p(text, constants);
  p.sum();
  p.skip();
  if (p.pos != text.size()) p.fail("trailing characters");
  if (p.max_depth > kExprMaxStack) p.fail("expression too deep");
  return p.prog;
}

std::map<std::string, double> expr_parse_constants(const std::string &s) {
  std::map<std::string, double> out;
  size_t b = 0;
  while (b <= s.size()) {
    size_t e = s.find(',', b);
    if (e == std::string::npos) e = s.size();
    const std::string item = trim(s.substr(b, e - b));
    if (!item.empty()) {
      const size_t eq = item.find('=');
      if (eq == std::string::npos) throw ParseError("bad constant '" + item + "'");
      out[trim(item.substr(0, eq))] = std::strtod(item.c_str() + eq + 1, nullptr);
    }
    b = e + 1;
  }
  return out;
}

std::vector<std::string> expr_split_components(const std::string &s) {
  std::vector<std::string> out;
  size_t b = 0;
  for (;;) {
    size_t e = s.find(';', b);
    if (e == std::string::npos) { out.push_back(trim(s.substr(b))); break; }
    out.push_back(trim(s.substr(b, e - b)));
    b = e + 1;
  }
  return out;
}

}  // namespace msfec
