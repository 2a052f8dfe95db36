#pragma once
#include <deal.II/base/shim_common.h>
#include <istream>
#include <map>
#include <sstream>
namespace dealii {
namespace Patterns {
struct PatternBase { virtual ~PatternBase() = default; };
struct Integer : PatternBase { Integer(long = 0, long = 0) {} };
struct Double : PatternBase { Double(double = 0, double = 0) {} };
struct Bool : PatternBase {};
struct Anything : PatternBase {};
struct Selection : PatternBase { explicit Selection(const std::string &) {} };
}  // namespace Patterns

// The documented .prm grammar: `subsection <name>` / `end`, `set <key> = <value>`, `#` comments, a trailing
// backslash continues a line.  Undeclared entries are skipped when skip_undefined is set (the reference
// always sets it), otherwise they are an error.
class ParameterHandler {
 public:
  void enter_subsection(const std::string &s) { path_.push_back(s); }
  void leave_subsection() { path_.pop_back(); }
  void declare_entry(const std::string &key, const std::string &def,
                     const Patterns::PatternBase & = Patterns::Anything(), const std::string & = "") {
    entries_[full(key)] = def;
  }
  void parse_input(std::istream &in, const std::string & = "", const std::string &last_line = "",
                   const bool skip_undefined = false) {
    if (!in) throw std::runtime_error("ParameterHandler: cannot read the parameter file");
    const std::vector<std::string> saved = path_;
    std::string line, acc;
    while (std::getline(in, line)) {
      const std::size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      line = trim(line);
      if (!line.empty() && line.back() == '\\') { acc += line.substr(0, line.size() - 1); continue; }
      line = acc + line;
      acc.clear();
      if (line.empty()) continue;
      if (!last_line.empty() && line == last_line) break;
      if (starts(line, "subsection")) { path_.push_back(squeeze(trim(line.substr(10)))); continue; }
      if (line == "end") {
        if (path_.size() <= saved.size()) throw std::runtime_error("ParameterHandler: stray end");
        path_.pop_back();
        continue;
      }
      if (starts(line, "set")) {
        const std::size_t eq = line.find('=');
        if (eq == std::string::npos) throw std::runtime_error("ParameterHandler: set without =");
        const std::string key = squeeze(trim(line.substr(3, eq - 3))), val = trim(line.substr(eq + 1));
        auto it = entries_.find(full(key));
        if (it == entries_.end()) {
          if (skip_undefined) continue;
          throw std::runtime_error("ParameterHandler: undeclared entry " + key);
        }
        it->second = val;
        continue;
      }
      throw std::runtime_error("ParameterHandler: cannot parse line: " + line);
    }
    path_ = saved;
  }
  std::string get(const std::string &key) const {
    auto it = entries_.find(full(key));
    if (it == entries_.end()) throw std::runtime_error("ParameterHandler: entry not declared: " + key);
    return it->second;
  }
  long get_integer(const std::string &key) const { return std::stol(get(key)); }
  double get_double(const std::string &key) const { return std::stod(get(key)); }
  bool get_bool(const std::string &key) const {
    const std::string s = get(key);
    if (s == "true" || s == "yes" || s == "on") return true;
    if (s == "false" || s == "no" || s == "off") return false;
    throw std::runtime_error("ParameterHandler: not a bool: " + s);
  }

 private:
  static bool starts(const std::string &s, const std::string &w) {
    return s.compare(0, w.size(), w) == 0 && (s.size() == w.size() || s[w.size()] == ' ' || s[w.size()] == '\t');
  }
  static std::string trim(const std::string &s) {
    const std::size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  static std::string squeeze(const std::string &s) {  // runs of blanks -> one blank
    std::string o;
    bool sp = false;
    for (char c : s) {
      if (c == ' ' || c == '\t') { sp = true; continue; }
      if (sp && !o.empty()) o += ' ';
      sp = false;
      o += c;
    }
    return o;
  }
  std::string full(const std::string &key) const {
    std::string p;
    for (const auto &s : path_) p += s + "/";
    return p + key;
  }
  std::vector<std::string> path_;
  std::map<std::string, std::string> entries_;
};
}  // namespace dealii
