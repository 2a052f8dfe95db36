// Reader for deal.II ParameterHandler text files (.prm), the subset the reference
// uses: "subsection NAME" ... "end", "set KEY = VALUE", '#' comments, trailing '\'
// line continuation.  Mirrors the skip_undefined=true re-parsing that every consumer
// of the reference performs (source/Ned_RT/ned_rt_parameters.cc:35-38,153-156,
// source/equation_data/eqn_coeff_A.cc:17-23): unknown keys are kept and ignored.
#pragma once
#include <map>
#include <string>

namespace msfec {

class PrmFile {
 public:
  // Throws std::runtime_error on I/O or syntax errors.
  static PrmFile parse_file(const std::string &path);
  static PrmFile parse_text(const std::string &text);

  // path = "Sub A/Sub B/key"; returns dflt if absent.
  std::string get(const std::string &path, const std::string &dflt) const;
  bool has(const std::string &path) const;
  long get_integer(const std::string &path, long dflt, long lo, long hi) const;
  double get_double(const std::string &path, double dflt, double lo, double hi) const;
  bool get_bool(const std::string &path, bool dflt) const;

 private:
  std::map<std::string, std::string> kv_;   // flattened "a/b/key" -> value
};

}  // namespace msfec
