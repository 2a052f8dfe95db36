#include "prm.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace msfec {
namespace {

std::string squeeze(const std::string &s) {   // trim + collapse inner whitespace
  std::string out;
  bool space = false;
  for (char c : s) {
    if (std::isspace((unsigned char)c)) { space = !out.empty(); continue; }
    if (space) out.push_back(' ');
    space = false;
    out.push_back(c);
  }
  return out;
}

std::string trim(const std::string &s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace((unsigned char)s[b])) ++b;
  while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

bool starts_with_word(const std::string &line, const char *word) {
  const size_t n = std::char_traits<char>::length(word);
  if (line.size() < n) return false;
  for (size_t i = 0; i < n; ++i)
    if (std::tolower((unsigned char)line[i]) != word[i]) return false;
  return line.size() == n || std::isspace((unsigned char)line[n]);
}

}  // namespace

PrmFile PrmFile::parse_file(const std::string &path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open parameter file '" + path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  return parse_text(ss.str());
}

PrmFile PrmFile::parse_text(const std::string &text) {
  PrmFile out;
  std::vector<std::string> stack;
  std::istringstream in(text);
  std::string raw, pending;
  int lineno = 0;
  while (std::getline(in, raw)) {
    ++lineno;
    const size_t hash = raw.find('#');
    std::string line = trim(hash == std::string::npos ? raw : raw.substr(0, hash));
    if (line.empty()) continue;
    if (line.back() == '\\') { pending += line.substr(0, line.size() - 1) + " "; continue; }
    line = trim(pending + line);
    pending.clear();
    if (starts_with_word(line, "subsection")) {
      stack.push_back(squeeze(line.substr(10)));
    } else if (starts_with_word(line, "end")) {
      if (stack.empty()) throw std::runtime_error("prm line " + std::to_string(lineno) + ": unbalanced 'end'");
      stack.pop_back();
    } else if (starts_with_word(line, "set")) {
      const size_t eq = line.find('=');
      if (eq == std::string::npos) throw std::runtime_error("prm line " + std::to_string(lineno) + ": missing '='");
      std::string key;
      for (const auto &s : stack) key += s + "/";
      key += squeeze(line.substr(3, eq - 3));
      out.kv_[key] = trim(line.substr(eq + 1));
    } else {
      throw std::runtime_error("prm line " + std::to_string(lineno) + ": cannot parse '" + raw + "'");
    }
  }
  if (!stack.empty()) throw std::runtime_error("prm: unterminated subsection '" + stack.back() + "'");
  return out;
}

bool PrmFile::has(const std::string &path) const { return kv_.count(path) != 0; }

std::string PrmFile::get(const std::string &path, const std::string &dflt) const {
  auto it = kv_.find(path);
  return it == kv_.end() ? dflt : it->second;
}

long PrmFile::get_integer(const std::string &path, long dflt, long lo, long hi) const {
  auto it = kv_.find(path);
  if (it == kv_.end()) return dflt;
  char *end = nullptr;
  const long v = std::strtol(it->second.c_str(), &end, 10);
  if (end == it->second.c_str() || *end != 0) throw std::runtime_error("prm: '" + path + "' is not an integer");
  if (v < lo || v > hi) throw std::runtime_error("prm: '" + path + "' out of range");
  return v;
}

double PrmFile::get_double(const std::string &path, double dflt, double lo, double hi) const {
  auto it = kv_.find(path);
  if (it == kv_.end()) return dflt;
  char *end = nullptr;
  const double v = std::strtod(it->second.c_str(), &end);
  if (end == it->second.c_str() || *end != 0) throw std::runtime_error("prm: '" + path + "' is not a number");
  if (v < lo || v > hi) throw std::runtime_error("prm: '" + path + "' out of range");
  return v;
}

bool PrmFile::get_bool(const std::string &path, bool dflt) const {
  auto it = kv_.find(path);
  if (it == kv_.end()) return dflt;
  std::string v;
  for (char c : it->second) v.push_back((char)std::tolower((unsigned char)c));
  if (v == "true" || v == "yes" || v == "on") return true;
  if (v == "false" || v == "no" || v == "off") return false;
  throw std::runtime_error("prm: '" + path + "' is not a bool");
}

}  // namespace msfec
