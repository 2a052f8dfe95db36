#pragma once
#include <deal.II/base/shim_common.h>
namespace dealii {
template <typename Number>
class Vector {
 public:
  Vector() = default;
  explicit Vector(std::size_t n) : v_(n, Number(0)) {}
  Number &operator()(std::size_t i) { return v_[i]; }
  const Number &operator()(std::size_t i) const { return v_[i]; }
  Vector &operator=(Number s) { for (auto &x : v_) x = s; return *this; }
  Vector &operator/=(Number s) { for (auto &x : v_) x /= s; return *this; }
  std::size_t size() const { return v_.size(); }
 private:
  std::vector<Number> v_;
};
}  // namespace dealii
