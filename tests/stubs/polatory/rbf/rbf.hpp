#pragma once
#include <polatory/rbf/rbf_base.hpp>
#include <string>
#include <vector>
namespace polatory::rbf {
template <int Dim>
class Rbf {
 public:
  using Mat = polatory::Mat<Dim>;
  const Mat& anisotropy() const { return aniso_; }
  const std::vector<double>& parameters() const { return params_; }
  std::string short_name() const { return name_; }
 private:
  Mat aniso_;
  std::vector<double> params_;
  std::string name_;
};
}  // namespace polatory::rbf
