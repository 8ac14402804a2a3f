#pragma once
#include <polatory/geometry/point3d.hpp>
namespace polatory::geometry {
template <int Dim>
class Bbox {
 public:
  using Point = geometry::Point<Dim>;
  const Point& max() const { return max_; }
  const Point& min() const { return min_; }
 private:
  Point min_, max_;
};
}  // namespace polatory::geometry
