#pragma once
#include <polatory/types.hpp>
namespace polatory::geometry {
template <int Dim>
using Point = Mat<1, Dim>;
template <int Dim>
using Points = Mat<Eigen::Dynamic, Dim>;   // row-major N x Dim, contiguous
}  // namespace polatory::geometry
