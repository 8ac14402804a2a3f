#pragma once
namespace polatory::fmm {
template <class Rbf>
struct GradientKernel;   // the ScalFMM matrix-kernel adaptor: only named by the interface headers
}  // namespace polatory::fmm
