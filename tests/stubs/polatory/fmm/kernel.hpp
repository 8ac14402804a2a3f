#pragma once
namespace polatory::fmm {
template <class Rbf>
struct Kernel;   // the ScalFMM matrix-kernel adaptor: only named by the interface headers
}  // namespace polatory::fmm
