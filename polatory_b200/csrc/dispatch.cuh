// Runtime (family, kind, dim) -> template instantiation dispatch.
#pragma once

#include "common.cuh"
#include "rbf.cuh"

namespace plt {

// F is a generic lambda taking three std::integral_constant tags: family, kind, dim.
template <class F>
void dispatch_fkd(int family, int kind, int dim, F&& f) {
  auto with_dim = [&](auto fam, auto knd) {
    switch (dim) {
      case 1: f(fam, knd, std::integral_constant<int, 1>{}); break;
      case 2: f(fam, knd, std::integral_constant<int, 2>{}); break;
      case 3: f(fam, knd, std::integral_constant<int, 3>{}); break;
      default: throw Error(PLT_ERR_INVALID, "dim must be 1, 2 or 3");
    }
  };
  auto with_kind = [&](auto fam) {
    switch (kind) {
      case KIND_K: with_dim(fam, std::integral_constant<int, KIND_K>{}); break;
      case KIND_F: with_dim(fam, std::integral_constant<int, KIND_F>{}); break;
      case KIND_FT: with_dim(fam, std::integral_constant<int, KIND_FT>{}); break;
      case KIND_H: with_dim(fam, std::integral_constant<int, KIND_H>{}); break;
      default: throw Error(PLT_ERR_INVALID, "unknown kernel kind");
    }
  };
  switch (family) {
    case FAM_BH3: with_kind(std::integral_constant<int, FAM_BH3>{}); break;
    case FAM_TH3: with_kind(std::integral_constant<int, FAM_TH3>{}); break;
    case FAM_BH2: with_kind(std::integral_constant<int, FAM_BH2>{}); break;
    case FAM_TH2: with_kind(std::integral_constant<int, FAM_TH2>{}); break;
    case FAM_EXP: with_kind(std::integral_constant<int, FAM_EXP>{}); break;
    case FAM_GAU: with_kind(std::integral_constant<int, FAM_GAU>{}); break;
    case FAM_IMQ: with_kind(std::integral_constant<int, FAM_IMQ>{}); break;
    case FAM_SPD_FULL: with_kind(std::integral_constant<int, FAM_SPD_FULL>{}); break;
    case FAM_SPD_DIRECT: with_kind(std::integral_constant<int, FAM_SPD_DIRECT>{}); break;
    case FAM_SPH: with_kind(std::integral_constant<int, FAM_SPH>{}); break;
    case FAM_CUB: with_kind(std::integral_constant<int, FAM_CUB>{}); break;
    default: throw Error(PLT_ERR_INVALID, "unknown RBF family");
  }
}

}  // namespace plt
