#pragma once

#include "common.cuh"
#include "rbf.cuh"

namespace plt {

struct DirectArgs {
  RbfConst k;
  const double* spos;  // SoA [dim][ns], isotropic space
  const double* swt;   // SoA [km][ns], anisotropy folded in
  int64_t ns;
  const double* tpos;  // SoA [dim][nt]
  int64_t nt;
  double* out;         // SoA [kn][nt] raw (pre-output-transform) sums
  double* partial;     // [n_chunks][kn][nt] scratch when n_chunks > 1
  int n_chunks;
  int64_t chunk;       // filled by launch_direct
  int symmetric;       // skip source index == target index
};

// Number of source chunks (blockIdx.y) that fills the GPU for this shape.
int direct_plan_chunks(int64_t ns, int64_t nt);

void launch_direct(int kind, int dim, DirectArgs a, cudaStream_t stream, LaunchCounter& ctr);

// Batched Gram matrices of the value kernel over B padded point sets (row-major [B][m][dim], original
// coordinates; aniso row-major dim x dim): the RAS domain matrices (preconditioner/mat_a.hpp).
void launch_gram_batched(int dim, const RbfConst& k, const double* aniso, const double* pts, const int* counts,
                         int64_t n_batch, int m, double nugget, double* out, cudaStream_t stream, LaunchCounter& ctr);

// The same with mixed value / gradient rows: types[b][r] = 0 value, 1 + c gradient component c, < 0 padding.
void launch_gram_mixed(int dim, const RbfConst& k, const double* aniso, const double* pts, const signed char* types,
                       int64_t n_batch, int m, double nugget, double* out, cudaStream_t stream, LaunchCounter& ctr);

}  // namespace plt
