// Interaction plan of one (source tree, target tree) pair, built on the device.
//
// Replaces what the reference rebuilds on every evaluate() with
//   scalfmm::list::sequential::build_m2l_interaction_list(src, trg, 1, 1)
//   scalfmm::list::sequential::build_p2p_interaction_list(src, trg, 1, mutual)
//   (src/fmm/fmm_evaluator.hpp:92-97, src/fmm/fmm_symmetric_evaluator.hpp:75-80).
// The lists themselves are implicit (dense key -> cell maps, separation criterion 1); the
// plan only compacts *which* target cells have work and resolves their source-cell ids once,
// so that the M2L and P2P kernels run over dense work lists.  A plan stays valid as long as
// both trees do: the matvec inside the Krylov solver builds it once per fit.
#pragma once

#include "common.cuh"
#include "tree.cuh"

namespace plt {

// Device view, passed by value to kernels.
struct PlanView {
  // M2L: active target parents, grouped by the level of their children.
  //   level l (2 <= l < height): slots [level_begin[l], level_begin[l + 1])
  const int* active;             // [n_active] compact id of the parent at level l - 1 (target tree)
  const int* src_ids;            // [n_active][3^dim * 2^dim] Mhat cell index (global id - cell_off[2]) or -1
  const unsigned char* trg_mask; // [n_active] bit ct set <=> target child ct exists
  const int* leaf_slot;          // [n_cells(leaf - 1)] slot - level_begin[leaf] of that parent, or -1
  // Per parent of the leaf level, everything the fused leaf pass needs in ONE coalesced read instead of a chain of
  // dependent lookups (key -> dense map -> leaf_start): [n_cells(leaf - 1)][2 + 3 * 2^dim] ints =
  //   { Morton key, leaf slot, then per child { compact leaf id or -1, first point, point count } }
  const int* leaf_meta;
  int level_begin[25];
  int n_active;
  // Sibling groups of active parents (3-D parent-block M2L, fmm_blk.cu): maximal runs of active parents with the same
  // parent (they are consecutive in Morton order).  Level l: groups [grp_level_begin[l], grp_level_begin[l + 1]).
  const int* grp_first = nullptr;  // [n_groups] first slot
  const int* grp_slot = nullptr;   // [n_groups][8] slot of the sibling at position tp = key & 7, or -1
  const int* grp_src = nullptr;    // [n_groups][64] source cell (global id - cell_off[1]) at position sp = (sx * 4 + sy) * 4 + sz
                                   // of the 4^3 block of cells around the group (s = coordinate - 2 * grandparent + 1), or -1
  int grp_level_begin[25];
  int n_groups = 0;
  // P2P: target leaves with at least one non-empty adjacent source leaf (ascending).
  const int* p2p_leaves;
  int n_p2p;
};

class Plan {
 public:
  void build(const Tree& src, const Tree& trg, cudaStream_t stream, LaunchCounter& ctr);
  bool built() const { return built_; }
  void reset() { built_ = false; }
  PlanView view() const { return view_; }
  int n_active(int level) const { return view_.level_begin[level + 1] - view_.level_begin[level]; }
  int n_p2p() const { return view_.n_p2p; }

 private:
  bool built_ = false;
  PlanView view_{};
  DevBuf<int> flags_, active_, src_ids_, leaf_slot_, leaf_meta_, p2p_flags_, p2p_leaves_, counts_;
  DevBuf<int> grp_flags_, grp_first_, grp_slot_, grp_src_;
  DevBuf<unsigned char> trg_mask_, tmp_;
};

}  // namespace plt
