// M2L by parent blocks, 3-D (declared in fmm_ops.cuh, where the method is described).
//
// Replaces most of the work of the reference's `fmm(src_tree, trg_tree, op, m2l)` pass
// (src/fmm/fmm_evaluator.hpp:92-99, src/fmm/fmm_symmetric_evaluator.hpp:81-82) for the uniform (equispaced-node)
// interpolator ScalFMM applies through FFTs: same nodes, same kernel values, same sums, regrouped.
//
//   upward    k_mblk3        children's multipoles -> one (2p-1)^3 grid per parent -> half spectrum of length 4p-3 per axis
//   M2L       k_m2l_blk3     Lhat_blk[target parent] += Kblk[source parent - target parent] . Mblk[source parent]
//   downward  k_idft_blk3    pruned inverse DFT back to the (2p-1)^3 grid, added to the local expansions of the children
//
// The adjacent child pairs of different parents that the block sums contain are taken out again by the same Hadamard
// kernel run one level down with the negated operators of the 3^3 near offsets (k_m2l_grouped3<.., NEAR = true>).
#include <cmath>
#include <cstdlib>

#include "fmm_ops.cuh"

namespace plt {
namespace {

constexpr int kBlkMaxOrder = 8;
struct TwBlk {
  double2 w[4 * kBlkMaxOrder - 3];  // forward twiddles (cos, -sin) of length 4 order - 3
};

TwBlk make_tw_blk(int order) {
  TwBlk t{};
  const int nb = blk_nf(order);
  const double pi = 3.14159265358979323846264338327950288;
  for (int i = 0; i < nb; ++i) {
    const double th = 2.0 * pi * i / nb;
    t.w[i] = make_double2(std::cos(th), -std::sin(th));
  }
  return t;
}

__device__ __forceinline__ int level_of_cell(const TreeView& tr, int g, int last) {
  int l = 1;
  while (l < last && g >= tr.cell_off[l + 1]) ++l;
  return l;
}

// ------------------------------------------------------------------------------------
// Upward: block spectrum of one source parent and weight component per CTA pass.
//   G[g0][g1][g2] = sum of the children's M at the grid node (a node on a shared plane belongs to 2, 4 or 8 children)
//   stage A: axis 2, 2p-1 real -> 2p-1 complex (half spectrum);  B: axis 1, 2p-1 -> 4p-3;  C: axis 0, 2p-1 -> 4p-3
// Zero padding from 2p-1 to 4p-3 points is implicit.  Register-blocked like k_m2hat3: a thread owns a column, the
// twiddles are kernel-parameter constants.
// ------------------------------------------------------------------------------------
template <int ORDER>
__global__ void __launch_bounds__(256) k_mblk3(TreeView src, int km, int first_cell, int n_par,
                                               const double* __restrict__ M, double2* __restrict__ Mblk, TwBlk tw) {
  constexpr int p = ORDER, n = 2 * p - 1, NB = 4 * p - 3, P = p * p * p;
  constexpr size_t FB = static_cast<size_t>(NB) * NB * n;
  extern __shared__ double2 sm2[];
  double2* Y1 = sm2;                           // [n][n][n]   (g0, g1, k2)
  double2* Y2 = Y1 + n * n * n;                // [n][NB][n]  (g0, k1, k2)
  double* G = reinterpret_cast<double*>(Y2);   // [n][n][n] real; dead before stage B writes Y2
  __shared__ int s_child[8];
  __shared__ int s_on;
  const int leaf = src.height - 1;
  for (int w = blockIdx.x; w < n_par * km; w += gridDim.x) {
    const int cell = first_cell + w / km, comp = w % km;  // cell: global id - cell_off[1]
    __syncthreads();  // previous pass done with the buffers and the child table
    if (threadIdx.x < 8) {
      const int g = src.cell_off[1] + cell;
      const int l = level_of_cell(src, g, leaf - 1);
      const int ci = src.dense[src.dense_off[l + 1] + ((src.keys[g] << 3) | threadIdx.x)];
      s_child[threadIdx.x] = ci >= 0 ? src.cell_off[l + 1] + ci : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int on = src.flags == nullptr;
      if (!on)  // partitioned upward pass: a block is needed where the spectra of its children are
        for (int c = 0; c < 8; ++c) on = on || (s_child[c] >= 0 && (src.flags[s_child[c]] & kCellFlagMhat));
      s_on = on;
    }
    __syncthreads();
    if (!s_on) continue;
    for (int e = threadIdx.x; e < n * n * n; e += blockDim.x) {
      const int g3[3] = {e / (n * n), (e / n) % n, e % n};
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int gid = s_child[c];
        bool ok = gid >= 0;
        int idx = 0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const int i = g3[x] - ((c >> (2 - x)) & 1) * (p - 1);
          ok = ok && i >= 0 && i < p;
          idx = idx * p + i;
        }
        if (ok) v += M[(static_cast<size_t>(gid) * km + comp) * P + idx];
      }
      G[e] = v;
    }
    __syncthreads();
    // stage A: column = (g0, g1)
    for (int col = threadIdx.x; col < n * n; col += blockDim.x) {
      double x[n];
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = G[col * n + i];
#pragma unroll
      for (int k = 0; k < n; ++k) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double2 t = tw.w[(k * i) % NB];
          re = fma(x[i], t.x, re);
          im = fma(x[i], t.y, im);
        }
        Y1[col * n + k] = make_double2(re, im);
      }
    }
    __syncthreads();
    // stage B: column = (g0, k2), stride n
    for (int col = threadIdx.x; col < n * n; col += blockDim.x) {
      const int g0 = col / n, k2 = col % n;
      double2 x[n];
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = Y1[(g0 * n + i) * n + k2];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double2 t = tw.w[(k * i) % NB];
          re = fma(x[i].x, t.x, re);
          re = fma(-x[i].y, t.y, re);
          im = fma(x[i].x, t.y, im);
          im = fma(x[i].y, t.x, im);
        }
        Y2[(g0 * NB + k) * n + k2] = make_double2(re, im);
      }
    }
    __syncthreads();
    // stage C: column = (k1, k2), stride NB * n
    double2* out = Mblk + (static_cast<size_t>(cell) * km + comp) * FB;
    for (int col = threadIdx.x; col < NB * n; col += blockDim.x) {
      double2 x[n];
#pragma unroll
      for (int i = 0; i < n; ++i) x[i] = Y2[i * (NB * n) + col];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double2 t = tw.w[(k * i) % NB];
          re = fma(x[i].x, t.x, re);
          re = fma(-x[i].y, t.y, re);
          im = fma(x[i].x, t.y, im);
          im = fma(x[i].y, t.x, im);
        }
        out[static_cast<size_t>(k) * (NB * n) + col] = make_double2(re, im);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Hadamard accumulation with the operators in registers (block sums and near-pair correction).
//
// Both passes have the same shape: 8 targets at the positions {1, 2}^3 of a 4^3 block of same-size cells, sources at
// any of the 64 positions, and an operator that depends only on the offset D = source - target in {-1, 0, 1}^3 \ 0:
//   block sums  (NEAR = false): targets = a sibling GROUP of active target parents (plan.cuh), sources = the parents
//               around it, spectra Mblk (FB frequencies), operators Kblk, result Lhat_blk;
//   correction  (NEAR = true):  targets = the children of ONE active parent, sources = the children of the
//               neighbouring parents that touch them (the parent's own children are left out), spectra Mhat
//               (F frequencies), operators = the negated near entries of Khat, result Lhat.
// A warp = 32 consecutive frequencies.  It keeps the operator values of all offsets at its frequencies in registers
// (13 complex values for the scalar kinds: K[-D] = conj K[D] for an even real kernel; 26 for the vector kinds) and the
// 8 targets as 8 complex accumulators.  The rows of the present source positions are compacted into a pointer list and
// stream, in that order, through a per-warp shared-memory ring filled by cp.async (kGrpDepth rows in flight per warp,
// no staging registers); the arithmetic is straight-line code over the 64 positions with one warp-uniform branch per
// position, so every offset is a compile-time constant: no shared-memory operand, 1 .. 8 complex FMAs per row (3.25
// on average).
//
// The order of the sum for a target is the source position ascending: it does not depend on which other targets
// exist, so sharded evaluations stay bit-identical to the single-GPU one.
// ------------------------------------------------------------------------------------
#ifndef PLT_GRP_WARPS
#define PLT_GRP_WARPS 8
#endif
constexpr int kGrpWarps = PLT_GRP_WARPS;
constexpr int kGrpTF = 32;
constexpr int kGrpRing = 8, kGrpDepth = kGrpRing - 1;  // ring slots per warp; rows in flight

__device__ __forceinline__ void cfma_n(double2& acc, const double2& k, const double2& m) {  // acc += k * m
  acc.x = fma(k.x, m.x, acc.x);
  acc.x = fma(-k.y, m.y, acc.x);
  acc.y = fma(k.x, m.y, acc.y);
  acc.y = fma(k.y, m.x, acc.y);
}
__device__ __forceinline__ void cfma_c(double2& acc, const double2& k, const double2& m) {  // acc += conj(k) * m
  acc.x = fma(k.x, m.x, acc.x);
  acc.x = fma(k.y, m.y, acc.x);
  acc.y = fma(k.x, m.y, acc.y);
  acc.y = fma(-k.y, m.x, acc.y);
}

// All complex FMAs of the source position sp = (s0 * 4 + s1) * 4 + s2 (a compile-time constant after unrolling).
template <bool VEC, int NK>
__device__ __forceinline__ void grp_body(int sp, double2 (&acc)[8], const double2 (&kk)[NK], const double2& m) {
  const int s0 = sp >> 4, s1 = (sp >> 2) & 3, s2 = sp & 3;
#pragma unroll
  for (int tp = 0; tp < 8; ++tp) {
    const int d0 = s0 - (((tp >> 2) & 1) + 1), d1 = s1 - (((tp >> 1) & 1) + 1), d2 = s2 - ((tp & 1) + 1);
    if (d0 < -1 || d0 > 1 || d1 < -1 || d1 > 1 || d2 < -1 || d2 > 1) continue;
    const int di = ((d0 + 1) * 3 + (d1 + 1)) * 3 + (d2 + 1);
    if (di == 13) continue;
    if (VEC) {
      cfma_n(acc[tp], kk[di % NK], m);
    } else if (di < 13) {
      cfma_n(acc[tp], kk[di % NK], m);
    } else {
      cfma_c(acc[tp], kk[(26 - di) % NK], m);
    }
  }
}

template <bool VEC, bool NEAR>
__global__ void __launch_bounds__(kGrpWarps * 32, VEC ? 1 : 2) k_m2l_grouped3(M2LArgs a, int F, int n_ftiles) {
  constexpr int NK = VEC ? 27 : 13;
  __shared__ __align__(512) double2 s_ring[kGrpWarps][kGrpRing][32];
  __shared__ const double2* s_ptr[kGrpWarps][64 + kGrpRing];
  __shared__ unsigned long long s_need[256];
  __shared__ int s_g[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    if (NEAR) {
      s_g[0] = 0;
      s_g[1] = a.n_active;
    } else {
      // groups that intersect the chunk of slots [slot0, slot0 + n_active)
      int l = a.grp_lo, h = a.grp_hi;
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (a.grp_first[mid] <= a.slot0) l = mid + 1; else h = mid;
      }
      s_g[0] = max(a.grp_lo, l - 1);
      l = a.grp_lo, h = a.grp_hi;
      const int end = a.slot0 + a.n_active;
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (a.grp_first[mid] < end) l = mid + 1; else h = mid;
      }
      s_g[1] = l;
    }
  }
  // s_need[target mask]: source positions within one cell of a present target (and not on it)
  for (int m = threadIdx.x; m < 256; m += blockDim.x) {
    unsigned long long need = 0;
#pragma unroll 1
    for (int sp = 0; sp < 64; ++sp) {
      bool any = false;
#pragma unroll 1
      for (int tp = 0; tp < 8; ++tp) {
        if (!((m >> tp) & 1)) continue;
        const int d0 = (sp >> 4) - (((tp >> 2) & 1) + 1), d1 = ((sp >> 2) & 3) - (((tp >> 1) & 1) + 1),
                  d2 = (sp & 3) - ((tp & 1) + 1);
        any = any || (abs(d0) <= 1 && abs(d1) <= 1 && abs(d2) <= 1 && (d0 | d1 | d2) != 0);
      }
      if (any) need |= 1ull << sp;
    }
    s_need[m] = need;
  }
  __syncthreads();
  const int glo = s_g[0], n_g = s_g[1] - s_g[0];
  if (n_g <= 0) return;

  // Schedule: the warps of a CTA take the SAME group at kGrpWarps neighbouring frequency tiles (identical control flow:
  // they walk through the ~40 KB of straight-line position code together and share its instruction-cache lines; 4 KB
  // contiguous of every spectrum row per CTA), and the whole grid works on ONE block of kGrpWarps tiles at a time, CTA c
  // on the groups c, c + gridDim.x, ...: the CTAs in flight then cover a window of consecutive (Morton-neighbouring)
  // groups, whose source rows overlap and stay in L2.  With contiguous (tile block, group) ranges per CTA the rows
  // were re-read from DRAM 4 - 5 times (ncu: 13.4 GB against 2.9 GB of unique block spectra on a dense 1M-point cloud).
  const int n_tb = (n_ftiles + kGrpWarps - 1) / kGrpWarps;
  const double2** ptrs = s_ptr[warp];
  const double2* ring = &s_ring[warp][0][lane];
  const unsigned ring_addr = static_cast<unsigned>(__cvta_generic_to_shared(ring));
  const unsigned ptrs_addr = static_cast<unsigned>(__cvta_generic_to_shared(ptrs));
  const int kn = VEC ? a.kn : 1, km = VEC ? a.km : 1;
  const double2* spectra = NEAR ? a.Mhat : a.Mblk;
  const double2* ops = NEAR ? a.Khat : a.Kblk;
  double2* result = NEAR ? a.Lhat : a.Lhat_blk;
  for (int tb = 0; tb < n_tb; ++tb) {
    const int g_lo = blockIdx.x, g_hi = n_g;
    const int ftile = tb * kGrpWarps + warp;
    if (ftile >= n_ftiles) continue;  // last tile block: no CTA-wide barrier below
    const int f = ftile * kGrpTF + lane;
    const bool fok = f < F;
    const int fl = fok ? f : F - 1;  // idle lanes of the last tile compute on a valid address and store nothing
    for (int comp = 0; comp < kn * km; ++comp) {
      const int cb = comp / km, ca = comp - cb * km;
      double2 kk[NK];
#pragma unroll
      for (int d = 0; d < NK; ++d) {
        // operator index of offset d: the block table is indexed by the 3^3 offsets, the child table by the 7^3 ones
        const int oi = NEAR ? ((d / 9 + 2) * 7 + ((d / 3) % 3 + 2)) * 7 + (d % 3 + 2) : d;
        kk[d] = ops[(static_cast<size_t>(oi) * kn * km + comp) * F + fl];
      }
      const double2* rbase = spectra + static_cast<size_t>(ca) * F;
      const size_t rstride = static_cast<size_t>(km) * F;
      const unsigned long long lane_off = static_cast<unsigned long long>(fl) * sizeof(double2);

      for (int g = g_lo; g < g_hi; g += gridDim.x) {
        // ---- targets and source rows of this group ----
        int local = -1;     // lane tp < 8: index of target tp in the result (or -1)
        int r0, r1;         // spectrum rows of the source positions lane and lane + 32 (or -1)
        if (NEAR) {
          const unsigned tm = a.trg_mask[g];
          if (lane < 8 && ((tm >> lane) & 1u)) local = g * 8 + lane;
          const int* tab = a.src_ids + static_cast<size_t>(g) * 216;
          auto row_of = [&](int sp) {
            // position -> (neighbour parent e, child c): v = s - 1 in [-1, 2] is the child coordinate relative to the parent
            int nb = 0, cs = 0;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
              const int v = ((sp >> (4 - 2 * x)) & 3) - 1;
              const int e = v < 0 ? -1 : (v > 1 ? 1 : 0);
              nb = nb * 3 + e + 1;
              cs = cs * 2 + (v - 2 * e);
            }
            return nb == 13 ? -1 : tab[nb * 8 + cs];
          };
          r0 = row_of(lane);
          r1 = row_of(lane + 32);
        } else {
          const int gg = glo + g;
          if (lane < 8) {
            const int slot = a.grp_slot[static_cast<size_t>(gg) * 8 + lane];
            if (slot >= a.slot0 && slot < a.slot0 + a.n_active) local = slot - a.slot0;
          }
          r0 = a.grp_src[static_cast<size_t>(gg) * 64 + lane];
          r1 = a.grp_src[static_cast<size_t>(gg) * 64 + 32 + lane];
        }
        const unsigned tmask = __ballot_sync(0xffffffffu, local >= 0) & 0xffu;
        const unsigned long long need = s_need[tmask];
        const unsigned b0 = __ballot_sync(0xffffffffu, r0 >= 0) & static_cast<unsigned>(need);
        const unsigned b1 = __ballot_sync(0xffffffffu, r1 >= 0) & static_cast<unsigned>(need >> 32);
        const unsigned lt = (1u << lane) - 1u;
        __syncwarp();  // previous group done with the pointer list
        // (row starts: the list is read by all lanes, each adds its own frequency)
        if ((b0 >> lane) & 1u) ptrs[__popc(b0 & lt)] = rbase + static_cast<size_t>(r0) * rstride;
        if ((b1 >> lane) & 1u) ptrs[__popc(b0) + __popc(b1 & lt)] = rbase + static_cast<size_t>(r1) * rstride;
        const int n = __popc(b0) + __popc(b1);
        __syncwarp();

        double2 acc[8];
#pragma unroll
        for (int tp = 0; tp < 8; ++tp) {
          acc[tp] = make_double2(0.0, 0.0);
          if (VEC && ca > 0) {
            const int sl = __shfl_sync(0xffffffffu, local, tp);
            if (sl >= 0 && fok) acc[tp] = result[(static_cast<size_t>(sl) * kn + cb) * F + f];
          }
        }

        // ---- rows through the cp.async ring: row i of the list goes to ring slot i mod kGrpRing; one commit group per
        // list position (empty past the end), so "at most kGrpDepth - 1 groups pending" means row `cnt` has landed ----
        auto issue = [&](int i) {
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              ".reg .b64 src;\n\t"
              "setp.lt.s32 p, %0, %1;\n\t"
              "ld.shared.b64 src, [%2];\n\t"
              "add.u64 src, src, %4;\n\t"
              "@p cp.async.ca.shared.global [%3], [src], 16;\n\t"
              "cp.async.commit_group;\n\t"
              "}" ::"r"(i), "r"(n), "r"(ptrs_addr + 8u * i), "r"(ring_addr + ((i & (kGrpRing - 1)) << 9)), "l"(lane_off)
              : "memory");
        };
#pragma unroll
        for (int i = 0; i < kGrpDepth; ++i) issue(i);
        int cnt = 0;
        // the 64 source positions in ascending order, straight-line code: a warp-uniform branch per position, every offset
        // a compile-time constant
#pragma unroll
        for (int sp = 0; sp < 64; ++sp) {
          if (NEAR && ((sp >> 4) == 1 || (sp >> 4) == 2) && (((sp >> 2) & 3) == 1 || ((sp >> 2) & 3) == 2) &&
              ((sp & 3) == 1 || (sp & 3) == 2))
            continue;  // the parent's own children are never sources of the correction pass
          if (((sp < 32 ? b0 : b1) >> (sp & 31)) & 1u) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kGrpDepth - 1) : "memory");
            const double2 m = ring[(cnt & (kGrpRing - 1)) * 32];
            issue(cnt + kGrpDepth);  // slot (cnt + kGrpDepth) mod kGrpRing: not the one just read
            ++cnt;
            grp_body<VEC, NK>(sp, acc, kk, m);
          }
        }
#pragma unroll
        for (int tp = 0; tp < 8; ++tp) {
          const int sl = __shfl_sync(0xffffffffu, local, tp);
          if (sl >= 0 && fok) __stcs(&result[(static_cast<size_t>(sl) * kn + cb) * F + f], acc[tp]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Downward: pruned inverse DFT of one block spectrum per CTA, added to the children's expansions.
//   stage A: axis 0, 4p-3 -> 2p-1 (inputs streamed from global memory, 2p-1 accumulators per column in registers);
//   B: axis 1, 4p-3 -> 2p-1;  C: axis 2, half spectrum -> real, then scatter: grid node g of the block is node
//   g - c (p - 1) of child c for every child c that contains it.
// ------------------------------------------------------------------------------------
template <int ORDER>
__global__ void __launch_bounds__(256) k_idft_blk3(M2LArgs a, TwBlk tw) {
  constexpr int p = ORDER, n = 2 * p - 1, NB = 4 * p - 3, P = p * p * p;
  constexpr size_t FB = static_cast<size_t>(NB) * NB * n;
  extern __shared__ double2 sm2[];
  double2* Y = sm2;                // [n][NB][n]  (m0, k1, k2)
  double2* Z = Y + n * NB * n;     // [n][n][n]   (m0, m1, k2)
  __shared__ double* s_out[8];
  const int slot = blockIdx.x / a.kn, b = blockIdx.x % a.kn;
  if (threadIdx.x < 8) {
    const int ct = threadIdx.x;
    double* out = nullptr;
    if ((a.trg_mask[slot] >> ct) & 1u) {
      if (a.L) {
        const int pidx = a.active[slot];
        const uint32_t pkey = a.trg.keys[a.trg.cell_off[a.level - 1] + pidx];
        const int cidx = a.trg.dense[a.trg.dense_off[a.level] + ((pkey << 3) | ct)];
        out = a.L + (static_cast<size_t>(a.trg.cell_off[a.level] + cidx) * a.kn + b) * P;
      } else {
        out = a.Lc + ((static_cast<size_t>(slot) * 8 + ct) * a.kn + b) * P;
      }
    }
    s_out[ct] = out;
  }
  const double2* in = a.Lhat_blk + (static_cast<size_t>(slot) * a.kn + b) * FB;
  // stage A: column = (k1, k2)
  for (int col = threadIdx.x; col < NB * n; col += blockDim.x) {
    double2 acc[n];
#pragma unroll
    for (int m = 0; m < n; ++m) acc[m] = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const double2 x = __ldcs(in + static_cast<size_t>(k) * (NB * n) + col);
#pragma unroll
      for (int m = 0; m < n; ++m) {
        const double2 t = tw.w[(m * k) % NB];  // conj: e^{+i theta}
        acc[m].x = fma(x.x, t.x, acc[m].x);
        acc[m].x = fma(x.y, t.y, acc[m].x);
        acc[m].y = fma(x.y, t.x, acc[m].y);
        acc[m].y = fma(-x.x, t.y, acc[m].y);
      }
    }
#pragma unroll
    for (int m = 0; m < n; ++m) Y[m * (NB * n) + col] = acc[m];
  }
  __syncthreads();
  // stage B: column = (m0, k2)
  for (int col = threadIdx.x; col < n * n; col += blockDim.x) {
    const int m0 = col / n, k2 = col % n;
    double2 acc[n];
#pragma unroll
    for (int m = 0; m < n; ++m) acc[m] = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const double2 x = Y[(m0 * NB + k) * n + k2];
#pragma unroll
      for (int m = 0; m < n; ++m) {
        const double2 t = tw.w[(m * k) % NB];
        acc[m].x = fma(x.x, t.x, acc[m].x);
        acc[m].x = fma(x.y, t.y, acc[m].x);
        acc[m].y = fma(x.y, t.x, acc[m].y);
        acc[m].y = fma(-x.x, t.y, acc[m].y);
      }
    }
#pragma unroll
    for (int m = 0; m < n; ++m) Z[(m0 * n + m) * n + k2] = acc[m];
  }
  __syncthreads();
  // stage C: column = (m0, m1)
  for (int col = threadIdx.x; col < n * n; col += blockDim.x) {
    const int m0 = col / n, m1 = col % n;
    double2 x[n];
#pragma unroll
    for (int k = 0; k < n; ++k) x[k] = Z[col * n + k];
#pragma unroll
    for (int m2 = 0; m2 < n; ++m2) {
      double v = x[0].x;
#pragma unroll
      for (int k = 1; k < n; ++k) {
        const double2 t = tw.w[(k * m2) % NB];
        v = fma(2.0 * x[k].x, t.x, v);
        v = fma(2.0 * x[k].y, t.y, v);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int i0 = m0 - ((c >> 2) & 1) * (p - 1), i1 = m1 - ((c >> 1) & 1) * (p - 1), i2 = m2 - (c & 1) * (p - 1);
        if (i2 < 0 || i2 >= p) continue;  // compile-time
        if (i0 < 0 || i0 >= p || i1 < 0 || i1 >= p) continue;
        double* out = s_out[c];
        if (out) out[(i0 * p + i1) * p + i2] += v;
      }
    }
  }
}

void smem_opt_in(const void* fn, size_t bytes) {
  if (bytes > 40 * 1024)  // static shared memory counts towards the 48 KiB default limit
    PLT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
}

__global__ void k_check_finite(const double* __restrict__ x, size_t n, int* __restrict__ flag) {
  bool bad = false;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    bad = bad || !isfinite(x[i]);
  if (bad) atomicOr(flag, 1);
}

template <int ORDER>
void launch_mblk_t(int km, const TreeView& src, int par_lo, int par_hi, const double* M, double2* Mblk, cudaStream_t s,
                   LaunchCounter& c) {
  constexpr int n = 2 * ORDER - 1, NB = 4 * ORDER - 3;
  const int first_cell = src.cell_off[par_lo] - src.cell_off[1];
  const int n_par = src.cell_off[par_hi + 1] - src.cell_off[par_lo];
  if (n_par <= 0) return;
  const size_t smem = sizeof(double2) * (n * n * n + n * NB * n);
  smem_opt_in((const void*)k_mblk3<ORDER>, smem);
  const int grid = static_cast<int>(std::min<long long>(static_cast<long long>(n_par) * km, num_sm() * 16));
  PLT_LAUNCH(c, (k_mblk3<ORDER>), grid, 256, smem, s, src, km, first_cell, n_par, M, Mblk, make_tw_blk(ORDER));
}

template <int ORDER>
void launch_idft_blk_t(const M2LArgs& a, cudaStream_t s, LaunchCounter& c) {
  constexpr int n = 2 * ORDER - 1, NB = 4 * ORDER - 3;
  const size_t smem = sizeof(double2) * (n * NB * n + n * n * n);
  smem_opt_in((const void*)k_idft_blk3<ORDER>, smem);
  PLT_LAUNCH(c, (k_idft_blk3<ORDER>), a.n_active * a.kn, 256, smem, s, a, make_tw_blk(ORDER));
}

}  // namespace

namespace {
double g_blk_min_fill = [] {
  const char* e = getenv("PLT_BLK_MIN_FILL");
  return e ? atof(e) : 0.25;
}();
}  // namespace
double blk_min_fill() { return g_blk_min_fill; }
double blk_set_min_fill(double v) {
  const double old = g_blk_min_fill;
  g_blk_min_fill = v;
  return old;
}

bool blk_supported(int dim, int order) {
  static const bool off = getenv("PLT_DEBUG_NO_BLK") != nullptr;  // A/B switch for parity bisection
  return !off && dim == 3 && (order == 6 || order == 8);
}

void launch_check_finite(const double* x, size_t n, int* flag, cudaStream_t s, LaunchCounter& c) {
  if (n == 0) return;
  PLT_LAUNCH(c, k_check_finite, static_cast<int>(std::min<size_t>((n + 255) / 256, num_sm() * 8)), 256, 0, s, x, n, flag);
}

void launch_mblk(int km, const TreeView& src, int order, int par_lo, int par_hi, const double* M, double2* Mblk,
                 cudaStream_t s, LaunchCounter& c) {
  if (src.height <= 2 || par_hi < par_lo) return;
  PLT_REQUIRE(par_lo >= 1 && par_hi <= src.height - 2, "block spectra: parent levels out of range");
  switch (order) {
    case 6: launch_mblk_t<6>(km, src, par_lo, par_hi, M, Mblk, s, c); return;
    case 8: launch_mblk_t<8>(km, src, par_lo, par_hi, M, Mblk, s, c); return;
    default: throw Error(PLT_ERR_INVALID, "parent-block M2L: unsupported order");
  }
}

namespace {
template <bool NEAR>
void launch_grouped(const M2LArgs& a, int F, int n_groups, cudaStream_t s, LaunchCounter& c) {
  const int n_ftiles = ceil_div(F, kGrpTF);
  if (a.kn * a.km == 1) {
    const int grid = std::max(1, std::min(2 * num_sm(), n_groups));
    PLT_LAUNCH(c, (k_m2l_grouped3<false, NEAR>), grid, kGrpWarps * 32, 0, s, a, F, n_ftiles);
  } else {
    const int grid = std::max(1, std::min(num_sm(), n_groups));
    PLT_LAUNCH(c, (k_m2l_grouped3<true, NEAR>), grid, kGrpWarps * 32, 0, s, a, F, n_ftiles);
  }
}
}  // namespace

void launch_m2l_blk_hadamard(const M2LArgs& a, cudaStream_t s, LaunchCounter& c) {
  if (a.n_active == 0 || a.grp_hi <= a.grp_lo) return;
  launch_grouped<false>(a, static_cast<int>(blk_freqs(a.order)), a.grp_hi - a.grp_lo, s, c);
}

void launch_m2l_hadamard_near(const M2LArgs& a, cudaStream_t s, LaunchCounter& c) {
  if (a.n_active == 0) return;
  PLT_REQUIRE(a.dim == 3, "near-pair correction pass: 3-D only");
  launch_grouped<true>(a, freqs_per_cell(a.order, a.dim), a.n_active, s, c);
}

void launch_m2l_blk_idft(const M2LArgs& a, cudaStream_t s, LaunchCounter& c) {
  if (a.n_active == 0) return;
  switch (a.order) {
    case 6: launch_idft_blk_t<6>(a, s, c); return;
    case 8: launch_idft_blk_t<8>(a, s, c); return;
    default: throw Error(PLT_ERR_INVALID, "parent-block M2L: unsupported order");
  }
}

}  // namespace plt
