// Scalar 3-D Hadamard M2L with most operator reads served from Tensor Memory (sm_100a).
//
// The list kernel (k_m2l_hadamard_tiled, fmm_ops.cu) is bound by the 128 B/clk shared-memory pipe: one complex
// multiply-add needs one 16-byte operator value per lane = one LDS.128 = 4 SM-cycles per warp, against 2.06 cycles for
// its four DFMAs (tools/ubench/kpath.cu).  tcgen05.ld reads Tensor Memory on a path of its own (1.45 SM-cycles per
// 512-byte warp load, overlapping LDS), but a warp only reaches the 32 TMEM lanes of its own quarter: 512 columns x 4 B
// = 128 operator values per lane (lane = frequency of the tile), and the tile's operator slice has 316 far offsets.
// Two facts make 128 slots enough for most of the work:
//   * the scalar kernels are even, so Khat[-o] = conj(Khat[o]) (same frequency): half of the offsets are stored,
//     the other half is read with the sign of the imaginary part flipped (one integer XOR on the high word);
//   * of the 316 far offsets only those with at most ONE axis at +-3 are kept in TMEM: 248 offsets = 124 slots.
// A source position (one of the 6^3 child positions around the target parent) whose far children all have such
// offsets is a "T" entry (160 of the 216 positions, ~70 % of the pairs of a full table) and takes its operators from
// TMEM; the others ("S" entries: two or three axes at the rim of the neighbourhood) run the list kernel's inner step
// on the shared-memory slice.  The per-parent list is compacted into the two classes with ballots (T from the front,
// S from the back); T entries first, then S entries, both in ascending position order -- the order of the sum for a
// target does not depend on anything but its own source table (bit-identical shards and batches).
//
// The TMEM loads of one half entry (4 children, 16 registers) are in flight while the other half is multiplied:
//   ld(B: e.hi) | cfma(A: e.lo) | wait | ld(A: (e+1).lo) | cfma(B: e.hi) | wait
// The operator slice is staged in shared memory first (the S entries need it anyway) and mirrored from there into the
// four TMEM quarters with tcgen05.st by one warp per quarter.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fmm_ops.cuh"

namespace plt {
namespace {

constexpr int kTF = 32;           // frequencies per tile = lanes
#ifndef PLT_TM_WARPS
#define PLT_TM_WARPS 16
#endif
#ifndef PLT_TM_G
#define PLT_TM_G 2
#endif
constexpr int kTmWarps = PLT_TM_WARPS;  // multiple of 4 (TMEM quarters); 128 registers per thread at 16
constexpr int kTmG = PLT_TM_G;          // Mhat rows in flight per warp and pipeline stage
#ifndef PLT_TM_PF
#define PLT_TM_PF 8
#endif
#ifndef PLT_TM_V3
#define PLT_TM_V3 1
#endif
constexpr int kTmPF = PLT_TM_PF;        // Mhat rows are prefetched into L2 this many list entries ahead (0: off)
constexpr int NC = 8, NN = 27, NOFF = 343, NE = NN * NC, NL = NE - NC, NCH = (NE + 31) / 32;
constexpr int kCenterOff = 171;   // offset index of (0, 0, 0); oi <-> 342 - oi is o <-> -o

__device__ __forceinline__ void cfma(double2& acc, const double2& k, const double2& m) {
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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Four 32x32b.x4 loads: thread i of the warp gets columns [a_q, a_q + 4) of TMEM lane (quarter base + i).
__device__ __forceinline__ void tm_ld4x4(uint32_t (&r)[16], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%17];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%18];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12, %13, %14, %15}, [%19];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3)
      : "memory");
}
// Retires every TMEM load of the thread; the registers go through the statement so that no use is scheduled above it.
__device__ __forceinline__ void tm_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// Is the far offset (o0, o1, o2) kept in TMEM?  (at most one axis at +-3)
__host__ __device__ inline bool tm_eligible(int o0, int o1, int o2) {
  const int a0 = o0 < 0 ? -o0 : o0, a1 = o1 < 0 ? -o1 : o1, a2 = o2 < 0 ? -o2 : o2;
  const bool far = a0 > 1 || a1 > 1 || a2 > 1;
  return far && ((a0 == 3) + (a1 == 3) + (a2 == 3)) <= 1;
}

__global__ void __launch_bounds__(kTmWarps * 32, 1) k_m2l_hadamard_tmem3(M2LArgs a, int F, int n_ftiles) {
  extern __shared__ double2 sm2[];
  double2* Ks = sm2;                                          // [NOFF][TF] operator slice of the current tile
  int2* s_meta = reinterpret_cast<int2*>(Ks + NOFF * kTF);    // [NE]: x = offset index base, y = far mask | T class << 8
  uint4* s_cols = reinterpret_cast<uint4*>(s_meta + NE);      // [NE][2]: TMEM column of the 8 children's operators
  int* s_cmask = reinterpret_cast<int*>(s_cols + 2 * NE);     // [NE]: conj mask of the 8 children
  int2* s_list = reinterpret_cast<int2*>(s_cmask + NE);       // [warps][NL]: T entries from the front, S from the back
  __shared__ unsigned char s_slot_of[kCenterOff];             // canonical offset index -> TMEM slot (0xFF: not in TMEM)
  __shared__ unsigned char s_ci_of[128];                      // TMEM slot -> canonical offset index
  __shared__ int s_n_slots;
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {  // slot numbering of the canonical half (oi < 171)
    int n = 0;
    for (int ci = 0; ci < kCenterOff; ++ci) {
      const int o0 = ci / 49 - 3, o1 = (ci / 7) % 7 - 3, o2 = ci % 7 - 3;
      if (tm_eligible(o0, o1, o2)) {
        s_slot_of[ci] = static_cast<unsigned char>(n);
        s_ci_of[n] = static_cast<unsigned char>(ci);
        ++n;
      } else {
        s_slot_of[ci] = 0xFF;
      }
    }
    s_n_slots = n;  // 124
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int code = threadIdx.x; code < NE; code += blockDim.x) {
    const int nb = code / NC, cs = code % NC;
    int u[3], r = nb;
#pragma unroll
    for (int d = 2; d >= 0; --d) {
      u[d] = 2 * ((r % 3) - 1) + ((cs >> (2 - d)) & 1);  // child-level coordinate of the source, -2 .. 3
      r /= 3;
    }
    const int base = ((u[0] + 3) * 7 + (u[1] + 3)) * 7 + (u[2] + 3);
    int far_mask = 0, conj_mask = 0;
    bool all_tm = true;
    uint32_t col[NC] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (int ct = 0; ct < NC; ++ct) {
      const int o0 = u[0] - ((ct >> 2) & 1), o1 = u[1] - ((ct >> 1) & 1), o2 = u[2] - (ct & 1);
      const bool far = o0 > 1 || o0 < -1 || o1 > 1 || o1 < -1 || o2 > 1 || o2 < -1;
      if (!far) continue;
      far_mask |= 1 << ct;
      const int oi = ((o0 + 3) * 7 + (o1 + 3)) * 7 + (o2 + 3);
      const bool cj = oi > kCenterOff;
      const int ci = cj ? 2 * kCenterOff - oi : oi;
      if (cj) conj_mask |= 1 << ct;
      const int slot = s_slot_of[ci];
      if (slot == 0xFF) all_tm = false;
      else col[ct] = 4u * static_cast<uint32_t>(slot);
    }
    const int cls = (far_mask != 0 && all_tm) ? 1 : 0;
    s_meta[code] = make_int2(base, far_mask | (cls << 8));
    s_cols[2 * code] = make_uint4(col[0], col[1], col[2], col[3]);
    s_cols[2 * code + 1] = make_uint4(col[4], col[5], col[6], col[7]);
    s_cmask[code] = conj_mask;
  }
  const uint32_t tb = s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);  // this warp's TMEM quarter

  const long long n_items = static_cast<long long>(n_ftiles) * a.n_active;
  const long long q_lo = n_items * blockIdx.x / gridDim.x, q_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int2* list = s_list + warp * NL;
  for (long long q0 = q_lo; q0 < q_hi;) {
    const int ftile = static_cast<int>(q0 / a.n_active);
    const int slot_lo = static_cast<int>(q0 - static_cast<long long>(ftile) * a.n_active);
    const long long seg_end = min(q_hi, static_cast<long long>(ftile + 1) * a.n_active);
    const int slot_hi = slot_lo + static_cast<int>(seg_end - q0);
    q0 = seg_end;
    const int f = ftile * kTF + lane;
    const bool fok = f < F;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // previous tile's operators (shared memory and TMEM) no longer in use; tables written
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int e = threadIdx.x; e < NOFF * kTF; e += blockDim.x) {
      const int oi = e / kTF, ff = ftile * kTF + (e % kTF);
      Ks[e] = ff < F ? a.Khat[static_cast<size_t>(oi) * F + ff] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (warp < 4) {  // one warp per TMEM quarter mirrors the canonical TMEM offsets of the slice
      const int ns = s_n_slots;
      for (int sidx = 0; sidx < ns; ++sidx) {
        const double2 v = Ks[static_cast<int>(s_ci_of[sidx]) * kTF + lane];
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb + 4u * sidx),
                     "r"(__double2loint(v.x)), "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)),
                     "r"(__double2hiint(v.y))
                     : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (int slot = slot_lo + warp; slot < slot_hi; slot += kTmWarps) {
      const int* tab = a.src_ids + static_cast<size_t>(slot) * NE;
      const int tmask = a.trg_mask[slot];
      // compact the present source cells that have a far target child, by class
      int nT = 0, nS = 0;
      __syncwarp();  // previous parent done with the list
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int idx = c * 32 + lane;
        const int sid = idx < NE ? __ldcs(tab + idx) : -1;
        int2 meta = make_int2(0, 0);
        if (idx < NE) meta = s_meta[idx];
        const int fm = meta.y & tmask & 0xff;
        const bool pres = sid >= 0 && fm != 0;
        const bool isT = (meta.y >> 8) & 1;
        const unsigned mT = __ballot_sync(0xffffffffu, pres && isT);
        const unsigned mS = __ballot_sync(0xffffffffu, pres && !isT);
        const unsigned lt = (1u << lane) - 1u;
        if (pres && isT) list[nT + __popc(mT & lt)] = make_int2(sid, idx | (fm << 16));
        if (pres && !isT) list[NL - 1 - (nS + __popc(mS & lt))] = make_int2(sid, meta.x | (fm << 16));
        nT += __popc(mT);
        nS += __popc(mS);
      }
      __syncwarp();
      double2 acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = make_double2(0.0, 0.0);

      // ---------------- T entries: operators from Tensor Memory ----------------
      if (nT > 0) {
        uint32_t rA[16], rB[16];
        int code = list[0].y & 0xffff;
        {
          const uint4 c = s_cols[2 * code];
          tm_ld4x4(rA, tb + c.x, tb + c.y, tb + c.z, tb + c.w);
        }
        // One half entry (4 children).  The conj flags of a half are the same for all its far children except where an
        // offset component is 0 (about 7 % of the halves): two straight variants without integer work, one with the
        // sign flip on the high word.
        auto half = [&](uint32_t (&r)[16], const int fm, const int cm, const double2 mh, auto hi) {
          constexpr int h0 = decltype(hi)::value * 4;
          const int fmh = (fm >> h0) & 0xf, cmh = (cm >> h0) & fmh & 0xf;
#if PLT_TM_V3
          // a full half (all four children far, the common case) runs two passes over the children, so that the second
          // multiply-add of a component is 8 DFMAs behind the first; no integer work, no predicates
          if (fmh == 0xf && (cmh == 0 || cmh == 0xf)) {
            double2 k[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              k[q] = make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2]));
              acc[h0 + q].x = fma(k[q].x, mh.x, acc[h0 + q].x);
              acc[h0 + q].y = fma(k[q].x, mh.y, acc[h0 + q].y);
            }
            if (cmh == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[h0 + q].x = fma(-k[q].y, mh.y, acc[h0 + q].x);
                acc[h0 + q].y = fma(k[q].y, mh.x, acc[h0 + q].y);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[h0 + q].x = fma(k[q].y, mh.y, acc[h0 + q].x);
                acc[h0 + q].y = fma(-k[q].y, mh.x, acc[h0 + q].y);
              }
            }
          } else if (cmh == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((fmh >> q) & 1)
                cfma(acc[h0 + q], make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2])), mh);
          } else if (cmh == fmh) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((fmh >> q) & 1)
                cfma_c(acc[h0 + q], make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2])), mh);
          } else {
#else
          if (cmh == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((fmh >> q) & 1)
                cfma(acc[h0 + q], make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2])), mh);
          } else if (cmh == fmh) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((fmh >> q) & 1)
                cfma_c(acc[h0 + q], make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2])), mh);
          } else {
#endif
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if ((fmh >> q) & 1) {
                const uint32_t flip = (static_cast<uint32_t>(cmh) << (31 - q)) & 0x80000000u;
                cfma(acc[h0 + q], make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3] ^ flip, r[4 * q + 2])), mh);
              }
            }
          }
        };
        auto load = [&](double2 (&mh)[kTmG], int base) {
#pragma unroll
          for (int g = 0; g < kTmG; ++g) {
            const int e = base + g;
            mh[g] = (e < nT && fok) ? a.Mhat[static_cast<size_t>(list[e < nT ? e : 0].x) * F + f] : make_double2(0.0, 0.0);
            if (kTmPF > 0 && e + kTmPF < nT && fok) prefetch_l2(&a.Mhat[static_cast<size_t>(list[e + kTmPF].x) * F + f]);
          }
        };
        auto compute = [&](const double2 (&mh)[kTmG], int base) {
#pragma unroll
          for (int g = 0; g < kTmG; ++g) {
            const int e = base + g;
            if (e >= nT) break;  // warp-uniform
            const int fm = list[e].y >> 16, cm = s_cmask[code];
            const uint4 chi = s_cols[2 * code + 1];
#if PLT_TM_V3
            // the next entry's columns are fetched now: their two dependent shared-memory reads are off the path of
            // the TMEM loads issued in the middle of this entry
            const int code_n = list[e + 1 < nT ? e + 1 : e].y & 0xffff;
            const uint4 cn = s_cols[2 * code_n];
#endif
            tm_wait(rA);  // lower half of entry e
            tm_ld4x4(rB, tb + chi.x, tb + chi.y, tb + chi.z, tb + chi.w);
            half(rA, fm, cm, mh[g], std::integral_constant<int, 0>{});
            tm_wait(rB);  // upper half of entry e
#if PLT_TM_V3
            if (e + 1 < nT) tm_ld4x4(rA, tb + cn.x, tb + cn.y, tb + cn.z, tb + cn.w);
            code = code_n;
#else
            if (e + 1 < nT) {
              code = list[e + 1].y & 0xffff;
              const uint4 c = s_cols[2 * code];
              tm_ld4x4(rA, tb + c.x, tb + c.y, tb + c.z, tb + c.w);
            }
#endif
            half(rB, fm, cm, mh[g], std::integral_constant<int, 1>{});
          }
        };
        double2 mhA[kTmG], mhB[kTmG];
        load(mhA, 0);
        for (int g0 = 0; g0 < nT; g0 += 2 * kTmG) {
          load(mhB, g0 + kTmG);
          compute(mhA, g0);
          load(mhA, g0 + 2 * kTmG);
          compute(mhB, g0 + kTmG);
        }
      }

      // ---------------- S entries: operators from the shared-memory slice ----------------
      if (nS > 0) {
        auto load = [&](double2 (&mh)[kTmG], int (&pk)[kTmG], int base) {
#pragma unroll
          for (int g = 0; g < kTmG; ++g) {
            const int e = base + g;
            int2 le = make_int2(0, 0);
            if (e < nS) le = list[NL - 1 - e];
            pk[g] = le.y;
            mh[g] = (e < nS && fok) ? a.Mhat[static_cast<size_t>(le.x) * F + f] : make_double2(0.0, 0.0);
            if (kTmPF > 0 && e + kTmPF < nS && fok) prefetch_l2(&a.Mhat[static_cast<size_t>(list[NL - 1 - e - kTmPF].x) * F + f]);
          }
        };
        auto compute = [&](const double2 (&mh)[kTmG], const int (&pk)[kTmG], int base) {
#pragma unroll
          for (int g = 0; g < kTmG; ++g) {
            if (base + g >= nS) break;  // warp-uniform
            const int fm = pk[g] >> 16;
            const double2* kp = Ks + (pk[g] & 0xffff) * kTF + lane;
#pragma unroll
            for (int ct = 0; ct < NC; ++ct) {
              const int cto = ((ct >> 2) & 1) * 49 + ((ct >> 1) & 1) * 7 + (ct & 1);
              if ((fm >> ct) & 1) cfma(acc[ct], kp[-cto * kTF], mh[g]);
            }
          }
        };
        double2 mhA[kTmG], mhB[kTmG];
        int pkA[kTmG], pkB[kTmG];
        load(mhA, pkA, 0);
        for (int g0 = 0; g0 < nS; g0 += 2 * kTmG) {
          load(mhB, pkB, g0 + kTmG);
          compute(mhA, pkA, g0);
          load(mhA, pkA, g0 + 2 * kTmG);
          compute(mhB, pkB, g0 + kTmG);
        }
      }
      if (fok) {
#pragma unroll
        for (int ct = 0; ct < NC; ++ct)
          if ((tmask >> ct) & 1) __stcs(&a.Lhat[(static_cast<size_t>(slot) * NC + ct) * F + f], acc[ct]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem) : "memory");
}

}  // namespace

// Opt-in (PLT_HAD_TMEM=1 in the environment, or plt_set_hadamard_tmem): measured on config #3 the kernel is parity-green
// but SLOWER than the shared-memory list kernel (15.2 ms against 11.4 ms, profiles/r02_k_hadamard_tmem.md): with 16 warps
// of 126 registers it issues ~18 instructions per complex multiply-add (TMEM addresses need an IADD + R2UR per load, the
// variant selection a few branches per half entry) and sits at 62 % issue utilisation with fixed-latency stalls on top.
namespace {
int g_tmem_mode = -1;  // -1: take the environment
}
bool hadamard_tmem_enabled() {
  if (g_tmem_mode < 0) {
    const char* e = getenv("PLT_HAD_TMEM");
    g_tmem_mode = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_tmem_mode == 1;
}
int hadamard_tmem_set(int on) {
  const int before = hadamard_tmem_enabled() ? 1 : 0;
  g_tmem_mode = on ? 1 : 0;
  return before;
}

// Scalar 3-D lists only; returns false when the caller has to take the shared-memory kernel.
bool launch_m2l_hadamard_tmem(const M2LArgs& a, int F, cudaStream_t s, LaunchCounter& c) {
  if (!hadamard_tmem_enabled() || a.dim != 3 || a.kn * a.km != 1) return false;
  const int n_ftiles = ceil_div(F, kTF);
  const long long rounds = static_cast<long long>(n_ftiles) * ceil_div(a.n_active, kTmWarps);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(num_sm(), rounds)));
  const size_t smem = sizeof(double2) * NOFF * kTF + (sizeof(int2) + 2 * sizeof(uint4) + sizeof(int)) * NE +
                      sizeof(int2) * NL * kTmWarps;
  static bool opted = false;
  if (!opted) {
    PLT_CUDA(cudaFuncSetAttribute((const void*)k_m2l_hadamard_tmem3, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    opted = true;
  }
  PLT_LAUNCH(c, k_m2l_hadamard_tmem3, grid, kTmWarps * 32, smem, s, a, F, n_ftiles);
  return true;
}

}  // namespace plt
