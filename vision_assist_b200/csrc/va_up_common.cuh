// Device helpers shared by the stand-alone upsample kernels (va_upsample.cu) and the fused
// TMA + tcgen05 kernel (va_fused_tc.cu): 0/1 byte packing, per-thread mask reductions, lattice
// sampling and the exact-4x bilinear arithmetic (see va_upsample.cu for the numerics contract).
#pragma once

#include <climits>
#include <type_traits>

#include "va_common.cuh"
#include "va_contour_core.h"

namespace va {

// ---------------------------------------------------------------------------------------------
// shared helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned pack4(float a, float b, float c, float d) {
  return (a > 0.f ? 1u : 0u) | (b > 0.f ? 0x100u : 0u) | (c > 0.f ? 0x10000u : 0u) | (d > 0.f ? 0x1000000u : 0u);
}

struct ThreadStats {
  unsigned acc = 0;                 // four byte-lane counters (<= 255 rows*words per lane between flushes)
  unsigned area = 0;
  unsigned orw[4] = {0, 0, 0, 0};   // OR of all rows' mask words (x extent)
  int miny = INT_MAX, maxy = -1;
  // rows must be added in increasing Y; at most 63 rows between flush() calls
  __device__ __forceinline__ void add_row(const uint4& w, int Y) {
    const unsigned any = w.x | w.y | w.z | w.w;
    if (any) {
      acc += (w.x + w.y) + (w.z + w.w);
      orw[0] |= w.x; orw[1] |= w.y; orw[2] |= w.z; orw[3] |= w.w;
      if (miny == INT_MAX) miny = Y;
      maxy = Y;
    }
  }
  __device__ __forceinline__ void flush() {
    area += (acc & 0xffu) + ((acc >> 8) & 0xffu) + ((acc >> 16) & 0xffu) + (acc >> 24);
    acc = 0;
  }
};

// Reduce over the warp (all lanes belong to the same (frame, instance)) and publish with atomics.
__device__ __forceinline__ void publish_stats(const ThreadStats& t, int xbase, InstStats* dst) {
  int minx = INT_MAX, maxx = -1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (t.orw[q]) {
      const int lo = (__ffs(t.orw[q]) - 1) >> 3, hi = (31 - __clz(t.orw[q])) >> 3;
      minx = min(minx, xbase + 4 * q + lo);
      maxx = max(maxx, xbase + 4 * q + hi);
    }
  }
  unsigned area = t.area + (t.acc & 0xffu) + ((t.acc >> 8) & 0xffu) + ((t.acc >> 16) & 0xffu) + (t.acc >> 24);
  int miny = t.miny, maxy = t.maxy;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    area += __shfl_xor_sync(0xffffffffu, area, o);
    minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, o));
    miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, o));
    maxx = max(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
    maxy = max(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
  }
  if ((threadIdx.x & 31) == 0 && area) {
    atomicAdd(&dst->area, area);
    atomicMin(&dst->minx, minx);
    atomicMin(&dst->miny, miny);
    atomicMax(&dst->maxx, maxx);
    atomicMax(&dst->maxy, maxy);
  }
}

// Sample the cell-centre lattice points that fall into this thread's 16 pixels of row Y (rare path:
// one dst row in gs is a lattice row).  Scalars by value so that the out-of-line copy needs no stack.
static __device__ __noinline__ void lattice_row_impl(unsigned w0, unsigned w1, unsigned w2, unsigned w3, int Y, int xbase, int gs,
                                                     int lat_cols, int lat_words, unsigned* lat) {
  const int half = gs >> 1;
  const int ly = (Y - half) / gs;
  int lx = (xbase - half + gs - 1) / gs;
  if (lx < 0) lx = 0;
  for (; lx < lat_cols; ++lx) {
    const int pos = gs * lx + half - xbase;
    if (pos > 15) break;
    const unsigned ww = (pos < 4) ? w0 : (pos < 8) ? w1 : (pos < 12) ? w2 : w3;
    if ((ww >> (8 * (pos & 3))) & 1u) atomicOr(&lat[ly * lat_words + (lx >> 5)], 1u << (lx & 31));
  }
}
__device__ __forceinline__ void lattice_row(const uint4& w, int Y, int xbase, const Dims& d, unsigned* lat) {
  lattice_row_impl(w.x, w.y, w.z, w.w, Y, xbase, d.gs, d.lat_cols, d.lat_words, lat);
}

// ---------------------------------------------------------------------------------------------
// per-row by-products for the contour step (va_contour_core.h): every mask kernel maps the 8 lanes of a lane group
// (lane & 7 = 16-pixel piece, lane >> 3 = row slot) to one 128-pixel block of one dst row
// ---------------------------------------------------------------------------------------------
// 16 mask bytes (0 / 1) -> 16 bits
__device__ __forceinline__ unsigned pat16(const uint4& w) {
  return ((w.x * 0x01020408u) >> 24) | (((w.y * 0x01020408u) >> 24) << 4) | (((w.z * 0x01020408u) >> 24) << 8) |
         (((w.w * 0x01020408u) >> 24) << 12);
}
// 16 bits -> 16 mask bytes
__device__ __forceinline__ uint4 bytes16(unsigned pat) {
  auto ex = [](unsigned n) { return ((n & 0xfu) * 0x00204081u) & 0x01010101u; };
  return make_uint4(ex(pat), ex(pat >> 4), ex(pat >> 8), ex(pat >> 12));
}

// 16-pixel patterns of the dst rows of one lane (slot s = row ordinal inside the task): up to 6 rows in three
// registers (exact 4x), up to 12 rows in six (generic scale: the first prototype row pair also owns the dst rows whose
// source row clamps to 0, about 1.5 x the vertical scale)
template <bool kWide>
struct RowPats;
template <>
struct RowPats<false> {
  static constexpr int kSlots = 6;
  unsigned long long lo = 0ull;
  unsigned hi = 0u;
  __device__ __forceinline__ void set(int s, unsigned pat) {
    if (s < 4) lo |= (unsigned long long)pat << (16 * s);
    else hi |= pat << (16 * (s - 4));
  }
  __device__ __forceinline__ unsigned get(int s) const {
    return (s < 4) ? (unsigned)(lo >> (16 * s)) & 0xffffu : (hi >> (16 * (s - 4))) & 0xffffu;
  }
  __device__ __forceinline__ void fill(int n) {      // the first n rows all ones (n = 2, 4 or 6)
    lo = (n >= 4) ? ~0ull : 0xffffffffull;
    hi = (n == 6) ? ~0u : 0u;
  }
  __device__ __forceinline__ bool is_zero() const { return (lo | hi) == 0; }
  __device__ __forceinline__ bool is_full(int n) const {
    RowPats f;
    f.fill(n);
    return n > 0 && lo == f.lo && hi == f.hi;
  }
};
template <>
struct RowPats<true> {
  static constexpr int kSlots = 12;
  unsigned long long q0 = 0ull, q1 = 0ull, q2 = 0ull;
  static __device__ __forceinline__ unsigned long long ones(int n) {      // n rows of 16 ones, 0 <= n <= 4
    return (n >= 4) ? ~0ull : ((1ull << (16 * n)) - 1ull);
  }
  __device__ __forceinline__ void set(int s, unsigned pat) {
    const unsigned long long v = (unsigned long long)pat << (16 * (s & 3));
    if (s < 4) q0 |= v; else if (s < 8) q1 |= v; else q2 |= v;
  }
  __device__ __forceinline__ unsigned get(int s) const {
    const unsigned long long q = (s < 4) ? q0 : (s < 8) ? q1 : q2;
    return (unsigned)(q >> (16 * (s & 3))) & 0xffffu;
  }
  __device__ __forceinline__ void fill(int n) {
    q0 = ones(min(n, 4));
    q1 = ones(max(min(n - 4, 4), 0));
    q2 = ones(max(min(n - 8, 4), 0));
  }
  __device__ __forceinline__ bool is_zero() const { return (q0 | q1 | q2) == 0; }
  __device__ __forceinline__ bool is_full(int n) const {
    RowPats f;
    f.fill(n);
    return n > 0 && q0 == f.q0 && q1 == f.q1 && q2 == f.q2;
  }
};

// what the leader lane (lane & 7 == 0) of a group accumulates for the instance's InstStats
struct LeaderStats {
  unsigned area = 0;
  int minx = INT_MAX, miny = INT_MAX, maxx = -1, maxy = -1;
};

// One dst row of a lane group: gather the 8 patterns into the block's 128-bit pattern (4 full-warp shuffles - a
// warp reduction with a partial member mask would be serialised group by group), derive the (row, block) summary at
// the leader, store it, feed the leader's stats.  ALL 32 lanes of the warp must call, converged; `exists` = this
// lane's row is a real row.  Returns the block pattern (valid in every lane of the group).
__device__ __forceinline__ uint4 emit_row_summary(unsigned pat, int gl, bool exists, int Y, int blk,
                                                  uint32_t* __restrict__ rowsum_inst, int nblk, LeaderStats& ls) {
  unsigned u = __shfl_xor_sync(0xffffffffu, pat, 1);
  const unsigned w32 = (gl & 1) ? (u | (pat << 16)) : (pat | (u << 16));
  u = __shfl_xor_sync(0xffffffffu, w32, 2);
  const unsigned lo = (gl & 2) ? u : w32, hi = (gl & 2) ? w32 : u;
  const unsigned ulo = __shfl_xor_sync(0xffffffffu, lo, 4), uhi = __shfl_xor_sync(0xffffffffu, hi, 4);
  uint4 w;
  w.x = (gl & 4) ? ulo : lo; w.y = (gl & 4) ? uhi : hi;
  w.z = (gl & 4) ? lo : ulo; w.w = (gl & 4) ? hi : uhi;
  if (exists && gl == 0) {
    const int c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    unsigned e = 0u;
    if (c) {
      const int f = w.x ? __ffs((int)w.x) - 1 : w.y ? 31 + __ffs((int)w.y) : w.z ? 63 + __ffs((int)w.z) : 95 + __ffs((int)w.w);
      const int l = w.w ? 127 - __clz((int)w.w) : w.z ? 95 - __clz((int)w.z) : w.y ? 63 - __clz((int)w.y) : 31 - __clz((int)w.x);
      e = cc::rowsum_pack(c, f, l);
      ls.area += (unsigned)c;
      ls.minx = min(ls.minx, cc::kRowBlock * blk + f);
      ls.maxx = max(ls.maxx, cc::kRowBlock * blk + l);
      ls.miny = min(ls.miny, Y);
      ls.maxy = max(ls.maxy, Y);
    }
    if (e) rowsum_inst[(size_t)Y * nblk + blk] = e;      // zero entries are the array's resting state
  }
  return w;
}

__device__ __forceinline__ void publish_leader(const LeaderStats& ls, InstStats* dst) {
  if (ls.area) {
    atomicAdd(&dst->area, ls.area);
    atomicMin(&dst->minx, ls.minx);
    atomicMin(&dst->miny, ls.miny);
    atomicMax(&dst->maxx, ls.maxx);
    atomicMax(&dst->maxy, ls.maxy);
  }
}

// lattice samples from a 16-bit pattern (see lattice_row)
static __device__ __noinline__ void lattice_row_pat(unsigned pat, int Y, int xbase, int gs, int lat_cols, int lat_words,
                                                    unsigned* lat) {
  const int half = gs >> 1;
  const int ly = (Y - half) / gs;
  int lx = (xbase - half + gs - 1) / gs;
  if (lx < 0) lx = 0;
  for (; lx < lat_cols; ++lx) {
    const int pos = gs * lx + half - xbase;
    if (pos > 15) break;
    if ((pat >> pos) & 1u) atomicOr(&lat[ly * lat_words + (lx >> 5)], 1u << (lx & 31));
  }
}

// ---------------------------------------------------------------------------------------------
// exact 4x kernel
// ---------------------------------------------------------------------------------------------
constexpr int kStrip = 8;           // proto row pairs per thread
constexpr int kUpThreads = 128;

// horizontal pass for one proto row: 6 source values (cols 4g-1 .. 4g+4, clamped) -> 16 outputs
__device__ __forceinline__ void hinterp4(const float (&s)[6], float (&h)[16], bool left_edge) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = s[k], b = s[k + 1], c = s[k + 2];
    h[4 * k + 0] = fmaf(a, 0.375f, __fmul_rn(b, 0.625f));
    h[4 * k + 1] = fmaf(a, 0.125f, __fmul_rn(b, 0.875f));
    h[4 * k + 2] = fmaf(b, 0.875f, __fmul_rn(c, 0.125f));
    h[4 * k + 3] = fmaf(b, 0.625f, __fmul_rn(c, 0.375f));
  }
  if (left_edge) {   // dst x = 0,1: src clamps to 0 -> l0 = 1, l1 = 0 -> a + b*0 = a
    h[0] = fmaf(s[1], 1.0f, __fmul_rn(s[2], 0.0f));
    h[1] = h[0];
  }
}

__device__ __forceinline__ void load6(const float* __restrict__ row, int g, int mw, float (&s)[6], float& mn, float& mx) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(row + 4 * g));
  s[1] = v.x; s[2] = v.y; s[3] = v.z; s[4] = v.w;
  s[0] = (g > 0) ? __ldg(row + 4 * g - 1) : v.x;
  s[5] = (4 * g + 4 < mw) ? __ldg(row + 4 * g + 4) : v.w;
  mn = fminf(fminf(fminf(s[0], s[1]), fminf(s[2], s[3])), fminf(s[4], s[5]));
  mx = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), fmaxf(s[4], s[5]));
}

// ATen's align_corners=False source index and lambdas (area_pixel_compute_source_index + guard_index_and_lambda)
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
  src = fmaxf(src, 0.f);
  i0 = min((int)src, in_size - 1);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  l0 = 1.0f - l1;
}

constexpr float kTiny = 1e-30f;   // below this a positive product could underflow: take the exact path

__device__ __forceinline__ uint4 vblend(const float (&hA)[16], const float (&hB)[16], float l0, float l1) {
  float o[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) o[x] = fmaf(hA[x], l0, __fmul_rn(hB[x], l1));
  return make_uint4(pack4(o[0], o[1], o[2], o[3]), pack4(o[4], o[5], o[6], o[7]), pack4(o[8], o[9], o[10], o[11]),
                    pack4(o[12], o[13], o[14], o[15]));
}

__device__ __forceinline__ uint4 hpack(const float (&h)[16]) {
  return make_uint4(pack4(h[0], h[1], h[2], h[3]), pack4(h[4], h[5], h[6], h[7]), pack4(h[8], h[9], h[10], h[11]),
                    pack4(h[12], h[13], h[14], h[15]));
}

}  // namespace va
