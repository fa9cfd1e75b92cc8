// Contour semantics of the reference without contour tracing - the algorithm shared by the CUDA kernels
// (va_tail.cu: the tail kernel runs both paths per frame) and a host build that the CPU tests drive (tests/native/contour_host.cpp).
//
// Reference behaviour reproduced (bit-exact, see oracle/contour.py for the model and its OpenCV pin):
//   masks2segments          vendored ultralytics ops.py:837-859   findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE),
//                                                                 contour with the most points (first maximum)
//   polygon selection       FrameProcessor.py:72-73               largest cv2.contourArea (first maximum)
//   boundingRect, fillPoly  FrameProcessor.py:75-86               bbox and raster of the kept contour
//
// Per instance mask:  G = complement of the 4-connected background region that touches the frame.  Its 8-connected
// components are the top-level components with holes filled; one RETR_EXTERNAL contour each.  Points and doubled
// shoelace area of a component are sums of a 3x3 table (va_contour_lut.h) over its pixels; the kept component is the
// one with the most points (ties: last in raster order = OpenCV's first); its pixels ARE the fillPoly raster.
//
// Two paths:
//   * certificate (cert_row): every mask row of the instance is ONE run and consecutive rows touch -> one
//     hole-free component; its doubled area follows in closed form from the per-row run ends (2N - L - 2 with L the
//     number of border moves).  Input: the per-row summaries the mask kernels emit - no pixel is re-read.
//   * general (Work / phase_*): run-based connected components on a bit image of the mask's bounding box:
//     background gaps united 4-connectedly (union-find, node 0 = the outside) -> holes; foreground runs united
//     8-connectedly and across holes -> components of G; table sums per component; selection; bbox; lattice samples.
//     Written as barrier-separated phases `phase(w, tid, nthreads)` so the same code runs as one CTA per instance on
//     the GPU and as a plain loop on the host.
#pragma once

#include <stdint.h>

#include "va_contour_lut.h"

#if defined(__CUDACC__)
#define VA_HD __host__ __device__ __forceinline__
// The phases run once per mask, mostly a row or a run per thread: unrolled loops only add to the instruction
// footprint of the kernel that inlines all of them (instruction fetch, not issue, bounds the single-thread sections).
#define VA_ROLL _Pragma("unroll 1")
#else
#define VA_HD inline
#define VA_ROLL
#endif

namespace va {
namespace cc {

// ---------------------------------------------------------------------------------------------
// portable intrinsics / atomics (the host build is single-threaded)
// ---------------------------------------------------------------------------------------------
VA_HD int popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
VA_HD int ffs32(uint32_t v) {   // 1-based index of the lowest set bit, 0 if none
#ifdef __CUDA_ARCH__
  return __ffs((int)v);
#else
  return v ? __builtin_ctz(v) + 1 : 0;
#endif
}
VA_HD int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
// Atomics on scratch that lives in shared memory when it fits and in a global slab otherwise.  A generic atomic
// (ATOM.E) that lands in the shared window is far slower than the shared-space instruction, so the device versions
// test the address space and issue atom / red .shared explicitly.
#ifdef __CUDA_ARCH__
#define VA_SH(p) ((unsigned)__cvta_generic_to_shared(p))
#endif
VA_HD int atom_min(int* p, int v) {        // returns the old value
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { int o; asm volatile("atom.shared.min.s32 %0, [%1], %2;" : "=r"(o) : "r"(VA_SH(p)), "r"(v) : "memory"); return o; }
  return atomicMin(p, v);
#else
  const int o = *p; if (v < o) *p = v; return o;
#endif
}
VA_HD void atom_max(int* p, int v) {
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(VA_SH(p)), "r"(v) : "memory"); return; }
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
VA_HD void atom_min_nr(int* p, int v) {
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(VA_SH(p)), "r"(v) : "memory"); return; }
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
VA_HD void atom_add(int* p, int v) {
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(VA_SH(p)), "r"(v) : "memory"); return; }
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
VA_HD void atom_or(unsigned* p, unsigned v) {
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(VA_SH(p)), "r"(v) : "memory"); return; }
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
VA_HD void atom_max64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
  if (__isShared(p)) { asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(VA_SH(p)), "l"(v) : "memory"); return; }
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
VA_HD int imin(int a, int b) { return a < b ? a : b; }
VA_HD int imax(int a, int b) { return a > b ? a : b; }
VA_HD int iabs(int a) { return a < 0 ? -a : a; }

// ---------------------------------------------------------------------------------------------
// per-row summaries written by the mask kernels: one entry per (row, 128-pixel block)
//   bits 0-7 pixels set in the block (0 = none), bits 8-14 first set pixel, bits 15-21 last set pixel (block-relative)
// Protocol: the array is all zero between calls; a mask kernel only stores the non-zero entries, and whoever consumes
// an instance's rows (certificate kernel, general path) stores zeros back.
// ---------------------------------------------------------------------------------------------
constexpr int kRowBlock = 128;
VA_HD uint32_t rowsum_pack(int cnt, int first, int last) { return (uint32_t)cnt | ((uint32_t)first << 8) | ((uint32_t)last << 15); }

struct RowRun { int cnt, a, b; };   // pixels set in the row, first and last set pixel (a > b when empty)

// fold block k's summary into the row's (blocks in ascending order)
VA_HD void rowsum_accumulate(RowRun& r, uint32_t v, int k) {
  const int c = (int)(v & 0xffu);
  if (c) {
    if (r.cnt == 0) r.a = k * kRowBlock + (int)((v >> 8) & 0x7fu);
    r.b = k * kRowBlock + (int)((v >> 15) & 0x7fu);
    r.cnt += c;
  }
}
VA_HD RowRun rowsum_combine(const uint32_t* e, int nblk) {
  RowRun r; r.cnt = 0; r.a = 1 << 30; r.b = -1;
  VA_ROLL
  for (int k0 = 0; k0 < nblk; k0 += 8) {
    uint32_t v[8];                                     // all loads of a batch are issued before the first use
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < nblk) ? e[k0 + j] : 0u;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) rowsum_accumulate(r, v[j], k0 + j);
  }
  return r;
}

// Result of the contour step for one instance (consumed by the tail kernel's selection).
struct InstContour {
  int area2;        // doubled cv2.contourArea of the kept polygon
  int state;        // kEmpty / kSimple / kGeneral / kPending / kOverflow
  int minx, miny, maxx, maxy;   // cv2.boundingRect of the kept polygon (pixel bbox of the kept component)
  int points;       // CHAIN_APPROX_SIMPLE points of the kept contour (general path; 0 on the certificate path)
  int n_components; // top-level components (general path; 1 on the certificate path)
};
enum { kEmpty = 0, kSimple = 1, kGeneral = 2, kPending = 3, kOverflow = 4 };

// Certificate terms of one row y (miny <= y <= maxy) given its run and the previous row's run.
struct CertTerms { int ok, n, l, minx, maxx; };
VA_HD CertTerms cert_row(const RowRun& cur, const RowRun& prev, bool first_row, bool last_row) {
  CertTerms t; t.ok = 1; t.n = cur.cnt; t.l = 0; t.minx = cur.a; t.maxx = cur.b;
  if (cur.cnt == 0 || cur.cnt != cur.b - cur.a + 1) { t.ok = 0; return t; }       // empty row inside the range / several runs
  if (!first_row) {
    if (prev.cnt == 0 || prev.cnt != prev.b - prev.a + 1) { t.ok = 0; return t; }
    if (!(cur.a <= prev.b + 1 && cur.b >= prev.a - 1)) { t.ok = 0; return t; }   // rows do not touch (8-connectivity)
    t.l += imax(1, iabs(cur.a - prev.a)) + imax(1, iabs(cur.b - prev.b));       // border moves down the left and the right side
  }
  if (first_row) t.l += cur.b - cur.a;                                          // moves along the top run
  if (last_row) t.l += cur.b - cur.a;                                           // ... and back along the bottom run
  return t;
}

// ---------------------------------------------------------------------------------------------
// general path
// ---------------------------------------------------------------------------------------------
// Rows whose summary says "one run" (or "empty") never touch the pixels: their bit words are synthesised from the run
// ends, their run table entry is written directly, and run look-ups on them are two comparisons.  Only rows with
// several runs are read back (u8 masks or bit rows) and scanned word by word.
struct Work {
  // ---- input ----
  int H, W;                 // frame
  int fmt;                  // 0: u8 mask (non-zero = set), 1: bit rows (bit x&31 of word x>>5)
  const uint8_t* px;        // fmt 0: [H][W] of this instance
  const uint32_t* bits;     // fmt 1: [H][bit_words] of this instance
  int bit_words;
  uint32_t* rowsum;         // [H][nblk] per-(row, 128 px block) summaries of this instance (read, then reset to 0)
  int nblk;
  int y0, x0w;              // region origin: first row, first 32-pixel word
  int R, Wd;                // region rows / words (covers the pixel bbox of the mask)
  int gs, lat_rows, lat_cols, lat_words;
  // ---- scratch (shared or global memory) ----
  uint32_t* Mfg;            // [NM][Wd] foreground bits of the rows with several runs (slot = position in mlist)
  uint32_t* G;              // [NM][Wd] foreground + holes
  uint16_t* S;              // [NM][Wd] runs of the row that start before word k
  int16_t* one_a;           // [R] first pixel of the row's only run (region-relative); kRowEmpty / kRowMulti otherwise
  int16_t* one_b;           // [R] last pixel of the row's only run; rows with several runs: their slot in Mfg / G / S
  int16_t* mlist;           // [R] rows with several runs (sc[W_NM] entries, any order)
  int16_t* nplist;          // [R] from the front: non-empty rows that are not "plain" (sc[W_NNP] entries), summed word by word;
                            //     from the back: plain rows with long candidate ranges (sc[W_NLONG] entries)
  int* rowoff;              // [R + 1] first run id of the row
  int cap;                  // run capacity
  uint16_t* rs;             // [cap] first / last pixel (region-relative) and row of a run; ids are in raster order
  uint16_t* re;
  uint16_t* ry;
  int* pF;                  // [cap] union-find over runs: components of G (root = smallest id = raster-first run)
  int* pG;                  // [cap + 1] union-find over background gaps: node 0 = outside, node id+1 = gap right of run id
  int* accP;                // [cap] per root: points
  int* accA;                // [cap] per root: signed doubled area
  int* seg;                 // [33] scan scratch
  // ---- scalars (one copy per CTA, shared memory) ----
  int* sc;                  // see enum
  unsigned long long* best; // [1] (points << 32) | root id
  // ---- output ----
  unsigned* lattice;        // [lat_rows][lat_words] of this instance (overwritten)
  InstContour* out;
};
enum { W_NR, W_OVERFLOW, W_HOLES, W_ROOTS, W_MINX, W_MINY, W_MAXX, W_MAXY, W_CHOSEN, W_NM, W_NNP, W_NLONG, W_LIGHT, W_COUNT };
constexpr int kRowEmpty = -1, kRowMulti = -2;

// bits [a, b] of word k (pixels 32k .. 32k+31), a <= b
VA_HD uint32_t span_bits(int a, int b, int k) {
  const int lo = imax(a, 32 * k), hi = imin(b, 32 * k + 31);
  if (lo > hi) return 0u;
  return (0xffffffffu << (lo & 31)) & (0xffffffffu >> (31 - (hi & 31)));
}
// first word of a row with several runs in Mfg / G / S
VA_HD int slot_base(const Work& w, int r) { return (int)w.one_b[r] * w.Wd; }
// word k of row r of a bit image: stored for rows with several runs, made up from the run ends otherwise
VA_HD uint32_t word_at(const uint32_t* bm, const Work& w, int r, int k) {
  if (r < 0 || r >= w.R || k < 0 || k >= w.Wd) return 0u;
  const int a = w.one_a[r];
  if (a == kRowMulti) return bm[slot_base(w, r) + k];
  return (a >= 0) ? span_bits(a, (int)w.one_b[r], k) : 0u;
}
VA_HD uint32_t rise_at(const Work& w, int r, int k) {      // bits where a foreground run starts
  const uint32_t m = word_at(w.Mfg, w, r, k), p = word_at(w.Mfg, w, r, k - 1);
  return m & ~((m << 1) | (p >> 31));
}
VA_HD uint32_t fall_at(const Work& w, int r, int k) {      // bits where a foreground run ends
  const uint32_t m = word_at(w.Mfg, w, r, k), n = word_at(w.Mfg, w, r, k + 1);
  return m & ~((m >> 1) | (n << 31));
}
VA_HD bool fg_at(const Work& w, int r, int x) {
  if (x < 0 || x >= 32 * w.Wd) return false;
  const int a = w.one_a[r];
  if (a != kRowMulti) return a >= 0 && x >= a && x <= (int)w.one_b[r];
  return (w.Mfg[slot_base(w, r) + (x >> 5)] >> (x & 31)) & 1u;
}
VA_HD int row_runs(const Work& w, int r) { return w.rowoff[r + 1] - w.rowoff[r]; }
// number of runs of row r that start at a pixel <= x
VA_HD int ns(const Work& w, int r, int x) {
  if (x < 0) return 0;
  const int a = w.one_a[r];
  if (a != kRowMulti) return (a >= 0 && x >= a) ? 1 : 0;
  if (x >= 32 * w.Wd) return row_runs(w, r);
  const int k = x >> 5;
  return (int)w.S[slot_base(w, r) + k] + popc32(rise_at(w, r, k) & (0xffffffffu >> (31 - (x & 31))));
}

// find with path halving: every visited node is re-pointed at its grandparent.  The concurrent writes are benign -
// a parent is only ever replaced by one of its ancestors, so chains only get shorter (a vertical stack of runs would
// otherwise leave a chain as long as the mask is tall).
VA_HD int uf_find(int* p, int x) {
  VA_ROLL
  while (true) {
    const int q = p[x];
    if (q == x) return x;
    const int g = p[q];
    if (g != q) p[x] = g;
    x = g;
  }
}
// roots are the smallest ids: lock-free union by atomicMin on the larger root
VA_HD void uf_union(int* p, int a, int b) {
  VA_ROLL
  while (true) {
    a = uf_find(p, a);
    b = uf_find(p, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atom_min(&p[a], b);
    if (old == a) return;
    a = old;
  }
}

// (row, word) tasks t = tid, tid + nt, ... without a division per task
struct WordIter {
  int t, r, k, dr, dk, nt, Wd, end;
  VA_HD WordIter(const Work& w, int tid, int nthreads) {
    Wd = w.Wd; nt = nthreads; end = w.R * w.Wd;
    t = tid; r = tid / Wd; k = tid - r * Wd; dr = nt / Wd; dk = nt - dr * Wd;
  }
  VA_HD bool valid() const { return t < end; }
  VA_HD void next() { t += nt; r += dr; k += dk; if (k >= Wd) { k -= Wd; ++r; } }
};

VA_HD bool row_is_multi(const Work& w, int r) { return r >= 0 && r < w.R && w.one_a[r] == kRowMulti; }
// a row with one run whose two neighbour rows have at most one run each: no hole touches it, its border pixels and
// their 3x3 codes follow from the three pairs of run ends
VA_HD bool row_is_plain(const Work& w, int r) { return w.one_a[r] >= 0 && !row_is_multi(w, r - 1) && !row_is_multi(w, r + 1); }

// ---- phase 0: scalars + row classes from the summaries ----
VA_HD void phase_init(Work& w, int tid, int nt) {
  if (tid == 0) {
    VA_ROLL
    for (int q = 0; q < W_COUNT; ++q) w.sc[q] = 0;
    w.sc[W_MINX] = 1 << 30; w.sc[W_MINY] = 1 << 30; w.sc[W_MAXX] = -1; w.sc[W_MAXY] = -1;
    w.sc[W_CHOSEN] = -1;
    w.sc[W_LIGHT] = 1;                                  // cleared by phase_light_check
    w.best[0] = 0ull;
  }
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {
    uint32_t* e = w.rowsum + (size_t)(w.y0 + r) * w.nblk;
    const RowRun rr = rowsum_combine(e, w.nblk);
    VA_ROLL
    for (int k = 0; k < w.nblk; ++k) e[k] = 0u;       // the summaries are all zero between calls: the mask kernels only write non-zero ones
    int a = kRowMulti, b = 0;
    if (rr.cnt == 0) { a = kRowEmpty; }
    else if (rr.cnt == rr.b - rr.a + 1) { a = rr.a - 32 * w.x0w; b = rr.b - 32 * w.x0w; }
    w.one_a[r] = (int16_t)a;
    w.one_b[r] = (int16_t)b;
  }
}

struct Span { int a, b; };     // a > b: empty
VA_HD Span row_span(const Work& w, int r) {
  Span s; s.a = 1; s.b = 0;
  if (r >= 0 && r < w.R && w.one_a[r] >= 0) { s.a = w.one_a[r]; s.b = w.one_b[r]; }
  return s;
}
// the two pixel ranges of a plain row that can hold border pixels: [c.a, la] and [lb, c.b] (second one possibly empty)
struct PlainRanges { Span c, u, d; int la, lb, n1, n2; };
VA_HD PlainRanges plain_ranges(const Work& w, int r) {
  PlainRanges g;
  g.c = row_span(w, r); g.u = row_span(w, r - 1); g.d = row_span(w, r + 1);
  const bool both = (g.u.a <= g.u.b) && (g.d.a <= g.d.b);
  g.la = both ? imin(g.c.b, imax(g.c.a, imax(g.u.a, g.d.a)) + 1) : g.c.b;
  g.lb = both ? imax(g.la + 1, imax(g.c.a, imin(g.c.b, imin(g.u.b, g.d.b)) - 1)) : g.c.b + 1;
  g.n1 = g.la - g.c.a + 1; g.n2 = g.c.b - g.lb + 1;
  return g;
}
constexpr int kPlainInline = 8;   // plain rows with more candidate pixels than this go to the long list
// ---- phase 0b: the (few) rows that need word-level work ----
VA_HD int atom_inc(int* p) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, 1);
#else
  return (*p)++;
#endif
}
VA_HD void phase_lists(Work& w, int tid, int nt) {
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {
    const int a = w.one_a[r];
    if (a == kRowMulti) {
      const int slot = atom_inc(&w.sc[W_NM]);
      w.mlist[slot] = (int16_t)r;
      w.one_b[r] = (int16_t)slot;
    }
    if (a == kRowEmpty) continue;
    if (!row_is_plain(w, r)) {
      w.nplist[atom_inc(&w.sc[W_NNP])] = (int16_t)r;
    } else {
      // plain rows with long candidate ranges (flat edges) are summed by a whole warp: listed from the back
      const PlainRanges g = plain_ranges(w, r);
      if (g.n1 + g.n2 > kPlainInline) w.nplist[w.R - 1 - atom_inc(&w.sc[W_NLONG])] = (int16_t)r;
    }
  }
}

// ---- phase 1: bit rows of the rows with several runs (every other row is made up from its run ends on demand) ----
VA_HD void phase_load(Work& w, int tid, int nt) {
  const int ntask = w.sc[W_NM] * w.Wd;
  VA_ROLL
  for (int q = tid; q < ntask; q += nt) {
    const int r = w.mlist[q / w.Wd], k = q % w.Wd, t = q;      // slot * Wd + k
    const int y = w.y0 + r, xw = w.x0w + k;
    uint32_t m = 0;
    if (w.fmt == 1) {
      m = (xw < w.bit_words) ? w.bits[(size_t)y * w.bit_words + xw] : 0u;
      const int rem = w.W - 32 * xw;
      if (rem < 32) m &= (rem <= 0) ? 0u : (0xffffffffu >> (32 - rem));
    } else {
      const uint8_t* p = w.px + (size_t)y * w.W + 32 * xw;
      const int rem = w.W - 32 * xw;
      if (rem >= 32 && (w.W & 3) == 0 && (reinterpret_cast<uintptr_t>(w.px) & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t u = q[j];
          u |= u >> 4; u |= u >> 2; u |= u >> 1;
          u &= 0x01010101u;                                   // byte != 0 -> 1
          m |= ((u * 0x01020408u) >> 24) << (4 * j);          // 4 byte flags -> 4 bits
        }
      } else {
        VA_ROLL
        for (int j = 0; j < 32 && j < rem; ++j) m |= (uint32_t)(p[j] != 0) << j;
      }
    }
    w.Mfg[t] = m;
    w.G[t] = m;
  }
}

// ---- phase 1b: the light check.  Most masks that miss the certificate are row-convex except for a few rows with a
//      notch (two or more runs).  When every row is non-empty, consecutive single-run rows touch, and every band of
//      adjacent rows with several runs passes the test spelled out below, the mask is ONE component WITHOUT holes: no
//      run table, union-find or lattice rebuild is needed, the table sums over the border pixels (phase_sums) give
//      points and area, the pixel bounding box is the component's.
//      Clears sc[W_LIGHT] otherwise; the full path then continues from phase_count. ----
constexpr int kLightRuns = 8;      // runs per row the light check follows
constexpr int kLightBand = 16;     // adjacent rows with several runs it follows ...
constexpr int kLightWords = 256;   // ... as long as the band has at most this many bit words (one thread walks it)
// runs of a row with several runs -> starts / ends (region-relative pixels); their number, or -1 when more than kLightRuns
VA_HD int light_row_runs(const Work& w, int r, int* rs, int* re) {
  const int base = slot_base(w, r);
  int n = 0;
  bool open = false;
  uint32_t carry = 0u;
  VA_ROLL
  for (int k = 0; k <= w.Wd; ++k) {                     // one word past the end closes a run that reaches the last pixel
    const uint32_t m = (k < w.Wd) ? w.Mfg[base + k] : 0u;
    uint32_t edges = m ^ ((m << 1) | carry);            // bit x: pixel x differs from pixel x - 1
    carry = m >> 31;
    VA_ROLL
    while (edges) {
      const int x = 32 * k + ffs32(edges) - 1;
      edges &= edges - 1;
      if (!open) {
        if (n == kLightRuns) return -1;
        rs[n] = x; open = true;
      } else {
        re[n++] = x - 1; open = false;
      }
    }
  }
  return n;
}
VA_HD void phase_light_check(Work& w, int tid, int nt) {
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {
    const int a = w.one_a[r];
    bool ok = true;
    if (a == kRowEmpty) {
      ok = false;
    } else if (a >= 0) {
      if (r > 0 && w.one_a[r - 1] >= 0) ok = a - 1 <= (int)w.one_b[r - 1] && (int)w.one_b[r] + 1 >= (int)w.one_a[r - 1];
    } else if (!row_is_multi(w, r - 1)) {
      // first row of a band of adjacent rows with several runs (a notch is usually a few rows deep).  The band is
      // followed only while it stays "parallel": the same number of runs in every row, run j touching run j of the
      // row above and gap j sharing a column with gap j of the row above - so run j forms a vertical strip and gap j a
      // vertical slit.  Then: a strip belongs to the component iff it touches the run above the band or the run
      // below it; a slit is open (not a hole) iff the run above or the run below does not cover its end; the parts
      // above and below the band meet iff some strip touches both.  Anything else goes to the full path.
      const Span u = row_span(w, r - 1);
      const bool hasu = u.a <= u.b;
      int depth = 1;                                    // size the band first: a deep or wide one is the full path's job
      VA_ROLL
      while (depth <= kLightBand && row_is_multi(w, r + depth)) ++depth;
      int ps[kLightRuns], pe[kLightRuns], cs[kLightRuns], ce[kLightRuns];
      const int m = (depth <= kLightBand && depth * w.Wd <= kLightWords) ? light_row_runs(w, r, ps, pe) : -1;
      ok = m >= 2;
      unsigned tu = 0u, open_top = 0u;                  // bit j: strip j touches the run above / slit j is open at the top
      VA_ROLL
      for (int j = 0; j < m && ok; ++j) {
        if (hasu && ps[j] - 1 <= u.b && pe[j] + 1 >= u.a) tu |= 1u << j;
        if (j + 1 < m && !(hasu && u.a <= pe[j] + 1 && u.b >= ps[j + 1] - 1)) open_top |= 1u << j;
      }
      int rr = r;
      VA_ROLL
      while (ok && row_is_multi(w, rr + 1)) {            // follow the band downwards
        ++rr;
        if (rr - r >= kLightBand || light_row_runs(w, rr, cs, ce) != m) { ok = false; break; }
        VA_ROLL
        for (int j = 0; j < m; ++j) {
          if (!(cs[j] <= pe[j] + 1 && ce[j] >= ps[j] - 1)) ok = false;                                    // strip j continues
          if (j + 1 < m && imax(ce[j] + 1, pe[j] + 1) > imin(cs[j + 1] - 1, ps[j + 1] - 1)) ok = false;   // slit j continues
        }
        VA_ROLL
        for (int j = 0; j < m; ++j) { ps[j] = cs[j]; pe[j] = ce[j]; }
      }
      if (ok) {
        const Span d = row_span(w, rr + 1);
        const bool hasd = d.a <= d.b;
        bool both = false;
        VA_ROLL
        for (int j = 0; j < m; ++j) {
          const bool td = hasd && ps[j] - 1 <= d.b && pe[j] + 1 >= d.a;
          const bool tuj = (tu >> j) & 1u;
          if (!tuj && !td) ok = false;                  // a strip that hangs in the air: another component
          both = both || (tuj && td);
          if (j + 1 < m && !((open_top >> j) & 1u) && hasd && d.a <= pe[j] + 1 && d.b >= ps[j + 1] - 1) ok = false;   // a hole
        }
        if (hasu && hasd && !both) ok = false;
      }
    }
    if (!ok) w.sc[W_LIGHT] = 0;
  }
}
// ---- phase 1c (light path only): one component, no holes - the state the later phases read ----
VA_HD void phase_light_setup(Work& w, int tid, int minx, int maxx) {
  if (tid != 0) return;
  w.sc[W_NR] = 1; w.sc[W_ROOTS] = 1; w.sc[W_HOLES] = 0; w.sc[W_CHOSEN] = 0;
  w.sc[W_MINX] = minx - 32 * w.x0w; w.sc[W_MAXX] = maxx - 32 * w.x0w; w.sc[W_MINY] = 0; w.sc[W_MAXY] = w.R - 1;
  w.pF[0] = 0; w.accP[0] = 0; w.accA[0] = 0;
}
// ---- phase 2: runs per word / row ----
VA_HD void phase_count(Work& w, int tid, int nt) {
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {
    const int a = w.one_a[r];
    int acc = 0;
    if (a != kRowMulti) {
      acc = (a >= 0) ? 1 : 0;
    } else {
      const int base = slot_base(w, r);
      uint32_t carry = 0u;                              // last pixel of the previous word
      VA_ROLL
      for (int k = 0; k < w.Wd; ++k) {
        const uint32_t m = w.Mfg[base + k];
        w.S[base + k] = (uint16_t)acc;
        acc += popc32(m & ~((m << 1) | carry));
        carry = m >> 31;
      }
    }
    w.rowoff[r] = acc;
  }
}
// ---- phase 3a-c: exclusive scan of the row counts (32 segments) ----
VA_HD void phase_scan_a(Work& w, int tid, int nt) {
  const int segl = (w.R + 31) / 32;
  VA_ROLL
  for (int s = tid; s < 32; s += nt) {
    int acc = 0;
    VA_ROLL
    for (int r = s * segl; r < imin((s + 1) * segl, w.R); ++r) acc += w.rowoff[r];
    w.seg[s] = acc;
  }
}
VA_HD void phase_scan_b(Work& w, int tid, int nt) {
  (void)nt;
  if (tid == 0) {
    int acc = 0;
    VA_ROLL
    for (int s = 0; s < 32; ++s) { const int v = w.seg[s]; w.seg[s] = acc; acc += v; }
    w.seg[32] = acc;
    w.sc[W_NR] = acc;
    if (acc > w.cap) w.sc[W_OVERFLOW] = 1;
  }
}
VA_HD void phase_scan_c(Work& w, int tid, int nt) {
  const int segl = (w.R + 31) / 32;
  VA_ROLL
  for (int s = tid; s < 32; s += nt) {
    int acc = w.seg[s];
    VA_ROLL
    for (int r = s * segl; r < imin((s + 1) * segl, w.R); ++r) { const int v = w.rowoff[r]; w.rowoff[r] = acc; acc += v; }
  }
  if (tid == 0) w.rowoff[w.R] = w.seg[32];
}

#if defined(__CUDACC__)
// phases 3a-c on the GPU, two barrier-separated steps: consecutive rows per thread, one shuffle scan per warp and the
// warp totals in seg[]; then every warp scans the totals again for its own offset
__device__ __forceinline__ void phase_scan_rows_a(Work& w, int tid, int nt, int& first, int& excl) {
  const int per = (w.R + nt - 1) / nt;
  const int r0 = imin(tid * per, w.R), r1 = imin(r0 + per, w.R);
  int acc = 0;
  VA_ROLL
  for (int r = r0; r < r1; ++r) acc += w.rowoff[r];
  int incl = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += u; }
  if ((tid & 31) == 31) w.seg[tid >> 5] = incl;       // nt <= 1024: at most 32 warps, seg holds 33 entries
  first = r0;
  excl = incl - acc;                                   // exclusive prefix inside the warp
}
__device__ __forceinline__ void phase_scan_rows_b(Work& w, int tid, int nt, int first, int excl) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int mine = lane < nwarps ? w.seg[lane] : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int base = __shfl_sync(0xffffffffu, incl - mine, warp);   // sum of the warps below this one
  const int per = (w.R + nt - 1) / nt;
  int run = base + excl;
  VA_ROLL
  for (int r = first; r < imin(first + per, w.R); ++r) { const int v = w.rowoff[r]; w.rowoff[r] = run; run += v; }
  if (tid == 0) {
    w.rowoff[w.R] = total;
    w.sc[W_NR] = total;
    if (total > w.cap) w.sc[W_OVERFLOW] = 1;
  }
}
#endif

VA_HD void run_init(Work& w, int id, int start, int r, bool last_in_row) {
  w.rs[id] = (uint16_t)start;
  w.ry[id] = (uint16_t)r;
  w.pF[id] = id;
  w.pG[id + 1] = last_in_row ? 0 : id + 1;           // the gap after the last run of a row is the outside
  w.accP[id] = 0;
  w.accA[id] = 0;
}
// ---- phase 4: run table.  The j-th run start of a row pairs with its j-th run end, so starts and ends are written
//      independently word by word (no search for the end of a run) ----
VA_HD void phase_runs(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  if (tid == 0) w.pG[0] = 0;                          // the outside
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {               // rows with one run
    const int a = w.one_a[r];
    if (a < 0) continue;
    const int id = w.rowoff[r];
    run_init(w, id, a, r, true);
    w.re[id] = (uint16_t)w.one_b[r];
  }
  const int ntask = w.sc[W_NM] * w.Wd;
  VA_ROLL
  for (int q = tid; q < ntask; q += nt) {              // rows with several runs
    const int r = w.mlist[q / w.Wd], k = q % w.Wd, t = q;      // slot * Wd + k
    const int base = w.rowoff[r], last_id = w.rowoff[r + 1] - 1;
    uint32_t rise = rise_at(w, r, k);
    int id = base + (int)w.S[t];
    VA_ROLL
    while (rise) {
      const int b = ffs32(rise) - 1;
      rise &= rise - 1;
      run_init(w, id, 32 * k + b, r, id == last_id);
      ++id;
    }
    // run ends: as many ends as starts lie before this word, minus the run still open at the word's left edge
    uint32_t fall = fall_at(w, r, k);
    const int open = (k > 0) ? (int)(w.Mfg[t - 1] >> 31) & (int)(w.Mfg[t] & 1u) : 0;
    int ie = base + (int)w.S[t] - open;
    VA_ROLL
    while (fall) {
      const int b = ffs32(fall) - 1;
      fall &= fall - 1;
      w.re[ie] = (uint16_t)(32 * k + b);
      ++ie;
    }
  }
}

// ---- phase 5: background gaps, 4-connectivity; gaps that reach the frame / region border join node 0 ----
VA_HD void phase_gaps(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  VA_ROLL
  for (int t = tid; t < 2 * NR; t += nt) {              // (gap, direction) tasks: the two neighbour rows side by side
    const int id = t >> 1;
    const int r = w.ry[id];
    if (id == w.rowoff[r + 1] - 1) continue;            // no gap to the right
    const int g0 = w.re[id] + 1, g1 = w.rs[id + 1] - 1;
    {
      const int rr = r + ((t & 1) ? 1 : -1);
      if (rr < 0 || rr >= w.R) { uf_union(w.pG, id + 1, 0); continue; }
      const int n2 = row_runs(w, rr), o2 = w.rowoff[rr];
      int xa = g0, xb = g1;
      if (fg_at(w, rr, xa)) xa = w.re[o2 + ns(w, rr, xa) - 1] + 1;       // first background pixel >= g0 in row rr
      if (xa > xb) continue;
      if (fg_at(w, rr, xb)) xb = w.rs[o2 + ns(w, rr, xb) - 1] - 1;       // last background pixel <= g1
      if (xa > xb) continue;
      const int qa = ns(w, rr, xa), qb = ns(w, rr, xb);                  // gap q lies between runs q-1 and q
      VA_ROLL
      for (int q = qa; q <= qb; ++q) uf_union(w.pG, id + 1, (q == 0 || q == n2) ? 0 : o2 + q);
    }
  }
}
// ---- phase 6: holes known -> runs separated by a hole belong together ----
VA_HD void phase_holes(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  VA_ROLL
  for (int id = tid; id < NR; id += nt) {
    const int r = w.ry[id];
    if (id == w.rowoff[r + 1] - 1) continue;
    if (uf_find(w.pG, id + 1) != 0) {
      uf_union(w.pF, id, id + 1);
      atom_add(&w.sc[W_HOLES], 1);
      // fill the hole pixels into G
      const int g0 = w.re[id] + 1, g1 = w.rs[id + 1] - 1;
      VA_ROLL
      for (int k = g0 >> 5; k <= (g1 >> 5); ++k) atom_or(&w.G[slot_base(w, r) + k], span_bits(g0, g1, k));
    }
  }
}
// ---- phase 7: foreground runs, 8-connectivity with the row above ----
// all runs of the row above that touch run id (its pixels and the two diagonal neighbours)
VA_HD void link_row_above(Work& w, int id, int r) {
  const int lo = (int)w.rs[id] - 1, hi = imin((int)w.re[id] + 1, 32 * w.Wd - 1);
  const int o2 = w.rowoff[r - 1];
  const int jlo = (lo >= 0 && fg_at(w, r - 1, lo)) ? ns(w, r - 1, lo) - 1 : ns(w, r - 1, lo);
  const int jhi = ns(w, r - 1, hi) - 1;
  VA_ROLL
  for (int j = jlo; j <= jhi; ++j) uf_union(w.pF, id, o2 + j);
}
VA_HD void phase_link(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
#ifdef __CUDA_ARCH__
  if ((nt & 31) == 0) {
    // A mask is mostly a vertical stack of single-run rows: run id sits on run id - 1.  Linking every run to the one
    // above at the same time leaves one list as long as the mask is tall, which the finds then walk.  Within a warp
    // (32 consecutive ids) the stack links are resolved by a vote instead: every run of a stretch of stack links is
    // united with the stretch's first run directly, only the first run of a stretch looks at the row above.
    const int lane = tid & 31;
    VA_ROLL
    for (int base = 0; base < NR; base += nt) {
      const int id = base + tid;
      const bool valid = id < NR;
      const int r = valid ? (int)w.ry[id] : 0;
      bool stack = false;
      if (valid && r > 0) {
        const int a = w.one_a[r], ap = w.one_a[r - 1];
        stack = a >= 0 && ap >= 0 && a - 1 <= (int)w.one_b[r - 1] && (int)w.one_b[r] + 1 >= ap;   // both rows one run, touching
      }
      const unsigned heads = __ballot_sync(0xffffffffu, !stack || lane == 0);
      const int head_lane = 31 - clz32(heads & (0xffffffffu >> (31 - lane)));
      if (!valid || r == 0) continue;
      if (lane != head_lane) uf_union(w.pF, id, id - (lane - head_lane));
      else link_row_above(w, id, r);
    }
    return;
  }
#endif
  // consecutive ids per thread: a strided second pass would start from the id - 1 chains the first pass left behind
  // (a vertical stack of runs links into one list as long as the mask is tall) and walk them alone
  const int per = (NR + nt - 1) / nt;
  VA_ROLL
  for (int id = tid * per; id < imin(NR, (tid + 1) * per); ++id) {
    const int r = w.ry[id];
    if (r == 0) continue;
    link_row_above(w, id, r);
  }
}
// ---- phase 8: flatten, in two barrier-separated steps: the finds of step a still re-point nodes (path halving) and
//      would undo another thread's flattened entry, so the roots are parked in accA first ----
VA_HD void phase_flatten_a(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  VA_ROLL
  for (int id = tid; id < NR; id += nt) w.accA[id] = uf_find(w.pF, id);
}
VA_HD void phase_flatten_b(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  VA_ROLL
  for (int id = tid; id < NR; id += nt) {
    const int root = w.accA[id];
    if (root == id) atom_add(&w.sc[W_ROOTS], 1);
    w.pF[id] = root;
    w.accA[id] = 0;
  }
}
// ---- phase 9: table sums over the border pixels of G ----
// Sums are kept per thread across all its work and flushed when the component changes; on the GPU the last flush
// is aggregated over the warp first (with one component per mask every thread would hit the same two words).
VA_HD void sums_flush(Work& w, int root, int pts, int a2) {
  if (root >= 0) { atom_add(&w.accP[root], pts); atom_add(&w.accA[root], a2); }
}
VA_HD uint32_t span_code(const Span& u, const Span& c, const Span& d, int x) {
  auto in = [](const Span& s, int v) -> uint32_t { return (v >= s.a && v <= s.b) ? 1u : 0u; };
  return in(u, x - 1) | (in(u, x) << 1) | (in(u, x + 1) << 2) | (in(c, x - 1) << 3) | (in(c, x + 1) << 4) |
         (in(d, x - 1) << 5) | (in(d, x) << 6) | (in(d, x + 1) << 7);
}
// per-thread running sums, flushed to the component's accumulators when the component changes
struct SumAcc {
  int root, pts, a2;
  VA_HD SumAcc() : root(-1), pts(0), a2(0) {}
  VA_HD void add(Work& w, int r, int p, int v) {
    if (r != root) { sums_flush(w, root, pts, a2); root = r; pts = 0; a2 = 0; }
    pts += p; a2 += v;
  }
  // last flush: aggregated over the warp when all its threads hold the same component (the usual case)
  VA_HD void finish(Work& w, int tid) {
#ifdef __CUDA_ARCH__
    __syncwarp();
    const int rmax = __reduce_max_sync(0xffffffffu, root);
    if (__all_sync(0xffffffffu, root == rmax || root < 0)) {
      const int sp = (int)__reduce_add_sync(0xffffffffu, (unsigned)(root >= 0 ? pts : 0));
      const int sa = (int)__reduce_add_sync(0xffffffffu, (unsigned)(root >= 0 ? a2 : 0));
      if ((tid & 31) == 0) sums_flush(w, rmax, sp, sa);
      return;
    }
#endif
    (void)tid;
    sums_flush(w, root, pts, a2);
  }
};
// table terms of pixel x of plain row r added to (p, v); a pixel the border does not visit adds zeros
VA_HD void plain_term(const uint16_t* lut, const PlainRanges& g, int r, int x, int& p, int& v) {
  const uint32_t e = lut[span_code(g.u, g.c, g.d, x)];
  const int dxs = (int)((e >> 3) & 7u) - 2, dys = (int)((e >> 6) & 7u) - 2;
  p += (int)(e & 7u);
  v += x * dys - r * dxs;
}
VA_HD void plain_pixel(Work& w, const uint16_t* lut, const PlainRanges& g, int r, int root, int j, SumAcc& acc) {
  int p = 0, v = 0;
  plain_term(lut, g, r, (j < g.n1) ? g.c.a + j : g.lb + (j - g.n1), p, v);
  if (p | v) acc.add(w, root, p, v);
}
// a word of G and its eight neighbour words, then the six shifted images whose bit b is a neighbour of pixel b
struct WordNb {
  uint32_t Up, U, Un, Mp, M, Mn, Dp, D, Dn, Ul, Ur, Ml, Mr, Dl, Dr;
  VA_HD void shift() {
    Ul = (U << 1) | (Up >> 31); Ur = (U >> 1) | (Un << 31);
    Ml = (M << 1) | (Mp >> 31); Mr = (M >> 1) | (Mn << 31);
    Dl = (D << 1) | (Dp >> 31); Dr = (D >> 1) | (Dn << 31);
  }
  VA_HD uint32_t border() const { return M & ~(Ul & U & Ur & Ml & Mr & Dl & D & Dr); }
  VA_HD uint32_t code(int b) const {
    return ((Ul >> b) & 1u) | (((U >> b) & 1u) << 1) | (((Ur >> b) & 1u) << 2) | (((Ml >> b) & 1u) << 3) |
           (((Mr >> b) & 1u) << 4) | (((Dl >> b) & 1u) << 5) | (((D >> b) & 1u) << 6) | (((Dr >> b) & 1u) << 7);
  }
};
// the run a pixel of word (r, k) belongs to: the row's only run, or the run starts up to the pixel (it lies in that
// run or in the hole after it) - fetched once per word
struct WordRuns {
  bool multi; const int* pF; int id0, before, root1; uint32_t rise;
  VA_HD WordRuns(const Work& w, int r, int k) {
    if (w.sc[W_LIGHT]) {                               // light path: one component, run 0 stands for it
      multi = false; pF = w.pF; id0 = 0; rise = 0u; before = 0; root1 = 0;
      return;
    }
    multi = w.one_a[r] == kRowMulti;
    pF = w.pF;
    id0 = w.rowoff[r];
    rise = multi ? rise_at(w, r, k) : 0u;
    before = multi ? (int)w.S[slot_base(w, r) + k] : 0;
    root1 = multi ? -1 : pF[id0];
  }
  VA_HD int root(int b) const { return multi ? pF[id0 + before + popc32(rise & (0xffffffffu >> (31 - b))) - 1] : root1; }
};
// the plain rows with long candidate ranges, one row per group of 32 threads (warps from the back: the word tasks
// below are dealt from the front)
VA_HD void sums_long_rows(Work& w, const uint16_t* lut, int tid, int nt, SumAcc& acc) {
  const int L = (nt % 32 == 0) ? 32 : 1;
  const int nlong = w.sc[W_NLONG], ngroups = nt / L;
  VA_ROLL
  for (int q = ngroups - 1 - tid / L; q < nlong; q += ngroups) {
    const int r = w.nplist[w.R - 1 - q];
    const PlainRanges g = plain_ranges(w, r);
    const int root = w.sc[W_LIGHT] ? 0 : w.pF[w.rowoff[r]];
    VA_ROLL
    for (int j = tid % L; j < g.n1 + g.n2; j += L) plain_pixel(w, lut, g, r, root, j, acc);
  }
}
VA_HD void word_pixel(Work& w, const uint16_t* lut, const WordNb& nb, const WordRuns& wr, int r, int k, int b, SumAcc& acc) {
  const uint32_t e = lut[nb.code(b)];
  const int p = (int)(e & 7u), dxs = (int)((e >> 3) & 7u) - 2, dys = (int)((e >> 6) & 7u) - 2;
  if (!(p | dxs | dys)) return;                         // a hole pixel touching the outside diagonally: no visit
  acc.add(w, wr.root(b), p, (32 * k + b) * dys - r * dxs);    // region-relative coordinates: the area is translation invariant
}
VA_HD void phase_sums(Work& w, const uint16_t* lut, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  SumAcc acc;
  // (a) plain rows.  All pixels of [a, b] that miss a neighbour lie in [a, la] and [lb, b] with la = max(a, ua, da) + 1,
  //     lb = min(b, ub, db) - 1 (the whole run when a neighbour row is empty).  Slanted outlines make both ranges a
  //     few pixels long: one thread per row.  Rows with long ranges (flat edges) are listed at the back of nplist
  //     and summed by phase_sums_long, 32 threads per row.
  VA_ROLL
  for (int r = tid; r < w.R; r += nt) {
    if (!row_is_plain(w, r)) continue;
    const PlainRanges g = plain_ranges(w, r);
    if (g.n1 + g.n2 > kPlainInline) continue;          // on the long list (phase_lists)
    const int root = w.sc[W_LIGHT] ? 0 : w.pF[w.rowoff[r]];
    int p1 = 0, v1 = 0, p2 = 0, v2 = 0;                // the two ranges side by side: two independent dependency chains
    VA_ROLL
    for (int j = 0; j < imax(g.n1, g.n2); ++j) {
      if (j < g.n1) plain_term(lut, g, r, g.c.a + j, p1, v1);
      if (j < g.n2) plain_term(lut, g, r, g.lb + j, p2, v2);
    }
    acc.add(w, root, p1 + p2, v1 + v2);
  }
  sums_long_rows(w, lut, tid, nt, acc);
  // (b) every other non-empty row (rows with several runs and their neighbours), word by word on the bit image
  const int ntask = w.sc[W_NNP] * w.Wd;
#ifdef __CUDA_ARCH__
  if ((nt & 31) == 0 && ntask <= 2 * (nt >> 3)) {
    // few words (notches in an otherwise row-convex outline): one word per group of 8 lanes and four pixels per lane -
    // a single thread walking the up to 32 border pixels of its word would be the whole mask's critical path.  The
    // eight lanes fetch the 3x3 words around the task's word (lane 0 two of them) and exchange them by shuffles.
    const int lane = tid & 31, gl = lane & 7;
    const unsigned gmask = 0xffu << (lane & 24);
    VA_ROLL
    for (int q = tid >> 3; q < ntask; q += nt >> 3) {
      const int r = w.nplist[q / w.Wd], k = q % w.Wd;
      const uint32_t mine = word_at(w.G, w, r - 1 + gl / 3, k - 1 + gl % 3);           // words 0..7 of the 3x3 block
      const uint32_t last = (gl == 0) ? word_at(w.G, w, r + 1, k + 1) : 0u;            // word 8
      WordNb nb;
      nb.Up = __shfl_sync(gmask, mine, 0, 8); nb.U = __shfl_sync(gmask, mine, 1, 8); nb.Un = __shfl_sync(gmask, mine, 2, 8);
      nb.Mp = __shfl_sync(gmask, mine, 3, 8); nb.M = __shfl_sync(gmask, mine, 4, 8); nb.Mn = __shfl_sync(gmask, mine, 5, 8);
      nb.Dp = __shfl_sync(gmask, mine, 6, 8); nb.D = __shfl_sync(gmask, mine, 7, 8); nb.Dn = __shfl_sync(gmask, last, 0, 8);
      if (!nb.M) continue;
      nb.shift();
      const uint32_t border = nb.border() & (0x01010101u << gl);                       // this lane's pixels: gl, gl + 8, ...
      if (!border) continue;
      const WordRuns wr(w, r, k);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((border >> (gl + 8 * j)) & 1u) word_pixel(w, lut, nb, wr, r, k, gl + 8 * j, acc);
    }
    acc.finish(w, tid);
    return;
  }
#endif
  VA_ROLL
  for (int q = tid; q < ntask; q += nt) {
    const int r = w.nplist[q / w.Wd], k = q % w.Wd;
    WordNb nb;
    nb.M = word_at(w.G, w, r, k);
    if (!nb.M) continue;
    nb.U = word_at(w.G, w, r - 1, k); nb.D = word_at(w.G, w, r + 1, k);
    nb.Up = word_at(w.G, w, r - 1, k - 1); nb.Un = word_at(w.G, w, r - 1, k + 1);
    nb.Mp = word_at(w.G, w, r, k - 1); nb.Mn = word_at(w.G, w, r, k + 1);
    nb.Dp = word_at(w.G, w, r + 1, k - 1); nb.Dn = word_at(w.G, w, r + 1, k + 1);
    nb.shift();
    uint32_t border = nb.border();
    if (!border) continue;
    const WordRuns wr(w, r, k);
    VA_ROLL
    while (border) {
      const int b = ffs32(border) - 1;
      border &= border - 1;
      word_pixel(w, lut, nb, wr, r, k, b, acc);
    }
  }
  acc.finish(w, tid);
}
// ---- phase 10: the component whose contour has the most points; ties: the last in raster order ----
// the lattice samples the mask kernel left are those of the kept component when the mask is one hole-free component
VA_HD bool lattice_stands(const Work& w) { return !w.sc[W_OVERFLOW] && w.sc[W_ROOTS] == 1 && w.sc[W_HOLES] == 0; }
VA_HD void phase_select(Work& w, int tid, int nt) {
  if (!lattice_stands(w)) {
    VA_ROLL
    for (int t = tid; t < w.lat_rows * w.lat_words; t += nt) w.lattice[t] = 0u;   // rebuilt by phase_output
  }
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  unsigned long long best = 0ull;
  bool any = false;
  VA_ROLL
  for (int id = tid; id < NR; id += nt) {
    if (w.pF[id] != id) continue;
    const unsigned long long key = ((unsigned long long)(unsigned)w.accP[id] << 32) | (unsigned)(id + 1);
    if (!any || key > best) { best = key; any = true; }
  }
  if (any) atom_max64(w.best, best);
}
// ---- phase 11: bounding box of the kept component ----
VA_HD void phase_bbox(Work& w, int tid, int nt) {
  if (w.sc[W_OVERFLOW]) return;
  const int NR = w.sc[W_NR];
  const int chosen = (int)(unsigned)(w.best[0] & 0xffffffffull) - 1;
  if (tid == 0) w.sc[W_CHOSEN] = chosen;
  if (chosen < 0) return;
  int minx = 1 << 30, miny = 1 << 30, maxx = -1, maxy = -1;
  VA_ROLL
  for (int id = tid; id < NR; id += nt) {
    if (w.pF[id] != chosen) continue;
    minx = imin(minx, (int)w.rs[id]); maxx = imax(maxx, (int)w.re[id]);
    miny = imin(miny, (int)w.ry[id]); maxy = imax(maxy, (int)w.ry[id]);
  }
  if (maxx >= 0) {
    atom_min_nr(&w.sc[W_MINX], minx); atom_max(&w.sc[W_MAXX], maxx);
    atom_min_nr(&w.sc[W_MINY], miny); atom_max(&w.sc[W_MAXY], maxy);
  }
}
// ---- phase 12: cell-centre lattice samples of the kept component (= the fillPoly raster) + result ----
VA_HD void phase_output(Work& w, int tid, int nt) {
  const int half = w.gs >> 1;
  const bool ok = !w.sc[W_OVERFLOW] && w.sc[W_CHOSEN] >= 0;
  const int chosen = w.sc[W_CHOSEN];
  if (ok && !lattice_stands(w)) {
    // lattice points inside the region only: rows ly0 .. ly1, columns lx0 .. lx1
    const int ly0 = imax(0, (w.y0 - half + w.gs - 1) / w.gs), ly1 = imin(w.lat_rows - 1, (w.y0 + w.R - 1 - half) / w.gs);
    const int lx0 = imax(0, (32 * w.x0w - half + w.gs - 1) / w.gs), lx1 = imin(w.lat_cols - 1, (32 * (w.x0w + w.Wd) - 1 - half) / w.gs);
    const int nly = ly1 - ly0 + 1, nlx = lx1 - lx0 + 1;
    if (w.y0 + w.R - 1 >= half && nly > 0 && nlx > 0) {
      VA_ROLL
      for (int t = tid; t < nly * nlx; t += nt) {
        const int ly = ly0 + t / nlx, lx = lx0 + t % nlx;
        const int r = w.gs * ly + half - w.y0;
        const int x = w.gs * lx + half - 32 * w.x0w;          // region-relative pixel
        if (r < 0 || r >= w.R || x < 0 || x >= 32 * w.Wd) continue;
        if (!((word_at(w.G, w, r, x >> 5) >> (x & 31)) & 1u)) continue;
        const int j = ns(w, r, x) - 1;                        // the run at or left of x (x is in it or in the hole after it)
        if (j >= 0 && w.pF[w.rowoff[r] + j] == chosen) atom_or(&w.lattice[ly * w.lat_words + (lx >> 5)], 1u << (lx & 31));
      }
    }
  }
  if (tid == 0) {
    InstContour o;
    o.area2 = 0; o.state = kEmpty; o.minx = 0; o.miny = 0; o.maxx = -1; o.maxy = -1; o.points = 0; o.n_components = 0;
    if (w.sc[W_OVERFLOW]) {
      o.state = kOverflow;
    } else if (chosen >= 0) {
      const int a = w.accA[chosen];
      o.area2 = a < 0 ? -a : a;
      o.points = w.accP[chosen];
      o.n_components = w.sc[W_ROOTS];
      o.state = (w.sc[W_ROOTS] == 1 && w.sc[W_HOLES] == 0) ? kSimple : kGeneral;
      o.minx = 32 * w.x0w + w.sc[W_MINX]; o.maxx = 32 * w.x0w + w.sc[W_MAXX];
      o.miny = w.y0 + w.sc[W_MINY]; o.maxy = w.y0 + w.sc[W_MAXY];
    }
    *w.out = o;
  }
}

// Scratch layout in three parts, each placed in shared memory when it fits and in a global slab otherwise: the row
// part (row classes, row offsets, row lists - touched by every phase), the word part (bit images and per-word run
// counts - only rows with several runs use it) and the run part (run table, union-find, sums).
struct RowLayout { size_t one_a, one_b, mlist, nplist, rowoff, seg, total; };
struct GridLayout { size_t Mfg, G, S, total; };
struct RunLayout { size_t rs, re, ry, pF, pG, accP, accA, total; };
VA_HD size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
VA_HD RowLayout row_layout(int R) {
  RowLayout l;
  size_t o = 0;
  l.one_a = o; o += align16(sizeof(int16_t) * R);
  l.one_b = o; o += align16(sizeof(int16_t) * R);
  l.mlist = o; o += align16(sizeof(int16_t) * R);
  l.nplist = o; o += align16(sizeof(int16_t) * R);
  l.rowoff = o; o += align16(sizeof(int) * (R + 1));
  l.seg = o; o += align16(sizeof(int) * 33);
  l.total = o;
  return l;
}
VA_HD GridLayout grid_layout(int R, int Wd) {
  GridLayout l;
  size_t o = 0;
  l.Mfg = o; o += align16(sizeof(uint32_t) * R * Wd);
  l.G = o; o += align16(sizeof(uint32_t) * R * Wd);
  l.S = o; o += align16(sizeof(uint16_t) * R * Wd);
  l.total = o;
  return l;
}
VA_HD RunLayout run_layout(int cap) {
  RunLayout l;
  size_t o = 0;
  l.rs = o; o += align16(sizeof(uint16_t) * cap);
  l.re = o; o += align16(sizeof(uint16_t) * cap);
  l.ry = o; o += align16(sizeof(uint16_t) * cap);
  l.pF = o; o += align16(sizeof(int) * cap);
  l.pG = o; o += align16(sizeof(int) * (cap + 1));
  l.accP = o; o += align16(sizeof(int) * cap);
  l.accA = o; o += align16(sizeof(int) * cap);
  l.total = o;
  return l;
}
VA_HD void bind_rows(Work& w, unsigned char* base, const RowLayout& l) {
  w.one_a = reinterpret_cast<int16_t*>(base + l.one_a); w.one_b = reinterpret_cast<int16_t*>(base + l.one_b);
  w.mlist = reinterpret_cast<int16_t*>(base + l.mlist); w.nplist = reinterpret_cast<int16_t*>(base + l.nplist);
  w.rowoff = reinterpret_cast<int*>(base + l.rowoff);
  w.seg = reinterpret_cast<int*>(base + l.seg);
}
VA_HD void bind_grid(Work& w, unsigned char* base, const GridLayout& l) {
  w.Mfg = reinterpret_cast<uint32_t*>(base + l.Mfg); w.G = reinterpret_cast<uint32_t*>(base + l.G);
  w.S = reinterpret_cast<uint16_t*>(base + l.S);
}
VA_HD void bind_runs(Work& w, unsigned char* base, const RunLayout& l) {
  w.rs = reinterpret_cast<uint16_t*>(base + l.rs); w.re = reinterpret_cast<uint16_t*>(base + l.re);
  w.ry = reinterpret_cast<uint16_t*>(base + l.ry);
  w.pF = reinterpret_cast<int*>(base + l.pF); w.pG = reinterpret_cast<int*>(base + l.pG);
  w.accP = reinterpret_cast<int*>(base + l.accP); w.accA = reinterpret_cast<int*>(base + l.accA);
}

}  // namespace cc
}  // namespace va
