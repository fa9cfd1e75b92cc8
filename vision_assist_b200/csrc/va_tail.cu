// Frame tail: instance selection -> occupancy grid -> penalty map -> top-edge peaks -> record.
//
// One CTA per frame.  Input is the per-instance reduction written by the mask kernels (pixel
// area, pixel bbox, cell-centre lattice bits); nothing here touches the full-resolution mask
// (except the optional Euler-number check).  Reference semantics reproduced bit-for-bit:
//
//   FrameProcessor._extract_grid_information   FrameProcessor.py:50-171   (bbox snap :79-83, centre
//       sampling :88-97, artificial band :126-165 incl. the `row_idx < len-1 ? replace : append`
//       duplicate-row / gap-compression / negative-index behaviour :162-165)
//   PenaltyCalculator._pre_compute_easy_segments :26-55, _calculate_segment_penalty :57-110,
//       calculate_penalty :112-142 (float64, the reference's operation order, no FMA contraction)
//   ProtrusionDetector._create_binary_image / _find_peak / __call__  :38-158, :419-535 in the
//       grid-level closed form (oracle/protrusion.py::peaks_closed_form)
//
//   utils.get_closest_grid_to_point as _find_paths calls it (FrameProcessor.py:236-239) and the grid_lookup
//       row table behind _create_graph (:184-207): path start / end cells and the implicit A* graph
//
// Rows are bit masks (32 columns per word): run extents along a row are clz/ffs operations, the
// vertical walks read one broadcast word per step.  Launched as a programmatic dependent of the mask kernel.
#include <climits>
#include <cmath>

#include <cstdio>
#include <cstdlib>

#include "va_common.cuh"
#include "va_contour_core.h"
#include "va_contour_lut.h"

namespace va {

constexpr int kTailThreads = 256;       // smallest CTA size (tuning aid VA_TAIL_THREADS); the launch uses kTailMaxThreads
constexpr int kTailMaxThreads = 1024;
constexpr size_t kContourGridSmem = 16 * 1024;   // contour step: bit rows of the rows with several runs

// developer diagnostic (VA_TAIL_TIMING=1): cycle stamps of block 0 at the phase boundaries, printed by the kernel
constexpr int kTailDebugFlag = 1 << 30;
constexpr int kTailTimelineFlag = 1 << 28;   // VA_TAIL_TIMING=2: one line per CTA (SM, start, end of the dependency wait, end)
__device__ long long g_tail_t[24];
__device__ __forceinline__ unsigned __smid() { unsigned v; asm volatile("mov.u32 %0, %smid;" : "=r"(v)); return v; }
#define TT(k) do { if ((d.flags & kTailDebugFlag) && (int)blockIdx.x == ((d.flags >> 8) & 0xfff) && (threadIdx.x == 0 || ((k) >= 100 && threadIdx.x == 64))) g_tail_t[(k) % 100] = clock64(); } while (0)

struct TailSmem {
  // "created rows" table: ids [0, 2*rmax)
  int* row_y;        // [T]
  int* row_attr;     // [T]
  int* row_ly;       // [T]      row_y / gs (filled by finish_record: the penalty cells use lookup-row units)
  unsigned* colocc;  // [cmax][plw] grid_lookup occupancy transposed: bit ly of column c (vertical runs by clz / ffs)
  unsigned* occ;     // [T][cwords]
  unsigned* art;     // [T][cwords]
  int* list_ids;     // [rmax]   FrameProcessor.grids (list order) -> created id
  int* plane_owner;  // [PL]     grid_lookup row y/gs -> created id, -1 = no such row
  int* erow_first;   // [rmax]   easy_rows[k]: first / last column, first = -1 when not easy
  int* erow_last;    // [rmax]
  int* ecol_first;   // [cmax]   easy_cols[c]: first / last LIST index
  int* ecol_last;    // [cmax]
  int* orphan_ids;   // [rmax]
  int* oflag;        // [T]      orphan marks; later: created id -> record row
  unsigned long long* best;   // [pmax + 1] closest-cell keys (block-wide search of large grids): 0 = start, 1.. = peaks
  int* sc;           // scalars, see enum
};
enum { S_FLAGS, S_SEL, S_X0, S_Y0, S_C, S_R, S_NORPH, S_NPEAKS, S_AREA, S_RM, S_MINX, S_MINY, S_MAXX, S_MAXY,
       S_EULER, S_NCREATED, S_USE_EASY, S_START, S_COUNT };

__host__ __device__ inline int plane_cap(const Dims& d) { return 2 * d.rmax; }

__host__ __device__ inline size_t tail_smem_layout(const Dims& d, TailSmem* s, unsigned char* base) {
  const int T = 2 * d.rmax, PL = plane_cap(d);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~size_t(15); return r; };
  const size_t o_y = take(sizeof(int) * T), o_a = take(sizeof(int) * T), o_ly = take(sizeof(int) * T);
  const size_t o_co = take(sizeof(unsigned) * d.cmax * ((PL + 31) / 32));
  const size_t o_occ = take(sizeof(unsigned) * T * d.cwords), o_art = take(sizeof(unsigned) * T * d.cwords);
  const size_t o_l = take(sizeof(int) * d.rmax), o_p = take(sizeof(int) * PL);
  const size_t o_ef = take(sizeof(int) * d.rmax), o_el = take(sizeof(int) * max(d.rmax, d.cwords));
  const size_t o_cf = take(sizeof(int) * d.cmax), o_cl = take(sizeof(int) * d.cmax);
  const size_t o_or = take(sizeof(int) * d.rmax), o_of = take(sizeof(int) * T), o_sc = take(sizeof(int) * S_COUNT);
  const size_t o_best = take(sizeof(unsigned long long) * (d.pmax + 1));
  if (s) {
    s->row_y = (int*)(base + o_y); s->row_attr = (int*)(base + o_a); s->row_ly = (int*)(base + o_ly); s->colocc = (unsigned*)(base + o_co);
    s->occ = (unsigned*)(base + o_occ); s->art = (unsigned*)(base + o_art);
    s->list_ids = (int*)(base + o_l); s->plane_owner = (int*)(base + o_p);
    s->erow_first = (int*)(base + o_ef); s->erow_last = (int*)(base + o_el);
    s->ecol_first = (int*)(base + o_cf); s->ecol_last = (int*)(base + o_cl);
    s->orphan_ids = (int*)(base + o_or); s->oflag = (int*)(base + o_of); s->sc = (int*)(base + o_sc);
    s->best = (unsigned long long*)(base + o_best);
  }
  return o;
}

size_t tail_smem_bytes(const Dims& d) { return tail_smem_layout(d, nullptr, nullptr); }

constexpr int kMaxColWords = 64;   // va_create rejects ceil(W/gs) > 2048

__device__ __forceinline__ bool bit_at(const unsigned* row, int c) { return (row[c >> 5] >> (c & 31)) & 1u; }

// first column of the run of set bits that contains c (bit c itself is not tested)
__device__ __forceinline__ int run_left(const unsigned* row, int c) {
  int w = c >> 5;
  // zeros strictly below c in word w
  unsigned z = ~row[w] & ((c & 31) ? (0xffffffffu >> (32 - (c & 31))) : 0u);
  VA_ROLL
  while (true) {
    if (z) return (w << 5) + (32 - __clz(z));
    if (w == 0) return 0;
    --w;
    z = ~row[w];
  }
}
// last column of the run of set bits that contains c; C = number of columns (bits >= C are 0)
__device__ __forceinline__ int run_right(const unsigned* row, int c, int C) {
  int w = c >> 5;
  const int nw = (C + 31) >> 5;
  unsigned z = ~row[w] & (((c & 31) == 31) ? 0u : (0xffffffffu << ((c & 31) + 1)));
  VA_ROLL
  while (true) {
    if (z) return min((w << 5) + (__ffs(z) - 1) - 1, C - 1);
    if (w == nw - 1) return C - 1;
    ++w;
    z = ~row[w];
  }
}

// PenaltyCalculator.py:98-110 with position and run ends in CELL units (m = pos - lo, den = hi - lo): the pixel
// coordinates the reference divides are these times gs, the same rational, hence the same correctly rounded
// double - taken from the host-built quotient table when both are in range (the usual case), divided otherwise.
__device__ __forceinline__ double seg_penalty(const Dims& d, int m, int den) {
  double ratio;
  if (den == 0) ratio = 0.5;
  else if ((unsigned)m <= (unsigned)d.ratio_n && (unsigned)den <= (unsigned)d.ratio_n) ratio = __ldg(d.ratio + (size_t)m * (d.ratio_n + 1) + den);
  else ratio = __ddiv_rn((double)m, (double)den);
  return __dmul_rn(2.0, fabs(__dsub_rn(ratio, 0.5)));
}

__device__ __forceinline__ double blend_penalty(double rp, double cp) {
  // PenaltyCalculator.py:127-142
  if (rp > 0.99 || cp > 0.99) return 1.0;
  const double tot = __dadd_rn(rp, cp);
  if (tot == 0.0) return 0.0;
  const double dom = __ddiv_rn(fabs(__dsub_rn(rp, cp)), tot);
  const double rw = __dadd_rn(0.5, (rp > cp) ? __dmul_rn(0.25, dom) : __dmul_rn(-0.25, dom));
  const double cw = __dsub_rn(1.0, rw);
  return __dadd_rn(__dmul_rn(rp, rw), __dmul_rn(cp, cw));
}

// ---------------------------------------------------------------------------------------------
// phases shared by the mask-driven and the grid-mode kernels (list / plane already in smem)
// ---------------------------------------------------------------------------------------------
__device__ void easy_segments(const Dims& d, const TailSmem& s) {
  const int R = s.sc[S_R], C = s.sc[S_C], cw = d.cwords;
  VA_ROLL
  for (int id = threadIdx.x; id < s.sc[S_NCREATED]; id += (int)blockDim.x) s.row_ly[id] = s.row_y[id] / d.gs;
  {
    // grid_lookup occupancy by column: the vertical traversal of PenaltyCalculator.py:73-95 becomes the same
    // clz / ffs run search as the horizontal one instead of a cell-by-cell walk
    const int PL = plane_cap(d), plw = (PL + 31) >> 5;
    // 32 x 32 bit blocks transposed with ballots: lane = lookup row of the block, bit q of its word = column q
    const int lane = threadIdx.x & 31, nwarps = (int)blockDim.x >> 5;
    VA_ROLL
    for (int blk = nwarps - 1 - ((int)threadIdx.x >> 5); blk < plw * cw; blk += nwarps) {   // last warps first: the first ones scan rows / columns below
      const int w = blk / cw, cwd = blk - w * cw;
      const int ly = 32 * w + lane;
      const int o = (ly < PL) ? s.plane_owner[ly] : -1;
      const unsigned word = (o >= 0) ? s.occ[(size_t)o * cw + cwd] : 0u;
      unsigned mine = 0;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, (word >> q) & 1u);
        if (lane == q) mine = m;
      }
      const int c = 32 * cwd + lane;
      if (c < d.cmax) s.colocc[(size_t)c * plw + w] = mine;
    }
  }
  const bool use = s.sc[S_USE_EASY] != 0;
  VA_ROLL
  for (int k = threadIdx.x; k < d.rmax; k += (int)blockDim.x) {
    int first = -1, last = -1, cnt = 0;
    if (use && k < R) {
      const unsigned* row = s.occ + (size_t)s.list_ids[k] * cw;
      VA_ROLL
      for (int w = 0; w < cw; ++w) {
        const unsigned v = row[w];
        if (v) {
          if (first < 0) first = (w << 5) + __ffs(v) - 1;
          last = (w << 5) + 31 - __clz(v);
          cnt += __popc(v);
        }
      }
      if (!(cnt > 0 && last - first == cnt - 1)) first = -1;
    }
    s.erow_first[k] = first;
    s.erow_last[k] = last;
  }
  VA_ROLL
  for (int c = threadIdx.x; c < d.cmax; c += (int)blockDim.x) {
    int first = -1, last = -1, cnt = 0;
    if (use && c < C) {
      VA_ROLL
      for (int k = 0; k < R; ++k) {
        if (bit_at(s.occ + (size_t)s.list_ids[k] * cw, c)) {
          if (first < 0) first = k;
          last = k;
          ++cnt;
        }
      }
      if (!(cnt > 0 && last - first == cnt - 1)) first = -1;
    }
    s.ecol_first[c] = first;
    s.ecol_last[c] = last;
  }
}

__device__ void penalties_and_record(const Dims& d, const TailSmem& s, uint8_t* rec, int tid, int nt) {
  const int R = s.sc[S_R], C = s.sc[S_C], cw = d.cwords, gs = d.gs, x0 = s.sc[S_X0];
  const int norph = s.sc[S_NORPH];
  const int PL = plane_cap(d);
  double* pen = reinterpret_cast<double*>(rec + d.off_penalty);
  uint8_t* occ_out = rec + d.off_occ;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int cells = d.rmax * d.cmax;
  // cell u = tid, tid + nt, ... decoded incrementally as (k, c0) - no division per cell - with the columns rotated by
  // the row, c = (c0 + k) mod cmax: a thread stride that is a multiple of cmax (224 = 7 * 32, 480 = 5 * 96) would
  // otherwise pin every thread to one column, and the cost of a cell depends on its column
  const int dk = nt / d.cmax, dc = nt - dk * d.cmax;
  int k = tid / d.cmax, c0 = tid - k * d.cmax, kmod = k % d.cmax;
  const int dkmod = dk % d.cmax;
  VA_ROLL
  for (int u = tid; u < cells; u += nt, k += dk, c0 += dc, kmod += dkmod) {
    if (c0 >= d.cmax) { c0 -= d.cmax; ++k; ++kmod; }
    if (kmod >= d.cmax) kmod -= d.cmax;
    if (kmod >= d.cmax) kmod -= d.cmax;
    int c = c0 + kmod;
    if (c >= d.cmax) c -= d.cmax;
    const int t = k * d.cmax + c;
    double p = qnan;
    uint8_t ob = 0;
    if (k < R + norph && c < C) {
      const int id = (k < R) ? s.list_ids[k] : s.orphan_ids[k - R];
      const unsigned* own = s.occ + (size_t)id * cw;
      const bool filled = bit_at(own, c);
      ob = (filled ? 1 : 0) | (bit_at(s.art + (size_t)id * cw, c) ? 2 : 0);
      if (filled && k < R) {
        const int attr = s.row_attr[id];
        const int ly = s.row_ly[id];                     // rows sit on multiples of gs (validated for caller-given grids)
        // ---- row direction (PenaltyCalculator.py:68-69, else :73-95 on grid_lookup), in column units ----
        int lo, hi;
        if (attr >= 0 && attr < R && s.erow_first[attr] >= 0) {
          lo = s.erow_first[attr];
          hi = s.erow_last[attr];
        } else {
          const unsigned* prow = s.occ + (size_t)s.plane_owner[ly] * cw;
          lo = run_left(prow, c);
          hi = run_right(prow, c, C);
        }
        const double rp = seg_penalty(d, c - lo, hi - lo);
        // ---- column direction, in lookup-row units ----
        if (s.ecol_first[c] >= 0) {
          lo = s.row_ly[s.list_ids[s.ecol_first[c]]];
          hi = s.row_ly[s.list_ids[s.ecol_last[c]]];
        } else {
          const unsigned* col = s.colocc + (size_t)c * ((PL + 31) >> 5);
          lo = run_left(col, ly);
          hi = run_right(col, ly, PL);
        }
        const double cp = seg_penalty(d, ly - lo, hi - lo);
        p = blend_penalty(rp, cp);
      }
    }
    pen[t] = p;
    occ_out[t] = ob;
  }
  int* ry = reinterpret_cast<int*>(rec + d.off_row_y);
  int* ra = reinterpret_cast<int*>(rec + d.off_row_attr);
  VA_ROLL
  for (int k = tid; k < d.rmax; k += nt) {
    int y = 0, a = 0;
    if (k < R + norph) {
      const int id = (k < R) ? s.list_ids[k] : s.orphan_ids[k - R];
      y = s.row_y[id];
      a = s.row_attr[id];
    }
    ry[k] = y;
    ra[k] = a;
  }
  // alignment padding: keep every byte of the record deterministic
  VA_ROLL
  for (int t = d.off_row_attr + 4 * d.rmax + tid; t < d.off_penalty; t += nt) rec[t] = 0;
  VA_ROLL
  for (int t = d.off_occ + d.rmax * d.cmax + tid; t < d.off_goals; t += nt) rec[t] = 0;
  VA_ROLL
  for (int t = d.off_lookup + 8 * d.rmax + tid; t < d.record_bytes; t += nt) rec[t] = 0;
}

// ProtrusionDetector closed form; executed by warp 0: one lane per list row finds the top-most occupied
// pixel row, the union of the rows painted on it is OR-reduced per 32-column word, lane 0 extracts the runs.
__device__ void find_peaks(const Dims& d, const TailSmem& s, uint8_t* rec) {
  int* peaks = reinterpret_cast<int*>(rec + d.off_peaks);
  const int R = s.sc[S_R], C = s.sc[S_C], cw = d.cwords, gs = d.gs, x0 = s.sc[S_X0];
  const int lane = threadIdx.x & 31;
  int ytop = INT_MAX;
  VA_ROLL
  for (int k = lane; k < R; k += 32) {
    const unsigned* row = s.occ + (size_t)s.list_ids[k] * cw;
    unsigned any = 0;
    VA_ROLL
    for (int w = 0; w < cw; ++w) any |= row[w];
    if (any) ytop = min(ytop, s.row_y[s.list_ids[k]]);
  }
  ytop = __reduce_min_sync(0xffffffffu, ytop);
  __shared__ unsigned uni[kMaxColWords];
  if (ytop != INT_MAX) {
    VA_ROLL
    for (int w = 0; w < cw; ++w) {
      unsigned v = 0;
      VA_ROLL
      for (int k = lane; k < R; k += 32)
        if (s.row_y[s.list_ids[k]] == ytop) v |= s.occ[(size_t)s.list_ids[k] * cw + w];
      v = __reduce_or_sync(0xffffffffu, v);
      if (lane == 0) uni[w] = v;
    }
  }
  __syncwarp();
  if (lane != 0) return;
  int np = 0;
  if (ytop != INT_MAX) {
    int c = 0;
    VA_ROLL
    while (c < C) {
      int w = c >> 5;                                   // next set bit at or after c
      unsigned v = uni[w] & (0xffffffffu << (c & 31));
      VA_ROLL
      while (!v && ++w < cw) v = uni[w];
      if (!v) break;
      c = (w << 5) + __ffs(v) - 1;
      if (c >= C) break;
      const int c1 = run_right(uni, c, C);
      const int xa = x0 + c * gs;
      const int xb = min(x0 + c1 * gs + gs, d.W - 1);
      if (np < d.pmax) {
        peaks[2 * np] = xa + (xb - xa + 1) / 2;
        peaks[2 * np + 1] = ytop;
      }
      ++np;
      c = c1 + 1;
    }
  }
  if (np > d.pmax) { s.sc[S_FLAGS] |= VA_FLAG_OVERFLOW; np = d.pmax; }
  VA_ROLL
  for (int q = np; q < d.pmax; ++q) { peaks[2 * q] = 0; peaks[2 * q + 1] = 0; }
  s.sc[S_NPEAKS] = np;
}

__device__ void write_header(const TailSmem& s, uint8_t* rec) {
  va_frame_header h;
  h.flags = s.sc[S_FLAGS]; h.sel = s.sc[S_SEL]; h.x0 = s.sc[S_X0]; h.y0 = s.sc[S_Y0];
  h.n_cols = s.sc[S_C]; h.n_rows = s.sc[S_R]; h.n_orphans = s.sc[S_NORPH]; h.n_peaks = s.sc[S_NPEAKS];
  h.area = s.sc[S_AREA]; h.n_mask_rows = s.sc[S_RM];
  h.minx = s.sc[S_MINX]; h.miny = s.sc[S_MINY]; h.maxx = s.sc[S_MAXX]; h.maxy = s.sc[S_MAXY];
  h.contour_area2 = s.sc[S_EULER]; h.start_cell = s.sc[S_START];
  *reinterpret_cast<va_frame_header*>(rec) = h;
}

// list rows no longer in the list but still owning their grid_lookup row, ordered by y.
// All threads mark, thread 0 gathers (a handful of rows at most).
__device__ void collect_orphans(const Dims& d, const TailSmem& s) {
  const int R = s.sc[S_R], n = s.sc[S_NCREATED];
  int* s_flags = s.oflag;
  VA_ROLL
  for (int id = threadIdx.x; id < n; id += (int)blockDim.x) s_flags[id] = 0;
  __syncthreads();
  VA_ROLL
  for (int k = threadIdx.x; k < R; k += (int)blockDim.x) s_flags[s.list_ids[k]] = 1;      // rows still in the list
  __syncthreads();
  VA_ROLL
  for (int id = threadIdx.x; id < n; id += (int)blockDim.x) {
    const bool in_list = s_flags[id] != 0;
    const int ly = s.row_y[id] / d.gs;
    s_flags[id] = (!in_list && ly >= 0 && ly < plane_cap(d) && s.plane_owner[ly] == id) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int no = 0;
    VA_ROLL
    for (int id = 0; id < n; ++id) {
      if (!s_flags[id] || R + no >= d.rmax) continue;
      int pos = no++;
      VA_ROLL
      while (pos > 0 && s.row_y[s.orphan_ids[pos - 1]] > s.row_y[id]) { s.orphan_ids[pos] = s.orphan_ids[pos - 1]; --pos; }
      s.orphan_ids[pos] = id;
    }
    s.sc[S_NORPH] = no;
  }
}

// Closest non-empty cell of one list row to the point (px, py): along a row the distance only depends on the column,
// so the candidates are the nearest set bit at or left of the point's column and the nearest one right of it
// (equal distances: the left one comes first in the reference's scan, the key order takes care of it).
__device__ __forceinline__ unsigned long long row_best_key(const unsigned* row, int cw, int C, int k, int cmax, int px, int py,
                                                           int x0, int y, int gs) {
  const int half = gs >> 1;
  int cfl = floor_div(px - x0 - half, gs);               // column whose centre is at or left of px
  cfl = max(-1, min(cfl, C - 1));
  int cl = -1, cr = -1;
  VA_ROLL
  for (int w = cfl >> 5; w >= 0 && cfl >= 0; --w) {      // nearest set bit <= cfl
    unsigned v = row[w];
    if (w == (cfl >> 5) && (cfl & 31) != 31) v &= (2u << (cfl & 31)) - 1u;
    if (v) { cl = (w << 5) + 31 - __clz(v); break; }
  }
  VA_ROLL
  for (int w = (cfl + 1) >> 5; w < cw; ++w) {            // nearest set bit > cfl
    unsigned v = row[w];
    if (w == ((cfl + 1) >> 5)) v &= 0xffffffffu << ((cfl + 1) & 31);
    if (v) { cr = (w << 5) + __ffs(v) - 1; break; }
  }
  const long long dy = py - (y + half);
  unsigned long long best = ~0ull;
  if (cl >= 0) {
    const long long dx = px - (x0 + cl * gs + half);
    best = ((unsigned long long)(dx * dx + dy * dy) << 32) | (unsigned)(k * cmax + cl);
  }
  if (cr >= 0 && cr < C) {
    const long long dx = px - (x0 + cr * gs + half);
    best = min(best, ((unsigned long long)(dx * dx + dy * dy) << 32) | (unsigned)(k * cmax + cr));
  }
  return best;
}

// SURVEY 8(f1), executed by warp 0 right after find_peaks (while the other warps compute penalty cells).
// Path start / end cells: utils.get_closest_grid_to_point (utils.py:6-32) as _find_paths calls it
// (FrameProcessor.py:236-239) - the non-empty LIST cell whose centre is closest to the point, first minimum in list
// order.  The reference compares np.sqrt of exact integers; the squared integer distances order the same way.
// One 64-bit key per candidate, (distance^2 << 32) | (k * cmax + c), minimised per lane and then over the warp.
// Also the grid_lookup row table that makes the _create_graph neighbourhood (:184-207) implicit in the record.
// `first_pt .. last_pt` of the points {0: path start (W/2, H), 1 + q: peak q}; one warp, lane = list row.
__device__ void closest_cells(const Dims& d, const TailSmem& s, uint8_t* rec, int first_pt, int last_pt) {
  const int lane = threadIdx.x & 31;
  const int R = s.sc[S_R], cw = d.cwords, gs = d.gs, x0 = s.sc[S_X0];
  const int* peaks = reinterpret_cast<const int*>(rec + d.off_peaks);     // written by lane 0 in find_peaks
  int* goals = reinterpret_cast<int*>(rec + d.off_goals);
  VA_ROLL
  for (int pt = first_pt; pt <= last_pt; ++pt) {
    const int px = pt ? peaks[2 * (pt - 1)] : d.W / 2, py = pt ? peaks[2 * (pt - 1) + 1] : d.H;
    unsigned long long best = ~0ull;
    VA_ROLL
    for (int k = lane; k < R; k += 32) {
      const int id = s.list_ids[k];
      best = min(best, row_best_key(s.occ + (size_t)id * cw, cw, s.sc[S_C], k, d.cmax, px, py, x0, s.row_y[id], gs));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
      const unsigned cell = (unsigned)best;
      const int gk = (best == ~0ull) ? -1 : (int)(cell / d.cmax), gc = (best == ~0ull) ? -1 : (int)(cell % d.cmax);
      if (pt == 0) s.sc[S_START] = (gk < 0) ? -1 : ((gk << 16) | gc);
      else { goals[2 * (pt - 1)] = gk; goals[2 * (pt - 1) + 1] = gc; }
    }
  }
}

// warp 0, after find_peaks: the end cell of every peak
__device__ void goals_for_peaks(const Dims& d, const TailSmem& s, uint8_t* rec) {
  const int lane = threadIdx.x & 31;
  __syncwarp();                                                           // lane 0's peaks / S_NPEAKS are visible
  const int npk = s.sc[S_NPEAKS];
  closest_cells(d, s, rec, 1, npk);
  int* goals = reinterpret_cast<int*>(rec + d.off_goals);
  VA_ROLL
  for (int q = npk + lane; q < d.pmax; q += 32) { goals[2 * q] = 0; goals[2 * q + 1] = 0; }
}

// warp 1 (independent of the peaks): the path start cell and the grid_lookup row table
__device__ void start_and_lookup(const Dims& d, const TailSmem& s, uint8_t* rec) {
  const int lane = threadIdx.x & 31;
  const int R = s.sc[S_R], norph = s.sc[S_NORPH], T = 2 * d.rmax, PL = plane_cap(d);
  closest_cells(d, s, rec, 0, 0);
  // created id -> record row (first list position, or R + j for the j-th orphan)
  VA_ROLL
  for (int t = lane; t < T; t += 32) s.oflag[t] = INT_MAX;
  __syncwarp();
  VA_ROLL
  for (int k = lane; k < R; k += 32) atomicMin(&s.oflag[s.list_ids[k]], k);
  VA_ROLL
  for (int j = lane; j < norph; j += 32) s.oflag[s.orphan_ids[j]] = R + j;
  __syncwarp();
  int* lookup = reinterpret_cast<int*>(rec + d.off_lookup);
  VA_ROLL
  for (int ly = lane; ly < PL; ly += 32) {
    const int owner = (R > 0) ? s.plane_owner[ly] : -1;
    const int v = (owner >= 0) ? s.oflag[owner] : -1;
    lookup[ly] = (v == INT_MAX) ? -1 : v;
  }
}

// The same for large grids (small cells / large frames), by the whole block after the penalty phase: one work item
// per (point, list row, 32-column word), keys minimised with shared-memory atomics.
__device__ void start_goals_lookup_block(const Dims& d, const TailSmem& s, uint8_t* rec) {
  const int R = s.sc[S_R], cw = d.cwords, gs = d.gs, x0 = s.sc[S_X0], half = gs >> 1;
  const int npk = s.sc[S_NPEAKS], P = 1 + npk, norph = s.sc[S_NORPH], T = 2 * d.rmax, PL = plane_cap(d);
  const int* peaks = reinterpret_cast<const int*>(rec + d.off_peaks);     // written by warp 0 before the barrier
  VA_ROLL
  for (int t = threadIdx.x; t < d.pmax + 1; t += (int)blockDim.x) s.best[t] = ~0ull;
  VA_ROLL
  for (int t = threadIdx.x; t < T; t += (int)blockDim.x) s.oflag[t] = INT_MAX;
  __syncthreads();
  VA_ROLL
  for (int pt = 0; pt < P; ++pt) {           // per point: per-thread minimum, warp minimum, one shared atomic per warp
    const int px = pt ? peaks[2 * (pt - 1)] : d.W / 2, py = pt ? peaks[2 * (pt - 1) + 1] : d.H;
    unsigned long long best = ~0ull;
    VA_ROLL
    for (int k = threadIdx.x; k < R; k += (int)blockDim.x) {
      const int id = s.list_ids[k];
      best = min(best, row_best_key(s.occ + (size_t)id * cw, cw, s.sc[S_C], k, d.cmax, px, py, x0, s.row_y[id], gs));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&s.best[pt], best);
  }
  VA_ROLL
  for (int k = threadIdx.x; k < R; k += (int)blockDim.x) atomicMin(&s.oflag[s.list_ids[k]], k);
  VA_ROLL
  for (int j = threadIdx.x; j < norph; j += (int)blockDim.x) s.oflag[s.orphan_ids[j]] = R + j;
  __syncthreads();
  int* goals = reinterpret_cast<int*>(rec + d.off_goals);
  VA_ROLL
  for (int q = threadIdx.x; q < d.pmax; q += (int)blockDim.x) {
    int gk = 0, gc = 0;
    if (q < npk) {
      const unsigned long long b = s.best[q + 1];
      const unsigned cell = (unsigned)b;
      gk = (b == ~0ull) ? -1 : (int)(cell / d.cmax);
      gc = (b == ~0ull) ? -1 : (int)(cell % d.cmax);
    }
    goals[2 * q] = gk;
    goals[2 * q + 1] = gc;
  }
  int* lookup = reinterpret_cast<int*>(rec + d.off_lookup);
  VA_ROLL
  for (int ly = threadIdx.x; ly < PL; ly += (int)blockDim.x) {
    const int owner = (R > 0) ? s.plane_owner[ly] : -1;
    const int v = (owner >= 0) ? s.oflag[owner] : -1;
    lookup[ly] = (v == INT_MAX) ? -1 : v;
  }
  if (threadIdx.x == 0) {
    const unsigned long long b = s.best[0];
    const unsigned cell = (unsigned)b;
    s.sc[S_START] = (R == 0 || b == ~0ull) ? -1 : (int)(((cell / d.cmax) << 16) | (cell % d.cmax));
  }
}

__device__ void finish_record(const Dims& d, const TailSmem& s, uint8_t* rec) {
  // common tail once list / plane / scalars are in shared memory
  __syncthreads();
  if (s.sc[S_FLAGS] & (VA_FLAG_EMPTY | VA_FLAG_CENTRE_OOB | VA_FLAG_LIST_OOB)) {
    if (threadIdx.x == 0) { s.sc[S_R] = 0; s.sc[S_C] = 0; s.sc[S_NORPH] = 0; s.sc[S_NPEAKS] = 0; s.sc[S_FLAGS] |= VA_FLAG_EMPTY; }
    __syncthreads();
  } else {
    collect_orphans(d, s);
    TT(5);
    easy_segments(d, s);
    __syncthreads();
  }
  TT(6);
  if (d.rmax * d.cwords <= 128) {
    // warp 0: peaks, the end cell of every peak, the path start cell and the lookup rows - concurrently with the other
    // warps' penalty cells; joined by a barrier before the header
    if (!(d.flags & (1 << 29))) {       // default: warp 0 does all the cell searches (VA_TAIL_ROLES=3 splits them; measured slower)
      if (threadIdx.x < 32) {
        find_peaks(d, s, rec);
        goals_for_peaks(d, s, rec);
        start_and_lookup(d, s, rec);
      } else {
        penalties_and_record(d, s, rec, (int)threadIdx.x - 32, (int)blockDim.x - 32);
      }
    } else if (threadIdx.x < 32) {
      find_peaks(d, s, rec);
      TT(7);
      goals_for_peaks(d, s, rec);
      TT(8);
    } else if (threadIdx.x < 64) {
      start_and_lookup(d, s, rec);
    } else {
      penalties_and_record(d, s, rec, (int)threadIdx.x - 64, (int)blockDim.x - 64);
      TT(110);
    }
    __syncthreads();
    if (threadIdx.x == 0) write_header(s, rec);
  } else {
    // large grids: the cell search is shared by the whole block
    if (threadIdx.x < 32) find_peaks(d, s, rec);
    penalties_and_record(d, s, rec, (int)threadIdx.x, (int)blockDim.x);   // measured: better than leaving warp 0 out
    __syncthreads();
    start_goals_lookup_block(d, s, rec);
    __syncthreads();
    if (threadIdx.x == 0) write_header(s, rec);
  }
}

// ---------------------------------------------------------------------------------------------
// mask-driven tail
// ---------------------------------------------------------------------------------------------
// What the contour step needs besides the per-instance reductions (va_contour_core.h).
struct TailContour {
  uint32_t* rowsum;          // [B][max_n][H][nblk] per-(row, 128 px block) summaries (read, then reset to 0)
  const uint8_t* masks;      // [B][max_n][H][W] u8 masks, or nullptr ...
  const uint32_t* bits;      // ... then [B][max_n][H][bit_words] bit-packed masks
  unsigned char* slab;       // [nslab][slab_bytes] global scratch of the general path (word part, big run tables)
  size_t slab_bytes;
  int nslab;
  int* slab_lock;            // [nslab] taken only when the batch has more frames than slabs
  int cap;                   // run capacity of a slab
  int smem_off;              // offset of the general path's shared-memory scratch behind the tail's own
  int smem_bytes;            // its size
};

__device__ const uint16_t g_contour_lut[256] = {VA_CONTOUR_LUT_VALUES};

// Contour step of one instance with the whole CTA (va_contour_core.h): run-based components with hole filling on the
// instance's bounding box, table sums, selection of the kept component; overwrites the instance's lattice samples.
// Per-row scratch and (when they fit) the run table live in shared memory, the bit rows of the few multi-run rows in
// the CTA's global slab.
__device__ void contour_general(const Dims& d, const TailContour& tc, size_t inst, const InstStats& st, unsigned* lattice_inst,
                                unsigned char* smem, unsigned char* slab, const uint16_t* lut, int* sc, unsigned long long* best,
                                cc::InstContour* out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  cc::Work w;
  w.H = d.H; w.W = d.W;
  w.fmt = tc.masks ? 0 : 1;
  w.px = tc.masks ? tc.masks + inst * (size_t)d.H * d.W : nullptr;
  w.bits = tc.masks ? nullptr : tc.bits + inst * (size_t)d.H * d.bit_words;
  w.bit_words = d.bit_words;
  w.rowsum = tc.rowsum + inst * (size_t)d.H * d.nblk; w.nblk = d.nblk;
  w.y0 = st.miny; w.x0w = st.minx >> 5;
  w.R = st.maxy - st.miny + 1; w.Wd = (st.maxx >> 5) - w.x0w + 1;
  w.gs = d.gs; w.lat_rows = d.lat_rows; w.lat_cols = d.lat_cols; w.lat_words = d.lat_words;
  w.cap = tc.cap;
  w.sc = sc; w.best = best;
  w.lattice = lattice_inst;
  w.out = out;
  const cc::RowLayout wl = cc::row_layout(w.R);
  const cc::RowLayout wl_full = cc::row_layout(d.H);
  const cc::GridLayout gl_full = cc::grid_layout(d.H, d.bit_words);
  size_t used = 0;
  const bool rows_in_smem = wl.total <= (size_t)tc.smem_bytes;
  cc::bind_rows(w, rows_in_smem ? smem : slab, wl);
  if (rows_in_smem) used += wl.total;
  cc::bind_grid(w, slab + wl_full.total, gl_full);
  cc::bind_runs(w, slab + wl_full.total + gl_full.total, cc::run_layout(tc.cap));
  __syncthreads();
  // per-phase clocks: compiled in with -DVA_TAIL_PHASE_TIMING only (VA_EXTRA_FLAGS of build.sh) - the array and the
  // printf would otherwise sit in the production kernel's stack frame and instruction footprint
#ifdef VA_TAIL_PHASE_TIMING
  const bool ptime = (d.flags & kTailDebugFlag) && tid == 0;
  long long pt[20]; int npt = 0;
#define PT(k) do { if (ptime) pt[npt++] = clock64(); } while (0)
#else
#define PT(k) do { } while (0)
#endif
  PT(0);
  cc::phase_init(w, tid, nt);       __syncthreads(); PT(1);
  cc::phase_lists(w, tid, nt);      __syncthreads(); PT(2);
  {
    // bit rows of the rows with several runs (usually a handful): shared memory when they fit
    const cc::GridLayout gl = cc::grid_layout(sc[cc::W_NM] > 0 ? sc[cc::W_NM] : 1, w.Wd);
    if (used + gl.total <= (size_t)tc.smem_bytes) {
      cc::bind_grid(w, smem + used, gl);
      used += gl.total;
    }
  }
  cc::phase_load(w, tid, nt);       __syncthreads(); PT(3);
  cc::phase_light_check(w, tid, nt); __syncthreads();
  const bool light = sc[cc::W_LIGHT] != 0;
  if (light) {
    // one component without holes, decided from the run ends: no run table, no union-find, the mask kernel's lattice
    // samples stand
    cc::bind_runs(w, smem + used, cc::run_layout(1));
    cc::phase_light_setup(w, tid, st.minx, st.maxx);  __syncthreads();
    PT(4); PT(5); PT(6); PT(7); PT(8); PT(9); PT(10); PT(11); PT(12); PT(13);
  } else {
    cc::phase_count(w, tid, nt);      __syncthreads(); PT(4);
    {
      int first, excl;
      cc::phase_scan_rows_a(w, tid, nt, first, excl);          __syncthreads(); PT(5);
      cc::phase_scan_rows_b(w, tid, nt, first, excl);          __syncthreads(); PT(6); PT(7);
    }
    {
      const int NR = sc[cc::W_NR];
      const cc::RunLayout rl = cc::run_layout(NR > 0 ? NR : 1);
      if (NR <= tc.cap && used + rl.total <= (size_t)tc.smem_bytes) {
        w.cap = NR > 0 ? NR : 1;
        cc::bind_runs(w, smem + used, rl);
      }
    }
    cc::phase_runs(w, tid, nt);       __syncthreads(); PT(8);
    cc::phase_gaps(w, tid, nt);       __syncthreads(); PT(9);
    cc::phase_holes(w, tid, nt);      __syncthreads(); PT(10);
    cc::phase_link(w, tid, nt);       __syncthreads(); PT(11);
    cc::phase_flatten_a(w, tid, nt);  __syncthreads(); PT(12);
    cc::phase_flatten_b(w, tid, nt);  __syncthreads(); PT(13);
  }
  cc::phase_sums(w, lut, tid, nt);    __syncthreads(); PT(14);
  if (!light) {
    cc::phase_select(w, tid, nt);     __syncthreads(); PT(15);
    cc::phase_bbox(w, tid, nt);       __syncthreads(); PT(16);
  } else {
    PT(15); PT(16);
  }
  cc::phase_output(w, tid, nt);       __syncthreads(); PT(17);
#ifdef VA_TAIL_PHASE_TIMING
  if (ptime) {
    printf("[va tail] frame %d phases (R=%d Wd=%d NR=%d NM=%d): init %lld lists %lld load %lld count %lld scan %lld %lld %lld runs %lld gaps %lld holes %lld link %lld flat %lld %lld sums %lld select %lld bbox %lld out %lld\n",
           (int)blockIdx.x, w.R, w.Wd, sc[cc::W_NR], sc[cc::W_NM], pt[1] - pt[0], pt[2] - pt[1], pt[3] - pt[2], pt[4] - pt[3],
           pt[5] - pt[4], pt[6] - pt[5], pt[7] - pt[6], pt[8] - pt[7], pt[9] - pt[8], pt[10] - pt[9], pt[11] - pt[10],
           pt[12] - pt[11], pt[13] - pt[12], pt[14] - pt[13], pt[15] - pt[14], pt[16] - pt[15], pt[17] - pt[16]);
  }
#endif
#undef PT
}

__global__ void __launch_bounds__(kTailMaxThreads)
tail_kernel(Dims d, const int* __restrict__ counts, InstStats* __restrict__ stats, unsigned* __restrict__ lattice,
            TailContour tc, const int* __restrict__ rects, const int* __restrict__ sel_in,
            uint8_t* __restrict__ records) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TailSmem s;
  tail_smem_layout(d, &s, smem_raw);
  const int b = blockIdx.x;
  uint8_t* rec = records + (size_t)b * d.record_bytes;
  InstStats* st = stats + (size_t)b * d.max_n;
  const int n = min(counts[b], d.max_n);
  const int gs = d.gs, cw = d.cwords;
  const int T = 2 * d.rmax, PL = plane_cap(d);
  TT(0);
  unsigned long long gt_start = 0, gt_wait = 0;
  if ((d.flags & kTailTimelineFlag) && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_start));

  __shared__ unsigned s_area[kMaxInst];
  __shared__ int s_area2[kMaxInst], s_state[kMaxInst];
  __shared__ int s_bbox[kMaxInst][4];
  __shared__ int s_cert[kMaxInst][5];              // certificate partials: ok, pixels, border moves, min x, max x
  __shared__ InstStats s_stats[kMaxInst];
  __shared__ int s_cc[cc::W_COUNT];
  __shared__ unsigned long long s_best;
  __shared__ uint16_t s_lut[256];
  __shared__ cc::InstContour s_out;
  VA_ROLL
  for (int t = threadIdx.x; t < PL; t += (int)blockDim.x) s.plane_owner[t] = -1;
  VA_ROLL
  for (int t = threadIdx.x; t < T * cw; t += (int)blockDim.x) { s.occ[t] = 0; s.art[t] = 0; }
  if (threadIdx.x < kMaxInst) {
    const int i = threadIdx.x;
    s_cert[i][0] = 1; s_cert[i][1] = 0; s_cert[i][2] = 0; s_cert[i][3] = INT_MAX; s_cert[i][4] = -1;
  }
  // Programmatic dependent launch: this kernel may become resident while the mask kernel is still draining;
  // everything above touched only shared memory and kernel inputs.  Wait here for its writes (statistics, lattice
  // bits, row summaries, masks).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if ((d.flags & kTailTimelineFlag) && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_wait));
  TT(1);
  if (d.flags & (1 << 27)) return;      // VA_TAIL_TIMING=3: floor of the stage (launch + dependency wait), records are not written
  if (threadIdx.x < kMaxInst) {      // all per-instance reductions in one parallel round of global loads
    const int i = threadIdx.x;
    InstStats v;
    v.area = 0; v.minx = 0; v.miny = 0; v.maxx = -1; v.maxy = -1;
    if (i < n) v = st[i];
    s_stats[i] = v;
    s_area[i] = v.area;
  }
  __syncthreads();
  TT(16);
  // ---- contour step, part 1: certificate (va_contour_core.h) - every mask row of an instance one run, consecutive
  //      rows touching => one hole-free component whose polygon area follows from the run ends in closed form.
  //      Reads only the per-(row, block) summaries the mask kernel wrote.  Thread groups of >= 32 share the instances. ----
  {
    // Work items = (instance, block of 31 consecutive mask rows), numbered through a prefix sum over the instances and
    // dealt to the warps round-robin: a frame with one tall mask and many small ones keeps every warp busy.  Lane 0 of
    // an item holds the row ABOVE the block (every lane gets its previous row by one shuffle), lanes 1..31 are certified.
    __shared__ int s_task0[kMaxInst + 1];
    if (threadIdx.x < 32) {
      const int i = threadIdx.x;
      const InstStats v = s_stats[i];
      int cnt = (i < n && v.area) ? (v.maxy - v.miny + 31) / 31 : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (i >= o) incl += u; }
      s_task0[i] = incl - cnt;
      if (i == 31) s_task0[32] = incl;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = (int)threadIdx.x >> 5, nwarps = (int)blockDim.x >> 5;
    const int ntask = s_task0[kMaxInst];
    const int my_task0 = s_task0[lane];                                // lane i: first task of instance i
    TT(17);
    // kU work items in flight per warp and all summary words of a chunk loaded before the first is used: the pass is a
    // chain of global round trips, so their number per warp is what counts
    constexpr int kU = 3, kChunk = 8;
    VA_ROLL
    for (int t0 = warp; t0 < ntask; t0 += kU * nwarps) {
      int inst_i[kU], yy[kU];
      bool live[kU];
      cc::RowRun cur[kU];
      const uint32_t* rowp[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = t0 + u * nwarps;
        // instance of task t: the last one whose first task is <= t (empty instances share their successor's start)
        const unsigned le = __ballot_sync(0xffffffffu, lane < n && my_task0 <= t);
        const int i = le ? 31 - __clz(le) : 0;
        inst_i[u] = i;
        const InstStats v = s_stats[i];
        yy[u] = v.miny + 31 * (t - s_task0[i]) + lane - 1;
        live[u] = t < ntask && yy[u] >= v.miny && yy[u] <= v.maxy;
        cur[u].cnt = 0; cur[u].a = 0; cur[u].b = -1;
        rowp[u] = tc.rowsum + (((size_t)b * d.max_n + i) * d.H + (live[u] ? yy[u] : 0)) * d.nblk;
      }
      VA_ROLL
      for (int k0 = 0; k0 < d.nblk; k0 += kChunk) {
        uint32_t raw[kU][kChunk];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) raw[u][j] = (live[u] && k0 + j < d.nblk) ? rowp[u][k0 + j] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) cc::rowsum_accumulate(cur[u], raw[u][j], k0 + j);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (t0 + u * nwarps >= ntask) break;                           // warp-uniform
        const int i = inst_i[u];
        const InstStats v = s_stats[i];
        cc::RowRun prev;
        prev.cnt = __shfl_up_sync(0xffffffffu, cur[u].cnt, 1);
        prev.a = __shfl_up_sync(0xffffffffu, cur[u].a, 1);
        prev.b = __shfl_up_sync(0xffffffffu, cur[u].b, 1);
        int ok = 1, npx = 0, l = 0, minx = INT_MAX, maxx = -1;
        if (live[u] && lane > 0) {
          const cc::CertTerms ct = cc::cert_row(cur[u], prev, yy[u] == v.miny, yy[u] == v.maxy);
          ok = ct.ok; npx = ct.n; l = ct.l; minx = ct.minx; maxx = ct.maxx;
        }
        ok = __all_sync(0xffffffffu, ok);
        npx = (int)__reduce_add_sync(0xffffffffu, (unsigned)npx);
        l = (int)__reduce_add_sync(0xffffffffu, (unsigned)l);
        minx = __reduce_min_sync(0xffffffffu, minx);
        maxx = __reduce_max_sync(0xffffffffu, maxx);
        if (lane == 0) {
          if (!ok) atomicAnd(&s_cert[i][0], 0);
          atomicAdd(&s_cert[i][1], npx); atomicAdd(&s_cert[i][2], l);
          atomicMin(&s_cert[i][3], minx); atomicMax(&s_cert[i][4], maxx);
        }
      }
    }
    __syncthreads();
    TT(18);
    int n_pending = 0;
    {
      // certified instances: put their summaries back to the resting state (all zero).  The warps are split over the
      // instances so that all of them are reset at once (one instance after the other is a chain of short loops).
      const int per = max(1, nwarps / max(n, 1));                     // warps per instance
      const bool vec = ((d.H * d.nblk) & 3) == 0 && (d.H & 3) == 0;   // groups of 4 rows are 16-byte aligned
      VA_ROLL
      for (int i = warp / per; i < n; i += max(1, nwarps / per)) {
        const InstStats v = s_stats[i];
        if (v.area == 0 || !s_cert[i][0]) continue;
        const int sub = warp % per;
        uint32_t* inst_rs = tc.rowsum + ((size_t)b * d.max_n + i) * d.H * d.nblk;
        if (vec) {
          // vector stores; the rows added by rounding to groups of 4 are zero already
          const int y_lo = v.miny & ~3, y_hi = (v.maxy + 4) & ~3;
          uint4* rs4 = reinterpret_cast<uint4*>(inst_rs + (size_t)y_lo * d.nblk);
          VA_ROLL
          for (int t = sub * 32 + lane; t < (y_hi - y_lo) * d.nblk / 4; t += per * 32) rs4[t] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          uint32_t* rs = inst_rs + (size_t)v.miny * d.nblk;
          VA_ROLL
          for (int t = sub * 32 + lane; t < (v.maxy - v.miny + 1) * d.nblk; t += per * 32) rs[t] = 0u;
        }
      }
    }
    TT(19);
    if (threadIdx.x < kMaxInst) {
      const int i = threadIdx.x;
      int state = cc::kEmpty, area2 = 0;
      int bx0 = 0, by0 = 0, bx1 = -1, by1 = -1;
      if (i < n && s_stats[i].area) {
        if (s_cert[i][0]) {
          state = cc::kSimple;
          area2 = 2 * s_cert[i][1] - s_cert[i][2] - 2;
          bx0 = s_cert[i][3]; bx1 = s_cert[i][4]; by0 = s_stats[i].miny; by1 = s_stats[i].maxy;
        } else {
          state = cc::kPending;
        }
      }
      s_state[i] = state; s_area2[i] = area2;
      s_bbox[i][0] = bx0; s_bbox[i][1] = by0; s_bbox[i][2] = bx1; s_bbox[i][3] = by1;
      n_pending = (state == cc::kPending);
    }
    n_pending = __syncthreads_or(n_pending);
    // ---- contour step, part 2: the general path for every instance the certificate did not cover ----
    if (n_pending) {
      VA_ROLL
      for (int t = threadIdx.x; t < 256; t += (int)blockDim.x) s_lut[t] = g_contour_lut[t];
      const int slot = b % tc.nslab;
      if (threadIdx.x == 0 && (int)gridDim.x > tc.nslab) {          // more frames than slabs: frames b and b + nslab share one
        VA_ROLL
        while (atomicCAS(&tc.slab_lock[slot], 0, 1) != 0) __nanosleep(200);
        __threadfence();
      }
      __syncthreads();
      unsigned char* slab = tc.slab + (size_t)slot * tc.slab_bytes;
      VA_ROLL
      for (int i = 0; i < n; ++i) {
        if (s_state[i] != cc::kPending) continue;                   // CTA-uniform
        const size_t inst = (size_t)b * d.max_n + i;
        contour_general(d, tc, inst, s_stats[i], lattice + inst * d.lat_rows * d.lat_words, smem_raw + tc.smem_off, slab, s_lut,
                        s_cc, &s_best, &s_out);
        if (threadIdx.x == 0) {
          const cc::InstContour o = s_out;
          s_state[i] = o.state;
          s_area2[i] = (o.state == cc::kSimple || o.state == cc::kGeneral) ? o.area2 : 0;
          s_bbox[i][0] = o.minx; s_bbox[i][1] = o.miny; s_bbox[i][2] = o.maxx; s_bbox[i][3] = o.maxy;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0 && (int)gridDim.x > tc.nslab) {
        __threadfence();
        atomicExch(&tc.slab_lock[slot], 0);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    VA_ROLL
    for (int q = 0; q < S_COUNT; ++q) s.sc[q] = 0;
    s.sc[S_USE_EASY] = 1;
    // ---- instance selection (FrameProcessor.py:71-73): the polygon with the largest cv2.contourArea, first maximum;
    //      a single instance is taken as it is.  Areas come from the contour step (doubled, exact integers); an empty
    //      polygon has area 0. ----
    int sel = -1;
    if (sel_in) {
      sel = sel_in[b];
      if (sel < 0 || sel >= n) sel = -1;
    } else if (n > 0) {
      sel = 0;
      VA_ROLL
      for (int i = 1; i < n; ++i) if (s_area2[i] > s_area2[sel]) sel = i;
    }
    s.sc[S_SEL] = sel;
    s.sc[S_AREA] = sel >= 0 ? (int)s_area[sel] : 0;
    s.sc[S_EULER] = sel >= 0 ? s_area2[sel] : 0;
    int flags = 0;
    if (sel < 0) {
      flags = VA_FLAG_EMPTY;
    } else if (s_state[sel] == cc::kEmpty) {
      flags = VA_FLAG_EMPTY | VA_FLAG_NO_POLYGON;     // cv2.fillPoly raises on the empty polygon (FrameProcessor.py:86)
    } else if (s_state[sel] != cc::kSimple && s_state[sel] != cc::kGeneral) {
      flags = VA_FLAG_EMPTY | VA_FLAG_OVERFLOW;       // run capacity of the contour step exceeded
    } else {
      if (s_state[sel] == cc::kGeneral) flags |= VA_FLAG_NON_SIMPLE;
      int x, y, w, h;
      if (rects) { x = rects[4 * b]; y = rects[4 * b + 1]; w = rects[4 * b + 2]; h = rects[4 * b + 3]; }
      else { x = s_bbox[sel][0]; y = s_bbox[sel][1]; w = s_bbox[sel][2] - x + 1; h = s_bbox[sel][3] - y + 1; }
      s.sc[S_MINX] = x; s.sc[S_MINY] = y; s.sc[S_MAXX] = x + w - 1; s.sc[S_MAXY] = y + h - 1;
      x -= x % gs;                                   // FrameProcessor.py:79
      y -= y % gs;                                   // :80
      if (w % gs != 0) w += gs - w % gs;             // :81
      if (w > d.W) w = d.W;                          // :82
      if (h % gs != 0) h += gs - h % gs;             // :83
      const int C = ceil_div(w, gs), Rm = ceil_div(h, gs);   // len(arange(x, x+w, gs))
      s.sc[S_X0] = x; s.sc[S_Y0] = y; s.sc[S_C] = C; s.sc[S_RM] = Rm;
      if (C <= 0 || Rm <= 0) flags |= VA_FLAG_EMPTY;
      else if (y + (Rm - 1) * gs + gs / 2 >= d.H || x + (C - 1) * gs + gs / 2 >= d.W) flags |= VA_FLAG_CENTRE_OOB;  // :97
      else if (C > d.cmax || Rm > d.rmax) flags |= VA_FLAG_OVERFLOW | VA_FLAG_EMPTY;
    }
    s.sc[S_FLAGS] = flags;
    TT(2);
  }
  __syncthreads();

  const int sel = s.sc[S_SEL];
  if (!(s.sc[S_FLAGS] & (VA_FLAG_EMPTY | VA_FLAG_CENTRE_OOB))) {
    // ---- centre sampling (FrameProcessor.py:88-97): shift the lattice rows of the selected
    //      instance so that column 0 is x0 ----
    const int C = s.sc[S_C], Rm = s.sc[S_RM];
    const int lx0 = s.sc[S_X0] / gs, ly0 = s.sc[S_Y0] / gs;
    const unsigned* lat = lattice + ((size_t)b * d.max_n + sel) * d.lat_rows * d.lat_words;
    int any = 0;
    VA_ROLL
    for (int t = threadIdx.x; t < Rm * cw; t += (int)blockDim.x) {
      const int r = t / cw, w = t - r * cw;
      const unsigned* lrow = lat + (size_t)(ly0 + r) * d.lat_words;
      const int bitpos = lx0 + 32 * w;
      const int wi = bitpos >> 5, sh = bitpos & 31;
      const unsigned lo = (wi < d.lat_words) ? lrow[wi] : 0u;
      const unsigned hi = (wi + 1 < d.lat_words) ? lrow[wi + 1] : 0u;
      unsigned v = __funnelshift_r(lo, hi, sh);
      const int rem = C - 32 * w;
      if (rem < 32) v &= (rem <= 0) ? 0u : (0xffffffffu >> (32 - rem));
      s.occ[(size_t)r * cw + w] = v;
      any |= (v != 0);
    }
    VA_ROLL
    for (int r = threadIdx.x; r < Rm; r += (int)blockDim.x) {
      s.row_y[r] = s.sc[S_Y0] + r * gs;
      s.row_attr[r] = r;
      s.list_ids[r] = r;
      s.plane_owner[ly0 + r] = r;
    }
    {
      // artificial columns (FrameProcessor.py:60-65) as a bit mask over this frame's columns, the same for every
      // band row: one column per thread, one ballot per 32 columns (erow_last is scratch until easy_segments runs)
      unsigned* am = reinterpret_cast<unsigned*>(s.erow_last);
      const int base = d.W / 2 - 8 * gs;
      VA_ROLL
      for (int c0 = threadIdx.x & ~31; c0 < 32 * cw; c0 += (int)blockDim.x) {
        const int c = c0 + (threadIdx.x & 31);
        const int delta = s.sc[S_X0] + c * gs - base;
        const bool bit = c < C && delta >= 0 && delta % gs == 0 && delta / gs <= 16;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        if ((threadIdx.x & 31) == 0) am[c0 >> 5] = m;
      }
    }
    any = __syncthreads_or(any);
    TT(3);
    if (threadIdx.x == 0) {
      if (!any) {
        s.sc[S_FLAGS] |= VA_FLAG_EMPTY;              // FrameProcessor.py:99-101
      } else {
        // ---- artificial band (FrameProcessor.py:126-165), replayed literally ----
        int list_len = Rm, ncreated = Rm;
        const unsigned* am = reinterpret_cast<const unsigned*>(s.erow_last);   // built by all warps above
        VA_ROLL
        for (int i = d.band_start; i < d.H; i += gs) {
          const int ly = i / gs;
          const int row_idx = floor_div(i - s.sc[S_Y0], gs);
          if (ncreated >= T || ly >= PL) { s.sc[S_FLAGS] |= VA_FLAG_OVERFLOW; break; }
          const int id = ncreated++;
          const int prev = s.plane_owner[ly];
          VA_ROLL
          for (int w = 0; w < cw; ++w) {
            const unsigned pv = (prev >= 0) ? s.occ[(size_t)prev * cw + w] : 0u;
            s.occ[(size_t)id * cw + w] = pv | am[w];
            s.art[(size_t)id * cw + w] = ~pv & am[w];
          }
          s.row_y[id] = i;
          s.row_attr[id] = row_idx;
          s.plane_owner[ly] = id;
          if (row_idx < list_len - 1) {
            int idx = row_idx;
            if (idx < 0) idx += list_len;
            if (idx < 0) { s.sc[S_FLAGS] |= VA_FLAG_LIST_OOB; break; }     // IndexError, :163
            s.list_ids[idx] = id;
          } else {
            if (list_len >= d.rmax) { s.sc[S_FLAGS] |= VA_FLAG_OVERFLOW; break; }
            s.list_ids[list_len++] = id;
          }
        }
        s.sc[S_R] = list_len;
        s.sc[S_NCREATED] = ncreated;
      }
    }
  }
  __syncthreads();
  TT(4);
  finish_record(d, s, rec);
  TT(13);
  TT(114);

  // ---- reset the reduction scratch for the next call ----
  __syncthreads();
  TT(9);
  TT(115);
  VA_ROLL
  for (int i = threadIdx.x; i < d.max_n; i += (int)blockDim.x) {
    InstStats z;
    z.area = 0; z.minx = INT_MAX; z.miny = INT_MAX; z.maxx = -1; z.maxy = -1; z.euler4 = 0; z.pad0 = 0; z.pad1 = 0;
    st[i] = z;
  }
  TT(11);
  unsigned* latb = lattice + (size_t)b * d.max_n * d.lat_rows * d.lat_words;
  VA_ROLL
  for (int t = threadIdx.x; t < d.max_n * d.lat_rows * d.lat_words; t += (int)blockDim.x) latb[t] = 0u;
  TT(12);
  if ((d.flags & kTailTimelineFlag) && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    printf("[va tail] block %d on sm %d: starts %llu waited %llu ends %llu ns\n", (int)blockIdx.x, (int)__smid(), gt_start, gt_wait, gt);
  }
  if ((d.flags & kTailDebugFlag) && (int)blockIdx.x == ((d.flags >> 8) & 0xfff) && threadIdx.x == 0) {
    const long long t0 = g_tail_t[0], te = clock64();
    printf("[va tail] reset: stats %lld lattice %lld rest %lld | t0 before barrier %lld t32 before %lld t32 after %lld (from t6)\n", g_tail_t[11] - g_tail_t[9], g_tail_t[12] - g_tail_t[11], te - g_tail_t[12],
           g_tail_t[13] - g_tail_t[6], g_tail_t[14] - g_tail_t[6], g_tail_t[15] - g_tail_t[6]);
    printf("[va tail] certificate: stats %lld tasks %lld rows %lld (%d tasks) zero %lld rest %lld\n", g_tail_t[16] - g_tail_t[1], g_tail_t[17] - g_tail_t[16],
           g_tail_t[18] - g_tail_t[17], 0, g_tail_t[19] - g_tail_t[18], g_tail_t[2] - g_tail_t[19]);
    printf("[va tail] block %d ", (int)blockIdx.x);
    printf("[va tail] cycles: wait %lld select %lld sample %lld band %lld | orphans %lld easy %lld | warp0: peaks %lld cells %lld"
           " | warp1 penalties %lld | end barrier %lld total %lld\n",
           g_tail_t[1] - t0, g_tail_t[2] - g_tail_t[1], g_tail_t[3] - g_tail_t[2], g_tail_t[4] - g_tail_t[3],
           g_tail_t[5] - g_tail_t[4], g_tail_t[6] - g_tail_t[5], g_tail_t[7] - g_tail_t[6], g_tail_t[8] - g_tail_t[7],
           g_tail_t[10] - g_tail_t[6], g_tail_t[9] - g_tail_t[6], te - t0);
  }
}

// ---------------------------------------------------------------------------------------------
// grid mode: list rows (+ optional lookup rows) given by the caller
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTailMaxThreads)
grid_mode_kernel(Dims d, const va_grid_input* __restrict__ hdr, const int* __restrict__ row_y,
                 const int* __restrict__ row_attr, const uint8_t* __restrict__ occ, const int* __restrict__ plane_y,
                 const uint8_t* __restrict__ plane_occ, uint8_t* __restrict__ records) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TailSmem s;
  tail_smem_layout(d, &s, smem_raw);
  const int b = blockIdx.x;
  uint8_t* rec = records + (size_t)b * d.record_bytes;
  const va_grid_input h = hdr[b];
  const int cw = d.cwords, T = 2 * d.rmax, PL = plane_cap(d), gs = d.gs;
  const int R = min(max(h.n_rows, 0), d.rmax), C = min(max(h.n_cols, 0), d.cmax);
  const int NP = (plane_y && plane_occ) ? min(max(h.n_plane, 0), d.rmax) : 0;

  VA_ROLL
  for (int t = threadIdx.x; t < PL; t += (int)blockDim.x) s.plane_owner[t] = -1;
  VA_ROLL
  for (int t = threadIdx.x; t < T * cw; t += (int)blockDim.x) { s.occ[t] = 0; s.art[t] = 0; }
  if (threadIdx.x == 0) {
    VA_ROLL
    for (int q = 0; q < S_COUNT; ++q) s.sc[q] = 0;
    s.sc[S_USE_EASY] = h.use_easy;
    s.sc[S_SEL] = -1;
    s.sc[S_X0] = h.x0; s.sc[S_C] = C; s.sc[S_R] = R; s.sc[S_RM] = R;
    s.sc[S_NCREATED] = R + NP;
    int flags = (R == 0 || C == 0) ? VA_FLAG_EMPTY : 0;
    if (h.n_rows > d.rmax || h.n_cols > d.cmax || h.n_plane > d.rmax) flags |= VA_FLAG_OVERFLOW | VA_FLAG_EMPTY;
    s.sc[S_FLAGS] = flags;
  }
  __syncthreads();
  // bit-pack rows: one thread per (row, word)
  VA_ROLL
  for (int t = threadIdx.x; t < (R + NP) * cw; t += (int)blockDim.x) {
    const int id = t / cw, w = t - id * cw;
    const uint8_t* src = (id < R) ? occ + ((size_t)b * d.rmax + id) * d.cmax
                                  : plane_occ + ((size_t)b * d.rmax + (id - R)) * d.cmax;
    unsigned vo = 0, va_ = 0;
    VA_ROLL
    for (int q = 0; q < 32; ++q) {
      const int c = 32 * w + q;
      if (c >= C) break;
      const uint8_t v = src[c];
      vo |= (unsigned)(v & 1u) << q;
      va_ |= (unsigned)((v >> 1) & 1u) << q;
    }
    s.occ[t] = vo;
    s.art[t] = va_;
  }
  VA_ROLL
  for (int id = threadIdx.x; id < R + NP; id += (int)blockDim.x) {
    s.row_y[id] = (id < R) ? row_y[(size_t)b * d.rmax + id] : plane_y[(size_t)b * d.rmax + (id - R)];
    s.row_attr[id] = (id < R) ? row_attr[(size_t)b * d.rmax + id] : -1;
    if (id < R) s.list_ids[id] = id;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // grid_lookup rows: explicit plane rows, else the list rows in order (later rows override)
    const int lo = NP ? R : 0, hi = NP ? R + NP : R;
    VA_ROLL
    for (int id = lo; id < hi; ++id) {
      const int y = s.row_y[id];
      if (y < 0 || y % gs != 0 || y / gs >= PL) { s.sc[S_FLAGS] |= VA_FLAG_OVERFLOW | VA_FLAG_EMPTY; break; }
      s.plane_owner[y / gs] = id;
    }
    VA_ROLL
    for (int id = 0; id < R && NP; ++id) {   // every list row must have a lookup row
      const int y = s.row_y[id];
      if (y < 0 || y % gs != 0 || y / gs >= PL || s.plane_owner[y / gs] < 0) { s.sc[S_FLAGS] |= VA_FLAG_OVERFLOW | VA_FLAG_EMPTY; break; }
    }
  }
  finish_record(d, s, rec);
}

// ---------------------------------------------------------------------------------------------
static int tail_threads(const Dims& d, int B) {
  if (const char* e = getenv("VA_TAIL_THREADS")) { const int v = atoi(e); if (v == 128 || v == 256 || v == 512 || v == 1024) return v; }   // tuning aid
  // measured: 512 threads beat 256 also at 640^2 / gs = 20 (31 -> 23 us per 256 frames).  1024 are slower when every SM
  // holds a frame, faster when most SMs would idle (cfg2: 32 frames of 6048 cells, 45 -> 35 us; one frame: 21.5 -> 20.5 us)
  const int sms = d.num_sms > 0 ? d.num_sms : 148;
  return (2 * B <= sms) ? kTailMaxThreads : 512;
}

size_t contour_slab_bytes(const Dims& d, int cap) {
  return cc::row_layout(d.H).total + cc::grid_layout(d.H, d.bit_words).total + cc::run_layout(cap).total + 256;
}

cudaError_t launch_tail(const Dims& d, const int* counts, int B, const Scratch& sc, const uint8_t* masks, const int* rects,
                        const int* sel, uint8_t* records, cudaStream_t st) {
  // shared memory: the tail's own tables, then the contour step's per-row scratch (+ the run table when it fits)
  const size_t own = (tail_smem_bytes(d) + 15) & ~(size_t)15;
  const size_t cc_smem = cc::row_layout(d.H).total + kContourGridSmem + cc::run_layout(2048).total;
  const size_t smem = own + cc_smem;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  TailContour tc;
  tc.rowsum = sc.rowsum; tc.masks = masks; tc.bits = sc.bits;
  tc.slab = sc.cc_slab; tc.slab_bytes = sc.cc_slab_bytes; tc.nslab = sc.nslab; tc.slab_lock = sc.slab_lock;
  tc.cap = sc.cc_cap; tc.smem_off = (int)own; tc.smem_bytes = (int)cc_smem;
  // launched with programmatic stream serialization: see griddepcontrol.wait in the kernel
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B);
  cfg.blockDim = dim3(tail_threads(d, B));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("VA_NO_PDL") != nullptr;   // tuning aid: plain stream-ordered launch
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, tail_kernel, d, counts, sc.stats, sc.lattice, tc, rects, sel, records);
}

cudaError_t launch_grid_mode(const Dims& d, const va_grid_input* hdr, const int* row_y, const int* row_attr,
                             const uint8_t* occ, const int* plane_y, const uint8_t* plane_occ, int B,
                             uint8_t* records, cudaStream_t st) {
  const size_t smem = tail_smem_bytes(d);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(grid_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  grid_mode_kernel<<<B, tail_threads(d, B), smem, st>>>(d, hdr, row_y, row_attr, occ, plane_y, plane_occ, records);
  return cudaGetLastError();
}

}  // namespace va
