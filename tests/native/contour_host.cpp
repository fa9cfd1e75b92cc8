// Host build of the contour algorithm (vision_assist_b200/csrc/va_contour_core.h) for the CPU tests: the same
// phase functions the CUDA kernel runs, driven by a loop over emulated thread ids.  TEST INFRASTRUCTURE ONLY - the
// product path is the CUDA tail kernel (va_tail.cu, which runs the same phases); nothing under vision_assist_b200/ links this file.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../vision_assist_b200/csrc/va_contour_core.h"

using namespace va::cc;

// Emulates what the mask kernels leave behind for one instance: per-(row, 128 px block) summaries and the pixel bbox.
static void summarise(const uint8_t* m, int H, int W, std::vector<uint32_t>& rowsum, int nblk, int bbox[4]) {
  bbox[0] = 1 << 30; bbox[1] = 1 << 30; bbox[2] = -1; bbox[3] = -1;
  for (int y = 0; y < H; ++y)
    for (int k = 0; k < nblk; ++k) {
      int cnt = 0, first = 0, last = 0;
      for (int x = k * kRowBlock; x < (k + 1) * kRowBlock && x < W; ++x)
        if (m[(size_t)y * W + x]) {
          if (!cnt) first = x - k * kRowBlock;
          last = x - k * kRowBlock;
          ++cnt;
          if (x < bbox[0]) bbox[0] = x;
          if (x > bbox[2]) bbox[2] = x;
          if (y < bbox[1]) bbox[1] = y;
          bbox[3] = y;
        }
      rowsum[(size_t)y * nblk + k] = rowsum_pack(cnt, first, last);
    }
}

// out: [0] state, [1] area2, [2..5] minx miny maxx maxy, [6] points, [7] n_components, [8] path (0 certificate, 1 general)
extern "C" int contour_host_instance(const uint8_t* mask, int H, int W, int gs, int fmt, int force_general, int nthreads,
                                     int cap, int* out, uint32_t* lattice) {
  const int nblk = (W + kRowBlock - 1) / kRowBlock;
  std::vector<uint32_t> rowsum((size_t)H * nblk);
  int bbox[4];
  summarise(mask, H, W, rowsum, nblk, bbox);
  const int half = gs / 2;
  const int lat_rows = (H - half + gs - 1) / gs, lat_cols = (W - half + gs - 1) / gs, lat_words = (lat_cols + 31) / 32;
  // lattice as the mask kernels sample it (raw mask)
  for (int i = 0; i < lat_rows * lat_words; ++i) lattice[i] = 0;
  for (int ly = 0; ly < lat_rows; ++ly)
    for (int lx = 0; lx < lat_cols; ++lx)
      if (mask[(size_t)(gs * ly + half) * W + gs * lx + half]) lattice[ly * lat_words + (lx >> 5)] |= 1u << (lx & 31);
  InstContour res;
  memset(&res, 0, sizeof(res));
  res.maxx = -1; res.maxy = -1;
  out[8] = 0;
  if (bbox[2] < 0) {
    res.state = kEmpty;
  } else {
    // certificate
    int ok = 1, n = 0, l = 0, minx = 1 << 30, maxx = -1;
    RowRun prev; prev.cnt = 0; prev.a = 0; prev.b = -1;
    for (int y = bbox[1]; y <= bbox[3]; ++y) {
      const RowRun cur = rowsum_combine(&rowsum[(size_t)y * nblk], nblk);
      const CertTerms t = cert_row(cur, prev, y == bbox[1], y == bbox[3]);
      ok &= t.ok; n += t.n; l += t.l;
      if (t.minx < minx) minx = t.minx;
      if (t.maxx > maxx) maxx = t.maxx;
      prev = cur;
    }
    if (ok && !force_general) {
      res.state = kSimple; res.area2 = 2 * n - l - 2;
      res.minx = minx; res.maxx = maxx; res.miny = bbox[1]; res.maxy = bbox[3]; res.n_components = 1;
    } else {
      out[8] = 1;
      Work w;
      memset(&w, 0, sizeof(w));
      w.H = H; w.W = W; w.fmt = fmt; w.gs = gs; w.lat_rows = lat_rows; w.lat_cols = lat_cols; w.lat_words = lat_words;
      std::vector<uint32_t> bits;
      const int bw = (W + 31) / 32;
      if (fmt == 1) {
        bits.assign((size_t)H * bw, 0u);
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x)
            if (mask[(size_t)y * W + x]) bits[(size_t)y * bw + (x >> 5)] |= 1u << (x & 31);
        w.bits = bits.data(); w.bit_words = bw;
      } else {
        w.px = mask;
      }
      w.rowsum = rowsum.data(); w.nblk = nblk;
      w.y0 = bbox[1]; w.x0w = bbox[0] >> 5; w.R = bbox[3] - bbox[1] + 1; w.Wd = (bbox[2] >> 5) - w.x0w + 1;
      w.cap = cap;
      const RowLayout wl = row_layout(w.R);
      const GridLayout gl = grid_layout(w.R, w.Wd);
      const RunLayout rl = run_layout(cap);
      std::vector<unsigned char> scratch0(wl.total + 64, 0xcd), scratch(gl.total + 64, 0xcd), scratch2(rl.total + 64, 0xcd);   // poison: the phases must initialise what they read
      bind_rows(w, scratch0.data(), wl);
      bind_grid(w, scratch.data(), gl);
      bind_runs(w, scratch2.data(), rl);
      int sc[W_COUNT];
      unsigned long long best = 0;
      w.sc = sc; w.best = &best; w.lattice = lattice; w.out = &res;
      const int nt = nthreads;
#define PHASE(call) for (int tid = 0; tid < nt; ++tid) { call; }
      PHASE(phase_init(w, tid, nt));
      PHASE(phase_lists(w, tid, nt));
      PHASE(phase_load(w, tid, nt));
      PHASE(phase_light_check(w, tid, nt));
      if (sc[W_LIGHT] && force_general != 2) {          // force_general == 2: the full path even where the light one applies
        out[8] = 2;
        PHASE(phase_light_setup(w, tid, bbox[0], bbox[2]));
        PHASE(phase_sums(w, kContourLutHost, tid, nt));
        PHASE(phase_output(w, tid, nt));
      } else {
        sc[W_LIGHT] = 0;
        PHASE(phase_count(w, tid, nt));
        PHASE(phase_scan_a(w, tid, nt));
        PHASE(phase_scan_b(w, tid, nt));
        PHASE(phase_scan_c(w, tid, nt));
        PHASE(phase_runs(w, tid, nt));
        PHASE(phase_gaps(w, tid, nt));
        PHASE(phase_holes(w, tid, nt));
        PHASE(phase_link(w, tid, nt));
        PHASE(phase_flatten_a(w, tid, nt));
        PHASE(phase_flatten_b(w, tid, nt));
        PHASE(phase_sums(w, kContourLutHost, tid, nt));
        PHASE(phase_select(w, tid, nt));
        PHASE(phase_bbox(w, tid, nt));
        PHASE(phase_output(w, tid, nt));
      }
#undef PHASE
    }
  }
  out[0] = res.state; out[1] = res.area2; out[2] = res.minx; out[3] = res.miny; out[4] = res.maxx; out[5] = res.maxy;
  out[6] = res.points; out[7] = res.n_components;
  return 0;
}
