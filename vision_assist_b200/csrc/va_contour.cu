// Contour step between the mask kernels and the frame tail: per (frame, instance) the polygon the reference keeps.
//
// Replaces, on the GPU and without tracing a contour:
//   masks2segments   vendored ultralytics ops.py:837-859   cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), contour
//                                                          with the most points
//   cv2.contourArea  FrameProcessor.py:72-73               (the selection key; the tail kernel takes the first maximum)
//   cv2.boundingRect / cv2.fillPoly   FrameProcessor.py:75-86   bbox and raster of the kept contour
// The algorithm and its proof obligations are in va_contour_core.h (shared with the host build the CPU tests run).
//
// Two kernels:
//   contour_certify_kernel   one warp per instance; reads only the per-(row, 128 px block) summaries the mask kernels
//                            wrote (4 B per block, a few KB per instance).  Every row one run + consecutive rows touch
//                            -> one hole-free component: area in closed form, the lattice samples of the mask kernel
//                            are already those of the filled polygon.  Anything else goes on a work list.
//   contour_general_kernel   persistent CTAs drain the work list: run-based connected components with hole filling on
//                            the bit image of the instance's bounding box (shared memory when it fits, an L2-resident
//                            slab otherwise), table sums, selection; overwrites the instance's lattice samples with
//                            those of the kept component.  Exits at once when the list is empty.
// Both are programmatic dependents of the kernel before them (griddepcontrol.wait before the first dependent read).
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "va_common.cuh"
#include "va_contour_core.h"

namespace va {

using cc::InstContour;

constexpr int kCertThreads = 128;
constexpr int kGenThreads = 512;

__device__ uint16_t g_contour_lut[256];

// One CTA per (frame, instance): thread t takes mask rows miny + t, miny + t + 128, ...; a row needs its own summary
// and the previous row's (a neighbour lane's, one extra load at the lane-0 seam).
__global__ void __launch_bounds__(kCertThreads)
contour_certify_kernel(Dims d, const int* __restrict__ counts, int B, const InstStats* __restrict__ stats,
                       uint32_t* __restrict__ rowsum, InstContour* __restrict__ out, int* __restrict__ worklist) {
  __shared__ int s_red[5];     // ok, n, l, minx, maxx
  const int inst = blockIdx.x;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = inst / d.max_n, i = inst - b * d.max_n;
  InstContour o;
  o.area2 = 0; o.state = cc::kEmpty; o.minx = 0; o.miny = 0; o.maxx = -1; o.maxy = -1; o.points = 0; o.n_components = 0;
  const InstStats st = stats[inst];
  if (i >= min(counts[b], d.max_n) || st.area == 0) {
    if (threadIdx.x == 0) out[inst] = o;
    return;
  }
  if (threadIdx.x == 0) { s_red[0] = 1; s_red[1] = 0; s_red[2] = 0; s_red[3] = INT_MAX; s_red[4] = -1; }
  __syncthreads();
  uint32_t* rs = rowsum + (size_t)inst * d.H * d.nblk;
  int ok = 1, n = 0, l = 0, minx = INT_MAX, maxx = -1;
  for (int y0 = st.miny; y0 <= st.maxy; y0 += kCertThreads) {
    const int y = y0 + (int)threadIdx.x;
    const bool live = y <= st.maxy;
    cc::RowRun cur; cur.cnt = 0; cur.a = 0; cur.b = -1;
    if (live) cur = cc::rowsum_combine(rs + (size_t)y * d.nblk, d.nblk);
    cc::RowRun prev;
    prev.cnt = __shfl_up_sync(0xffffffffu, cur.cnt, 1);
    prev.a = __shfl_up_sync(0xffffffffu, cur.a, 1);
    prev.b = __shfl_up_sync(0xffffffffu, cur.b, 1);
    if (live && lane == 0 && y > st.miny) prev = cc::rowsum_combine(rs + (size_t)(y - 1) * d.nblk, d.nblk);
    if (live) {
      const cc::CertTerms t = cc::cert_row(cur, prev, y == st.miny, y == st.maxy);
      ok &= t.ok; n += t.n; l += t.l;
      minx = min(minx, t.minx); maxx = max(maxx, t.maxx);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  n = (int)__reduce_add_sync(0xffffffffu, (unsigned)n);
  l = (int)__reduce_add_sync(0xffffffffu, (unsigned)l);
  minx = __reduce_min_sync(0xffffffffu, minx);
  maxx = __reduce_max_sync(0xffffffffu, maxx);
  if (lane == 0) {
    if (!ok) atomicAnd(&s_red[0], 0);
    atomicAdd(&s_red[1], n); atomicAdd(&s_red[2], l);
    atomicMin(&s_red[3], minx); atomicMax(&s_red[4], maxx);
  }
  __syncthreads();
  if (s_red[0]) {
    // certified: nobody else reads this instance's summaries - put them back to their resting state (pending
    // instances are reset by the general path after it has read them)
    for (int t = threadIdx.x; t < (st.maxy - st.miny + 1) * d.nblk; t += kCertThreads) rs[(size_t)st.miny * d.nblk + t] = 0u;
  }
  if (threadIdx.x == 0) {
    if (s_red[0]) {
      o.state = cc::kSimple;
      o.area2 = 2 * s_red[1] - s_red[2] - 2;
      o.minx = s_red[3]; o.maxx = s_red[4]; o.miny = st.miny; o.maxy = st.maxy;
      o.n_components = 1;
    } else {
      o.state = cc::kPending;
      o.minx = st.minx; o.maxx = st.maxx; o.miny = st.miny; o.maxy = st.maxy;
      const int slot = atomicAdd(&worklist[0], 1);
      worklist[2 + slot] = inst;
    }
    out[inst] = o;
  }
}

struct GenParams {
  Dims d;
  const uint8_t* masks;      // [B][max_n][H][W] or nullptr
  const uint32_t* bits;      // [B][max_n][H][bit_words] when masks == nullptr
  uint32_t* rowsum;          // [B][max_n][H][nblk]
  const InstStats* stats;
  unsigned* lattice;
  InstContour* out;
  int* worklist;
  unsigned char* slab;       // [gridDim.x][slab_bytes]
  size_t slab_bytes;
  int cap;                   // run capacity (global slab)
  int smem_bytes;            // dynamic shared memory available for the scratch parts
  int timing;                // developer diagnostic (VA_CC_TIMING=1): CTA 0 prints the cycles of every phase of its first item
};

__global__ void __launch_bounds__(kGenThreads)
contour_general_kernel(const GenParams p) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ int s_sc[cc::W_COUNT];
  __shared__ unsigned long long s_best;
  __shared__ uint16_t s_lut[256];
  __shared__ unsigned long long s_dbg[4];
  const Dims& d = p.d;
  const int tid = threadIdx.x, nt = blockDim.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int n_items = p.worklist[0];
  if (n_items > 0)
    for (int t = tid; t < 256; t += nt) s_lut[t] = g_contour_lut[t];
  unsigned char* slab = p.slab + (size_t)blockIdx.x * p.slab_bytes;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int inst = p.worklist[2 + item];
    const InstStats st = p.stats[inst];
    cc::Work w;
    w.H = d.H; w.W = d.W;
    w.fmt = p.masks ? 0 : 1;
    w.px = p.masks ? p.masks + (size_t)inst * d.H * d.W : nullptr;
    w.bits = p.masks ? nullptr : p.bits + (size_t)inst * d.H * d.bit_words;
    w.bit_words = d.bit_words;
    w.rowsum = p.rowsum + (size_t)inst * d.H * d.nblk; w.nblk = d.nblk;
    w.y0 = st.miny; w.x0w = st.minx >> 5;
    w.R = st.maxy - st.miny + 1; w.Wd = (st.maxx >> 5) - w.x0w + 1;
    w.gs = d.gs; w.lat_rows = d.lat_rows; w.lat_cols = d.lat_cols; w.lat_words = d.lat_words;
    w.cap = p.cap;
    w.sc = s_sc; w.best = &s_best;
    w.lattice = p.lattice + (size_t)inst * d.lat_rows * d.lat_words;
    w.out = p.out + inst;
    w.dbg = (p.timing && blockIdx.x == 0 && item == blockIdx.x) ? s_dbg : nullptr;
    if (w.dbg && tid < 4) s_dbg[tid] = 0ull;
    // row part, word part, run part: each in shared memory when it fits behind the previous one, else in the CTA's
    // slab (L2-resident); the run part is placed after the scan, when the number of runs is known
    const cc::RowLayout wl = cc::row_layout(w.R);
    const cc::GridLayout gl = cc::grid_layout(w.R, w.Wd);
    const cc::RowLayout wl_full = cc::row_layout(d.H);
    const cc::GridLayout gl_full = cc::grid_layout(d.H, d.bit_words);
    size_t used = 0;
    const bool rows_in_smem = wl.total <= (size_t)p.smem_bytes;
    cc::bind_rows(w, rows_in_smem ? smem_dyn : slab, wl);
    if (rows_in_smem) used += wl.total;
    const bool grid_in_smem = used + gl.total <= (size_t)p.smem_bytes;
    cc::bind_grid(w, grid_in_smem ? smem_dyn + used : slab + wl_full.total, gl);
    if (grid_in_smem) used += gl.total;
    unsigned char* run_slab = slab + wl_full.total + gl_full.total;
    cc::bind_runs(w, run_slab, cc::run_layout(p.cap));
    long long tstamp[20];
    int nst = 0;
#define VA_TS() do { if (p.timing && blockIdx.x == 0 && tid == 0 && item == 0) tstamp[nst++] = clock64(); } while (0)
    __syncthreads();                                       // previous item done with the shared scalars
    VA_TS();
    cc::phase_init(w, tid, nt);       __syncthreads();
    cc::phase_lists(w, tid, nt);      __syncthreads();
    cc::phase_load(w, tid, nt);       __syncthreads(); VA_TS();
    cc::phase_count(w, tid, nt);      __syncthreads(); VA_TS();
    cc::phase_scan_a(w, tid, nt);     __syncthreads();
    cc::phase_scan_b(w, tid, nt);     __syncthreads();
    cc::phase_scan_c(w, tid, nt);     __syncthreads(); VA_TS();
    {
      const int NR = s_sc[cc::W_NR];
      const cc::RunLayout rl = cc::run_layout(NR > 0 ? NR : 1);
      if (NR <= p.cap && used + rl.total <= (size_t)p.smem_bytes) {
        w.cap = NR > 0 ? NR : 1;
        cc::bind_runs(w, smem_dyn + used, rl);
      }
    }
    cc::phase_runs(w, tid, nt);       __syncthreads(); VA_TS();
    cc::phase_gaps(w, tid, nt);       __syncthreads(); VA_TS();
    cc::phase_holes(w, tid, nt);      __syncthreads(); VA_TS();
    cc::phase_link(w, tid, nt);       __syncthreads(); VA_TS();
    cc::phase_flatten_a(w, tid, nt);  __syncthreads();
    cc::phase_flatten_b(w, tid, nt);  __syncthreads(); VA_TS();
    cc::phase_sums(w, s_lut, tid, nt); __syncthreads(); VA_TS();
    cc::phase_select(w, tid, nt);     __syncthreads();
    cc::phase_bbox(w, tid, nt);       __syncthreads(); VA_TS();
    cc::phase_output(w, tid, nt);
    if (p.timing && blockIdx.x == 0 && item == 0) {
      __syncthreads();
      VA_TS();
      if (tid == 0) {
        printf("[va contour] sums sections (max cycles over threads): rows %llu words %llu flush %llu\n", s_dbg[0], s_dbg[1], s_dbg[2]);
        printf("[va contour] items %d inst %d R %d Wd %d runs %d holes %d roots %d grid_smem %d | cycles: load %lld count %lld scan %lld runs %lld "
               "gaps %lld holes %lld link %lld flatten %lld sums %lld select+bbox %lld output %lld | total %lld\n",
               n_items, inst, w.R, w.Wd, s_sc[cc::W_NR], s_sc[cc::W_HOLES], s_sc[cc::W_ROOTS], (int)grid_in_smem,
               tstamp[1] - tstamp[0], tstamp[2] - tstamp[1], tstamp[3] - tstamp[2], tstamp[4] - tstamp[3], tstamp[5] - tstamp[4],
               tstamp[6] - tstamp[5], tstamp[7] - tstamp[6], tstamp[8] - tstamp[7], tstamp[9] - tstamp[8], tstamp[10] - tstamp[9],
               tstamp[11] - tstamp[10], tstamp[11] - tstamp[0]);
      }
    }
#undef VA_TS
  }
  // the last CTA to finish re-arms the work list for the next call (every CTA has read the count by now)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&p.worklist[1], 1) == (int)gridDim.x - 1) {
      p.worklist[0] = 0;
      p.worklist[1] = 0;
    }
  }
}

size_t contour_slab_bytes(const Dims& d, int cap) {
  return cc::row_layout(d.H).total + cc::grid_layout(d.H, d.bit_words).total + cc::run_layout(cap).total + 256;
}

static bool g_lut_loaded[64] = {false};

cudaError_t launch_contour(const Dims& d, const int* counts, int B, const Scratch& sc, const uint8_t* masks, cudaStream_t st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && !g_lut_loaded[dev]) {
    e = cudaMemcpyToSymbol(g_contour_lut, kContourLutHost, sizeof(kContourLutHost));
    if (e != cudaSuccess) return e;
    g_lut_loaded[dev] = true;
  }
  static const bool no_pdl = getenv("VA_NO_PDL") != nullptr;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * d.max_n);
    cfg.blockDim = dim3(kCertThreads);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    e = cudaLaunchKernelEx(&cfg, contour_certify_kernel, d, counts, B, (const InstStats*)sc.stats, sc.rowsum,
                           reinterpret_cast<InstContour*>(sc.contour), sc.worklist);
    if (e != cudaSuccess) return e;
  }
  GenParams p;
  p.d = d; p.masks = masks; p.bits = sc.bits; p.rowsum = sc.rowsum; p.stats = sc.stats; p.lattice = sc.lattice;
  p.out = reinterpret_cast<InstContour*>(sc.contour); p.worklist = sc.worklist;
  p.slab = sc.cc_slab; p.slab_bytes = sc.cc_slab_bytes; p.cap = sc.cc_cap;
  static int smem_opt = -1;
  if (smem_opt < 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) v = 48 * 1024;
    smem_opt = v - 4096;                                    // static shared memory of the kernel + margin
    if (smem_opt > 160 * 1024) smem_opt = 160 * 1024;       // leave room for a resident tail CTA
    e = cudaFuncSetAttribute(contour_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_opt);
    if (e != cudaSuccess) return e;
  }
  p.smem_bytes = smem_opt;
  static const bool cc_timing = getenv("VA_CC_TIMING") != nullptr;
  p.timing = cc_timing ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sc.cc_ctas);
  cfg.blockDim = dim3(kGenThreads);
  cfg.dynamicSmemBytes = (size_t)smem_opt;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, contour_general_kernel, p);
}

}  // namespace va
