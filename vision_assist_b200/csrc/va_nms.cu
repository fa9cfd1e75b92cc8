// Front of the path: candidate filter + NMS on the raw segmentation-head output, one CTA per image.
//
// Follows `non_max_suppression` of the vendored ultralytics ops (testing/old/segmenting_using_tflite/ops.py:214-363)
// in the configuration FrameProcessor uses through model.predict(frame, conf=0.5) (FrameProcessor.py:322) - best
// class only, no class filter, not rotated - and the torchvision CPU NMS kernel it calls:
//   candidates = anchors with best class confidence > conf_thres, in anchor order (:277, :307-308); xywh -> xyxy in
//   fp32 (:472-480); boxes offset by class * max_wh unless agnostic (:319-325); greedy suppression in order of
//   descending score, stable; j is suppressed by a kept i iff inter / (area_i + area_j - inter) > iou_thres in fp32;
//   the first max_det survivors (:327).
// Output is what ops.process_mask / va_run_fused consume: boxes (xyxy, input pixels), mask coefficients, counts.
// Latency-bound integer / compare work on a few KB per image: no tensor cores, everything in shared memory.
#include "va_common.cuh"

namespace va {

constexpr int kNmsThreads = 256;
constexpr int kNmsCap = 512;             // candidates per image that survive the confidence filter
constexpr int kNmsWords = kNmsCap / 32;

struct NmsSmem {
  unsigned long long key[kNmsCap];       // (descending score, ascending candidate index) sort keys
  float box[kNmsCap][4];                 // class-offset xyxy, candidate order
  float area[kNmsCap];
  float score[kNmsCap];
  int anchor[kNmsCap];
  int cls[kNmsCap];
  unsigned mask[kNmsCap][kNmsWords];     // sorted position p: later positions q with IoU(p, q) > threshold
  int kept[kNmsCap];
  int warp_tot[kNmsThreads / 32];
  unsigned hist[256];                    // radix select of the kNmsCap best scores when more anchors pass the filter
  unsigned sel_prefix, sel_need;
  int n, nkept, n_eq;
};

size_t nms_smem_bytes() { return sizeof(NmsSmem); }

__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ pred, int A, int nc, int nm, float conf_thres, float iou_thres, float class_offset,
           int max_det, int max_n, float* __restrict__ coefs_out, float* __restrict__ boxes_out,
           float* __restrict__ conf_out, int* __restrict__ cls_out, int* __restrict__ counts_out) {
  extern __shared__ __align__(16) unsigned char nms_raw[];
  NmsSmem& s = *reinterpret_cast<NmsSmem*>(nms_raw);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = pred + (size_t)b * (4 + nc + nm) * A;

  // best class confidence of an anchor (first maximum; NaN is sticky like torch.max and fails the filter)
  auto anchor_conf = [&](int a, int& j) -> float {
    float conf = -INFINITY;
    j = 0;
    for (int c = 0; c < nc; ++c) {
      const float v = __ldg(P + (size_t)(4 + c) * A + a);
      if (c == 0 || v > conf || v != v) { conf = v; j = c; }
    }
    return conf;
  };
  auto orderable = [](float f) -> unsigned {            // ascending in the float order
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
  };
  // ---- 0. how many anchors pass the confidence filter ----
  int total = 0;
  {
    int mine = 0;
    for (int a = tid; a < A; a += kNmsThreads) { int j; mine += (anchor_conf(a, j) > conf_thres) ? 1 : 0; }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) s.warp_tot[warp] = mine;
    __syncthreads();
    for (int w = 0; w < kNmsThreads / 32; ++w) total += s.warp_tot[w];
    __syncthreads();
  }
  // More candidates than the kernel holds: keep the kNmsCap best by (score descending, anchor ascending) - the
  // order in which the greedy NMS visits them.  Suppression only flows from better to worse candidates, so the
  // survivors among the best kNmsCap are exactly the reference's first survivors; the result is complete when
  // max_det of them survive (checked below).  4-pass radix select on the orderable score bits.
  unsigned thr_key = 0;     // candidates with key > thr_key are in; n_eq of those with key == thr_key (anchor order)
  int n_eq = 0;
  const bool overflow = total > kNmsCap;
  if (overflow) {
    unsigned prefix = 0;
    int need = kNmsCap;                                  // still to be found among keys with the current prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int t = tid; t < 256; t += kNmsThreads) s.hist[t] = 0;
      __syncthreads();
      const unsigned himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
      for (int a = tid; a < A; a += kNmsThreads) {
        int j;
        const float conf = anchor_conf(a, j);
        if (!(conf > conf_thres)) continue;
        const unsigned key = orderable(conf);
        if ((key & himask) == (prefix & himask)) atomicAdd(&s.hist[(key >> shift) & 0xffu], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int acc = 0, d = 255;
        for (; d > 0; --d) {                            // largest digit first
          if (acc + (int)s.hist[d] >= need) break;
          acc += (int)s.hist[d];
        }
        s.sel_prefix = prefix | ((unsigned)d << shift);
        s.sel_need = (unsigned)(need - acc);
      }
      __syncthreads();
      prefix = s.sel_prefix;
      need = (int)s.sel_need;
      __syncthreads();
    }
    thr_key = prefix;
    n_eq = need;
  }
  // ---- 1. ordered compaction of the selected anchors (anchor order is the reference's tie-break) ----
  int base = 0, eq_base = 0;
  for (int a0 = 0; a0 < A; a0 += kNmsThreads) {
    const int a = a0 + tid;
    float conf = -INFINITY;
    int j = 0;
    if (a < A) conf = anchor_conf(a, j);
    bool flag = (a < A) && (conf > conf_thres);
    if (overflow) {
      const unsigned key = flag ? orderable(conf) : 0u;
      const bool eq = flag && key == thr_key;
      // rank of this anchor among the ties on the threshold key, in anchor order
      const unsigned eqb = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s.warp_tot[warp] = __popc(eqb);
      __syncthreads();
      int eoff = eq_base, etot = 0;
      for (int w = 0; w < kNmsThreads / 32; ++w) {
        if (w < warp) eoff += s.warp_tot[w];
        etot += s.warp_tot[w];
      }
      __syncthreads();
      const int erank = eoff + __popc(eqb & ((1u << lane) - 1u));
      flag = flag && (key > thr_key || (eq && erank < n_eq));
      eq_base += etot;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s.warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      if (w < warp) off += s.warp_tot[w];
      tot += s.warp_tot[w];
    }
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    if (flag && pos < kNmsCap) {
      s.anchor[pos] = a;
      s.score[pos] = conf;
      s.cls[pos] = j;
    }
    base += tot;
    __syncthreads();
  }
  const int n = base;
  // ---- 2. boxes (xywh -> xyxy, class offset), areas, sort keys ----
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int t = tid; t < n2; t += kNmsThreads) {
    if (t < n) {
      const int a = s.anchor[t];
      const float x = __ldg(P + a), y = __ldg(P + (size_t)A + a), w = __ldg(P + 2 * (size_t)A + a), h = __ldg(P + 3 * (size_t)A + a);
      const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
      const float c = __fmul_rn((float)s.cls[t], class_offset);
      const float x1 = __fadd_rn(__fsub_rn(x, hw), c), y1 = __fadd_rn(__fsub_rn(y, hh), c);
      const float x2 = __fadd_rn(__fadd_rn(x, hw), c), y2 = __fadd_rn(__fadd_rn(y, hh), c);
      s.box[t][0] = x1; s.box[t][1] = y1; s.box[t][2] = x2; s.box[t][3] = y2;
      s.area[t] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      s.key[t] = ((unsigned long long)(~orderable(s.score[t])) << 32) | (unsigned)t;   // ascending key = descending score
    } else {
      s.key[t] = ~0ull;
    }
  }
  __syncthreads();
  // bitonic sort (ascending key = descending score, ties by candidate index: torch's stable descending sort)
  for (int k = 2; k <= n2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int t = tid; t < n2; t += kNmsThreads) {
        const int q = t ^ jj;
        if (q > t) {
          const unsigned long long x = s.key[t], y = s.key[q];
          const bool up = (t & k) == 0;
          if ((x > y) == up) { s.key[t] = y; s.key[q] = x; }
        }
      }
      __syncthreads();
    }
  }
  // ---- 3. suppression matrix over sorted positions, then the sequential greedy scan by warp 0 ----
  const int nw = (n + 31) >> 5;
  for (int t = tid; t < n * nw; t += kNmsThreads) {
    const int p = t / nw, w = t - p * nw;
    const int i = (int)(unsigned)s.key[p];
    const float ix1 = s.box[i][0], iy1 = s.box[i][1], ix2 = s.box[i][2], iy2 = s.box[i][3], ia = s.area[i];
    unsigned m = 0;
    for (int q = max(32 * w, p + 1); q < min(32 * w + 32, n); ++q) {
      const int j = (int)(unsigned)s.key[q];
      const float xx1 = fmaxf(ix1, s.box[j][0]), yy1 = fmaxf(iy1, s.box[j][1]);
      const float xx2 = fminf(ix2, s.box[j][2]), yy2 = fminf(iy2, s.box[j][3]);
      const float ww = fmaxf(0.f, __fsub_rn(xx2, xx1)), hh = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(ww, hh);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, s.area[j]), inter));
      if (ovr > iou_thres) m |= 1u << (q & 31);
    }
    s.mask[p][w] = m;
  }
  __syncthreads();
  if (warp == 0) {
    unsigned remv = 0;                                   // lane w holds word w of the suppressed set
    int k = 0;
    for (int p = 0; p < n && k < max_det; ++p) {
      const unsigned word = __shfl_sync(0xffffffffu, remv, p >> 5);
      if (!((word >> (p & 31)) & 1u)) {
        if (lane == 0) s.kept[k] = p;
        ++k;
        if (lane < nw) remv |= s.mask[p][lane];
      }
    }
    if (lane == 0) s.nkept = k;
  }
  __syncthreads();
  // ---- 4. rows of the survivors in keep order: what process_mask / va_run_fused take ----
  const int k = min(s.nkept, min(max_det, max_n));
  // candidates were dropped and fewer than max_det survived: the dropped ones could have survived too - report
  if (tid == 0) counts_out[b] = (overflow && s.nkept < max_det) ? -total : k;
  for (int t = tid; t < max_n * 4; t += kNmsThreads) {
    const int slot = t >> 2, c = t & 3;
    float v = 0.f;
    if (slot < k) {
      const int i = (int)(unsigned)s.key[s.kept[slot]];
      const int a = s.anchor[i];
      // un-offset box, recomputed exactly as the reference's rows hold it (xy -+ wh / 2)
      const float ctr = __ldg(P + (size_t)(c & 1) * A + a), half = __fmul_rn(__ldg(P + (size_t)(2 + (c & 1)) * A + a), 0.5f);
      v = (c < 2) ? __fsub_rn(ctr, half) : __fadd_rn(ctr, half);
    }
    boxes_out[((size_t)b * max_n + slot) * 4 + c] = v;
  }
  for (int slot = tid; slot < max_n; slot += kNmsThreads) {
    const bool live = slot < k;
    const int i = live ? (int)(unsigned)s.key[s.kept[slot]] : 0;
    if (conf_out) conf_out[(size_t)b * max_n + slot] = live ? s.score[i] : 0.f;
    if (cls_out) cls_out[(size_t)b * max_n + slot] = live ? s.cls[i] : 0;
  }
  for (int t = tid; t < max_n * nm; t += kNmsThreads) {
    const int slot = t / nm, m = t - slot * nm;
    float v = 0.f;
    if (slot < k) {
      const int i = (int)(unsigned)s.key[s.kept[slot]];
      v = __ldg(P + (size_t)(4 + nc + m) * A + s.anchor[i]);
    }
    coefs_out[((size_t)b * max_n + slot) * nm + m] = v;
  }
}

cudaError_t launch_nms(const float* pred, int A, int nc, int nm, float conf_thres, float iou_thres, float class_offset,
                       int max_det, int max_n, int B, float* coefs_out, float* boxes_out, float* conf_out, int* cls_out,
                       int* counts_out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem));
  if (e != cudaSuccess) return e;
  nms_kernel<<<B, kNmsThreads, sizeof(NmsSmem), st>>>(pred, A, nc, nm, conf_thres, iou_thres, class_offset, max_det, max_n,
                                                      coefs_out, boxes_out, conf_out, cls_out, counts_out);
  return cudaGetLastError();
}

}  // namespace va
