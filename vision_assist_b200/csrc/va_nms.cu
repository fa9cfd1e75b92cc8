// Front of the path: candidate filter + NMS on the raw segmentation-head output, one CTA per image.
//
// Follows `non_max_suppression` of the vendored ultralytics ops (testing/old/segmenting_using_tflite/ops.py:214-363)
// in the configuration FrameProcessor uses through model.predict(frame, conf=0.5) (FrameProcessor.py:322) - best
// class only, no class filter, not rotated - and the torchvision CPU NMS kernel it calls:
//   candidates = anchors with best class confidence > conf_thres, in anchor order (:277, :307-308); xywh -> xyxy in
//   fp32 (:472-480); boxes offset by class * max_wh unless agnostic (:319-325); greedy suppression in order of
//   descending score, stable; j is suppressed by a kept i iff inter / (area_i + area_j - inter) > iou_thres in fp32;
//   the first max_det survivors (:327).
//   more than max_nms candidates: only the max_nms best scores are considered (:332-333).
// Output is what ops.process_mask / va_run_fused consume: boxes (xyxy, input pixels), mask coefficients, counts.
// Latency-bound integer / compare work on a few KB per image: no tensor cores, everything in shared memory.
#include "va_common.cuh"

namespace va {

constexpr int kNmsThreads = 256;
constexpr int kNmsCap = 512;             // candidates per tile (tiles are visited in order of descending score)
constexpr int kNmsWords = kNmsCap / 32;
constexpr int kNmsKeep = kMaxInst;       // survivors the mask path can take (max_det <= max_n <= kMaxInst)

struct NmsSmem {
  unsigned long long key[kNmsCap];       // (descending score, ascending candidate index) sort keys
  float box[kNmsCap][4];                 // class-offset xyxy, candidate order
  float area[kNmsCap];
  float score[kNmsCap];
  int anchor[kNmsCap];
  int cls[kNmsCap];
  unsigned mask[kNmsCap][kNmsWords];     // sorted position p: later positions q with IoU(p, q) > threshold
  unsigned pre[kNmsWords];               // sorted positions suppressed by a survivor of an earlier tile
  int kept[kNmsKeep];                    // sorted positions kept in this tile
  float kbox[kNmsKeep][4];               // survivors so far (class-offset boxes), in keep order
  float karea[kNmsKeep];
  float kscore[kNmsKeep];
  int kanchor[kNmsKeep];
  int kcls[kNmsKeep];
  int warp_tot[kNmsThreads / 32];
  unsigned hist[256];                    // radix select of the best (tile + 1) * kNmsCap scores
  unsigned sel_prefix, sel_need;
  int nkept, nkept_before;
};

size_t nms_smem_bytes() { return sizeof(NmsSmem); }

__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ pred, int A, int nc, int nm, float conf_thres, float iou_thres, float class_offset,
           int max_det, int max_n, int max_nms, float* __restrict__ coefs_out, float* __restrict__ boxes_out,
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
  // sum over the CTA of one int per thread, and the exclusive prefix over the warps below this one
  auto cta_offsets = [&](int warp_count, int& before, int& all) {
    if (lane == 0) s.warp_tot[warp] = warp_count;
    __syncthreads();
    before = 0; all = 0;
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      if (w < warp) before += s.warp_tot[w];
      all += s.warp_tot[w];
    }
    __syncthreads();
  };
  // ---- 0. how many anchors pass the confidence filter ----
  int total = 0;
  {
    int mine = 0;
    for (int a = tid; a < A; a += kNmsThreads) { int j; mine += (anchor_conf(a, j) > conf_thres) ? 1 : 0; }
    mine = __reduce_add_sync(0xffffffffu, mine);
    int before;
    cta_offsets(mine, before, total);
  }
  if (tid == 0) s.nkept = 0;
  __syncthreads();
  // The greedy NMS visits the candidates in order of (score descending, anchor ascending) and suppression only flows
  // from better to worse ones.  So the candidates are taken in tiles of kNmsCap in that order: a tile is first
  // thinned by the survivors of the earlier tiles, then scanned like the first one, until max_det have survived or
  // the candidates (the best max_nms of them, ops.py:332-333) run out.  One tile covers the usual case.
  const int limit = min(total, max_nms);
  unsigned prev_thr = 0xffffffffu;       // tile boundary above this tile: keys > prev_thr, and prev_eq of the ties on
  int prev_eq = 0;                       // prev_thr (in anchor order), belong to earlier tiles
  for (int lo = 0; lo < limit; lo += kNmsCap) {
    const int hi = min(limit, lo + kNmsCap);
    // ---- 1a. boundary below this tile: the hi-th best key (4-pass radix select on the orderable score bits) ----
    unsigned thr_key = 0;                // candidates with key > thr_key, and n_eq of the ties on it, rank below hi
    int n_eq = INT_MAX;
    if (hi < total) {
      unsigned prefix = 0;
      int need = hi;                                     // still to be found among keys with the current prefix
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
    // ---- 1b. ordered compaction of the tile's anchors (anchor order is the reference's tie-break) ----
    const bool select = lo > 0 || hi < total;
    int base = 0, eq_base = 0, peq_base = 0;
    for (int a0 = 0; a0 < A; a0 += kNmsThreads) {
      const int a = a0 + tid;
      float conf = -INFINITY;
      int j = 0;
      if (a < A) conf = anchor_conf(a, j);
      bool flag = (a < A) && (conf > conf_thres);
      if (select) {
        const unsigned key = flag ? orderable(conf) : 0u;
        // rank of this anchor among the ties on either boundary key, in anchor order
        const bool eq = flag && key == thr_key, peq = flag && key == prev_thr;
        const unsigned eqb = __ballot_sync(0xffffffffu, eq), peqb = __ballot_sync(0xffffffffu, peq);
        int eoff, etot, poff, ptot;
        cta_offsets(__popc(eqb), eoff, etot);
        cta_offsets(__popc(peqb), poff, ptot);
        const int erank = eq_base + eoff + __popc(eqb & ((1u << lane) - 1u));
        const int prank = peq_base + poff + __popc(peqb & ((1u << lane) - 1u));
        const bool above_lower = key > thr_key || (eq && erank < n_eq);
        const bool below_upper = key < prev_thr || (peq && prank >= prev_eq);
        flag = flag && above_lower && below_upper;
        eq_base += etot;
        peq_base += ptot;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      int off, tot;
      cta_offsets(__popc(bal), off, tot);
      const int pos = base + off + __popc(bal & ((1u << lane) - 1u));
      if (flag && pos < kNmsCap) {
        s.anchor[pos] = a;
        s.score[pos] = conf;
        s.cls[pos] = j;
      }
      base += tot;
    }
    const int n = min(base, kNmsCap);    // == hi - lo
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
    if (tid < kNmsWords) s.pre[tid] = 0u;
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
    // ---- 3. suppression: by the survivors of earlier tiles, then the matrix over the tile's sorted positions and
    //         the sequential greedy scan by warp 0 ----
    auto overlaps = [&](float ix1, float iy1, float ix2, float iy2, float ia, int j) -> bool {
      const float xx1 = fmaxf(ix1, s.box[j][0]), yy1 = fmaxf(iy1, s.box[j][1]);
      const float xx2 = fminf(ix2, s.box[j][2]), yy2 = fminf(iy2, s.box[j][3]);
      const float ww = fmaxf(0.f, __fsub_rn(xx2, xx1)), hh = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(ww, hh);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, s.area[j]), inter));
      return ovr > iou_thres;
    };
    const int nk0 = s.nkept;
    for (int p = tid; p < n && nk0 > 0; p += kNmsThreads) {
      const int j = (int)(unsigned)s.key[p];
      bool hit = false;
      for (int k = 0; k < nk0 && !hit; ++k) hit = overlaps(s.kbox[k][0], s.kbox[k][1], s.kbox[k][2], s.kbox[k][3], s.karea[k], j);
      if (hit) atomicOr(&s.pre[p >> 5], 1u << (p & 31));
    }
    const int nw = (n + 31) >> 5;
    for (int t = tid; t < n * nw; t += kNmsThreads) {
      const int p = t / nw, w = t - p * nw;
      const int i = (int)(unsigned)s.key[p];
      const float ix1 = s.box[i][0], iy1 = s.box[i][1], ix2 = s.box[i][2], iy2 = s.box[i][3], ia = s.area[i];
      unsigned m = 0;
      for (int q = max(32 * w, p + 1); q < min(32 * w + 32, n); ++q) {
        if (overlaps(ix1, iy1, ix2, iy2, ia, (int)(unsigned)s.key[q])) m |= 1u << (q & 31);
      }
      s.mask[p][w] = m;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned remv = (lane < kNmsWords) ? s.pre[lane] : 0u;   // lane w holds word w of the suppressed set
      int k = nk0;
      for (int p = 0; p < n && k < max_det; ++p) {
        const unsigned word = __shfl_sync(0xffffffffu, remv, p >> 5);
        if (!((word >> (p & 31)) & 1u)) {
          if (lane == 0) s.kept[k - nk0] = p;
          ++k;
          if (lane < nw) remv |= s.mask[p][lane];
        }
      }
      if (lane == 0) { s.nkept_before = nk0; s.nkept = k; }
    }
    __syncthreads();
    // the tile's survivors join the list
    for (int k = s.nkept_before + tid; k < s.nkept; k += kNmsThreads) {
      const int i = (int)(unsigned)s.key[s.kept[k - s.nkept_before]];
      s.kbox[k][0] = s.box[i][0]; s.kbox[k][1] = s.box[i][1]; s.kbox[k][2] = s.box[i][2]; s.kbox[k][3] = s.box[i][3];
      s.karea[k] = s.area[i];
      s.kscore[k] = s.score[i];
      s.kanchor[k] = s.anchor[i];
      s.kcls[k] = s.cls[i];
    }
    __syncthreads();
    if (s.nkept >= max_det) break;
    prev_thr = thr_key;
    prev_eq = n_eq;
  }
  // ---- 4. rows of the survivors in keep order: what process_mask / va_run_fused take ----
  const int k = min(s.nkept, min(max_det, max_n));
  if (tid == 0) counts_out[b] = k;
  for (int t = tid; t < max_n * 4; t += kNmsThreads) {
    const int slot = t >> 2, c = t & 3;
    float v = 0.f;
    if (slot < k) {
      const int a = s.kanchor[slot];
      // un-offset box, recomputed exactly as the reference's rows hold it (xy -+ wh / 2)
      const float ctr = __ldg(P + (size_t)(c & 1) * A + a), half = __fmul_rn(__ldg(P + (size_t)(2 + (c & 1)) * A + a), 0.5f);
      v = (c < 2) ? __fsub_rn(ctr, half) : __fadd_rn(ctr, half);
    }
    boxes_out[((size_t)b * max_n + slot) * 4 + c] = v;
  }
  for (int slot = tid; slot < max_n; slot += kNmsThreads) {
    const bool live = slot < k;
    if (conf_out) conf_out[(size_t)b * max_n + slot] = live ? s.kscore[slot] : 0.f;
    if (cls_out) cls_out[(size_t)b * max_n + slot] = live ? s.kcls[slot] : 0;
  }
  for (int t = tid; t < max_n * nm; t += kNmsThreads) {
    const int slot = t / nm, m = t - slot * nm;
    float v = 0.f;
    if (slot < k) v = __ldg(P + (size_t)(4 + nc + m) * A + s.kanchor[slot]);
    coefs_out[((size_t)b * max_n + slot) * nm + m] = v;
  }
}

// ops.scale_boxes (ops.py:139-174, padding = True, xyxy) + clip_boxes (:367-385) on the kept boxes: from the letterboxed
// model input back to the original frame.  gain and pad are computed on the host exactly as Python does (doubles,
// round-half-even); the tensor arithmetic is fp32: subtract the pad, divide by fp32(gain), clamp.
__global__ void scale_boxes_kernel(const float* __restrict__ boxes, const int* __restrict__ counts, int max_n, int B, float pad_x,
                                   float pad_y, float gain, float w0, float h0, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * max_n * 4) return;
  const int c = t & 3, slot = (t >> 2) % max_n, b = (t >> 2) / max_n;
  float v = 0.f;
  if (slot < counts[b]) {
    v = __fdiv_rn(__fsub_rn(boxes[t], (c & 1) ? pad_y : pad_x), gain);
    const float hi = (c & 1) ? h0 : w0;
    v = fminf(fmaxf(v, 0.f), hi);
  }
  out[t] = v;
}

cudaError_t launch_scale_boxes(const float* boxes, const int* counts, int max_n, int B, float pad_x, float pad_y, float gain,
                               float w0, float h0, float* out, cudaStream_t st) {
  const int total = B * max_n * 4;
  if (total == 0) return cudaSuccess;
  scale_boxes_kernel<<<(total + 255) / 256, 256, 0, st>>>(boxes, counts, max_n, B, pad_x, pad_y, gain, w0, h0, out);
  return cudaGetLastError();
}

cudaError_t launch_nms(const float* pred, int A, int nc, int nm, float conf_thres, float iou_thres, float class_offset,
                       int max_det, int max_n, int max_nms, int B, float* coefs_out, float* boxes_out, float* conf_out, int* cls_out,
                       int* counts_out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem));
  if (e != cudaSuccess) return e;
  nms_kernel<<<B, kNmsThreads, sizeof(NmsSmem), st>>>(pred, A, nc, nm, conf_thres, iou_thres, class_offset, max_det, max_n,
                                                      max_nms, coefs_out, boxes_out, conf_out, cls_out, counts_out);
  return cudaGetLastError();
}

}  // namespace va
