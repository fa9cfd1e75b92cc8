// Bilinear upsample (align_corners=False) of cropped proto-resolution logits to frame resolution,
// threshold > 0, u8 mask store, and the per-instance reductions the grid stage needs
// (pixel area, pixel bbox, cell-centre lattice samples).
//
// Follows ops.process_mask lines :736-737 (F.interpolate bilinear + gt_(0.0)); the 4-tap blend is
// implemented with exactly the roundings torch's CPU kernel performs (oracle/mask_assembly.py
// `bilinear_upsample_np`):   top = fma(a, l0x, fl(b*l1x)) ; out = fma(top, l0y, fl(bot*l1y))
// and the source index  src = max(fma(scale, dst+0.5, -0.5), 0).
//
// Two kernels: an exact-4x specialisation (H = 4*mh, W = 4*mw: every YOLOv8-seg head at its native
// input size) whose weights are the constants {.125,.375,.625,.875}, and a generic-scale kernel.
// Both are HBM-store bound: n*H*W bytes out per frame, logits re-read from L2.
#include "va_up_common.cuh"

namespace va {

struct RowState {
  float s[6];
  float h[16];
  float mn, mx;
  bool hvalid;
};

// Common sink of one finished dst row of a lane: lattice sample, bit-packed row (grid-only mode) and the
// (row, 128 px block) summary.  Called by ALL lanes of a lane group at the same point (warp reductions inside).
struct RowSink {
  const Dims& d;
  uint32_t* rowsum_inst;
  uint16_t* bits_inst;      // nullptr unless grid-only
  unsigned* lat;
  int g, gl;
  bool gok;                 // this lane's 16-pixel piece exists (g < NG)
  LeaderStats ls;
  __device__ __forceinline__ void row(unsigned pat, int Y, bool exists) {
    const bool ex = exists && gok;
    if (ex && pat) {
      const int t = Y - (d.gs >> 1);
      if (t >= 0 && t % d.gs == 0) lattice_row_pat(pat, Y, 16 * g, d.gs, d.lat_cols, d.lat_words, lat);
    }
    if (ex && bits_inst) bits_inst[(size_t)Y * (2 * d.bit_words) + g] = (uint16_t)pat;
    // Most rows of most blocks are all zeros or all ones: two votes find the 8-lane groups for which that holds and
    // their summaries are constants; only a warp with a block that crosses the outline gathers the patterns.
    // (A lane past the row end, or a row that does not exist, counts as both - its leader then writes nothing.)
    const unsigned p = ex ? pat : 0u;
    const unsigned zm = __ballot_sync(0xffffffffu, p == 0u), fm = __ballot_sync(0xffffffffu, p == 0xffffu || !ex);
    const unsigned gz = zm & (zm >> 4), gf = fm & (fm >> 4);
    const unsigned gz2 = gz & (gz >> 2), gf2 = gf & (gf >> 2);
    const unsigned gzall = gz2 & (gz2 >> 1) & 0x01010101u, gfall = gf2 & (gf2 >> 1) & 0x01010101u;
    if ((gzall | gfall) == 0x01010101u) {
      if (gl == 0 && ex) {             // the leader's piece exists whenever any piece of its group does
        const bool fullg = ((gfall >> ((threadIdx.x & 31) & 24)) & 1u) && !((gzall >> ((threadIdx.x & 31) & 24)) & 1u);
        const int blk = g >> 3;
        // a "full" group whose last pieces lie past the row end is a ragged block: count its real pixels
        const int npx = min(cc::kRowBlock, d.W - cc::kRowBlock * blk);
        if (fullg) {
          rowsum_inst[(size_t)Y * d.nblk + blk] = cc::rowsum_pack(npx, 0, npx - 1);
          ls.area += (unsigned)npx;
          ls.minx = min(ls.minx, cc::kRowBlock * blk);
          ls.maxx = max(ls.maxx, cc::kRowBlock * blk + npx - 1);
          ls.miny = min(ls.miny, Y);
          ls.maxy = max(ls.maxy, Y);
        }
      }
      return;
    }
    emit_row_summary(p, gl, exists, Y, g >> 3, rowsum_inst, d.nblk, ls);
  }
};

template <bool kWriteMasks>
__global__ void __launch_bounds__(kUpThreads, 3)
upsample4x_kernel(Dims d, const float* __restrict__ logits, const int* __restrict__ counts, uint8_t* __restrict__ masks,
                  MaskSinks sk) {
  const int b = blockIdx.z, i = blockIdx.y;
  if (i >= min(counts[b], d.max_n)) return;
  const int NG = d.W >> 4;                        // 16-px groups per row
  const int NG8 = ceil_div(NG, 8);
  const int NS4 = ceil_div(d.mh, 4 * kStrip);
  const int wt = blockIdx.x * (kUpThreads / 32) + (threadIdx.x >> 5);
  if (wt >= NG8 * NS4) return;
  const int lane = threadIdx.x & 31;
  const int g = (wt % NG8) * 8 + (lane & 7);
  const int strip = (wt / NG8) * 4 + (lane >> 3);
  const int r_begin = strip * kStrip;
  const int r_end = min(r_begin + kStrip, d.mh);
  if (r_begin >= d.mh) return;                    // the whole lane group leaves together
  const bool gok = g < NG;
  const int ge = gok ? g : NG - 1;                // lanes past the row width follow the group with a clamped piece

  const size_t inst = (size_t)b * d.max_n + i;
  const float* L = logits + inst * d.mh * d.mw;
  uint8_t* M = kWriteMasks ? masks + inst * (size_t)d.H * d.W + 16 * ge : nullptr;
  RowSink sink{d, sk.rowsum + inst * (size_t)d.H * d.nblk,
               (!kWriteMasks && sk.bits) ? reinterpret_cast<uint16_t*>(sk.bits) + inst * (size_t)d.H * (2 * d.bit_words) : nullptr,
               sk.lattice + inst * (size_t)d.lat_rows * d.lat_words, g, lane & 7, gok, LeaderStats()};

  const bool left = (ge == 0);
  RowState A, Bq;
  load6(L + (size_t)r_begin * d.mw, ge, d.mw, A.s, A.mn, A.mx);
  A.hvalid = false;
  auto store = [&](const uint4& w, int Y) {
    if (kWriteMasks && gok) *reinterpret_cast<uint4*>(M + (size_t)Y * d.W) = w;
  };
  const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  const uint4 zeros = make_uint4(0, 0, 0, 0);

  if (r_begin == 0) {   // dst rows 0,1: src y clamps to 0 -> l0 = 1, l1 = 0 -> value = h(row 0)
    uint4 w;
    if (A.mn > kTiny) w = ones;
    else if (A.mx <= 0.f) w = zeros;
    else {
      hinterp4(A.s, A.h, left);
      A.hvalid = true;
      w = hpack(A.h);
    }
    store(w, 0);
    store(w, 1);
    const unsigned pat = pat16(w);
    sink.row(pat, 0, true);
    sink.row(pat, 1, true);
  }

  auto step = [&](RowState& P, RowState& Q, int r) {
    // pair (r, r+1): dst rows 4r+2 .. 4r+5
    load6(L + (size_t)(r + 1) * d.mw, ge, d.mw, Q.s, Q.mn, Q.mx);
    Q.hvalid = false;
    const float mn = fminf(P.mn, Q.mn), mx = fmaxf(P.mx, Q.mx);
    const int Y0 = 4 * r + 2;
    unsigned pp[4];
    if (mn > kTiny) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { store(ones, Y0 + j); pp[j] = 0xffffu; }
    } else if (mx <= 0.f) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { store(zeros, Y0 + j); pp[j] = 0u; }
    } else {
      if (!P.hvalid) { hinterp4(P.s, P.h, left); P.hvalid = true; }
      hinterp4(Q.s, Q.h, left);
      Q.hvalid = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float l1 = 0.125f + 0.25f * (float)j;
        const uint4 w = vblend(P.h, Q.h, 1.0f - l1, l1);
        store(w, Y0 + j);
        pp[j] = pat16(w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sink.row(pp[j], Y0 + j, true);
  };
  auto last_step = [&](RowState& P) {
    // bottom edge r = mh-1: src row r+1 clamps to r, only dst rows 4r+2, 4r+3 exist
    const int Y0 = 4 * (d.mh - 1) + 2;
    unsigned pp[2];
    if (P.mn > kTiny) { store(ones, Y0); store(ones, Y0 + 1); pp[0] = pp[1] = 0xffffu; }
    else if (P.mx <= 0.f) { store(zeros, Y0); store(zeros, Y0 + 1); pp[0] = pp[1] = 0u; }
    else {
      if (!P.hvalid) { hinterp4(P.s, P.h, left); P.hvalid = true; }
      const uint4 w0 = vblend(P.h, P.h, 0.875f, 0.125f), w1 = vblend(P.h, P.h, 0.625f, 0.375f);
      store(w0, Y0); store(w1, Y0 + 1);
      pp[0] = pat16(w0); pp[1] = pat16(w1);
    }
    sink.row(pp[0], Y0, true);
    sink.row(pp[1], Y0 + 1, true);
  };
  const int r_pairs_end = min(r_end, d.mh - 1);   // pairs with a real second row
  int r = r_begin;
  for (; r + 1 < r_pairs_end; r += 2) {
    step(A, Bq, r);
    step(Bq, A, r + 1);
  }
  if (r < r_pairs_end) {
    step(A, Bq, r);
    if (r_end == d.mh) last_step(Bq);
  } else if (r_end == d.mh) {
    last_step(A);
  }
  if ((lane & 7) == 0) publish_leader(sink.ls, sk.stats + inst);
}

// ---------------------------------------------------------------------------------------------
// generic-scale kernel: one thread = 16 consecutive dst pixels of kGenRows dst rows
// ---------------------------------------------------------------------------------------------
constexpr int kGenRows = 8;   // dst rows per thread (same 16 columns): the column span and its sign range are reused

template <bool kWriteMasks>
__global__ void __launch_bounds__(kUpThreads, 8)
upsample_generic_kernel(Dims d, const float* __restrict__ logits, const float* __restrict__ boxes,
                        const int* __restrict__ counts, uint8_t* __restrict__ masks, MaskSinks sk) {
  const int b = blockIdx.z, i = blockIdx.y;
  if (i >= min(counts[b], d.max_n)) return;
  const size_t inst = (size_t)b * d.max_n + i;
  // warp tile: 8 column groups (128 px) x 4 row blocks of kGenRows rows; a store instruction writes four 128 B row pieces
  const int NG = ceil_div(d.W, 16);
  const int NG8 = ceil_div(NG, 8);
  const int NYB = ceil_div(d.H, 4 * kGenRows);
  const int wt = blockIdx.x * (kUpThreads / 32) + (threadIdx.x >> 5);
  if (wt >= NG8 * NYB) return;
  const int lane = threadIdx.x & 31;
  // Conservative dst-space rectangle outside of which no source tap can lie inside the instance's box
  // (crop_mask, ops.py:688-704, zeroed everything else): threads outside it only store zeros.  Lanes 0..3 compute one
  // bound each (no shared memory, no block barrier: the warps of a CTA are independent).
  int rXa = 0, rYa = 0, rXb = d.W - 1, rYb = d.H - 1;
  if (boxes) {
    const int c = lane & 3;                                       // 0: x1, 1: y1, 2: x2, 3: y2
    const float scaled = __fmul_rn(__ldg(boxes + inst * 4 + c), (c & 1) ? d.hr : d.wr);
    // kept proto cols [ceil(bx1), ceil(bx2)-1]; a dst pixel X reads cols x0(X), x0(X)+1 with x0 = floor(sx*(X+.5)-.5):
    // X touches cols [cA,cB] iff (cA-.5)/sx-.5 <= X < (cB+1.5)/sx-.5 ; widened by >= 1 on both sides
    const float e = ceilf(fminf(fmaxf(scaled, -4.f), 1e6f)) - ((c & 2) ? 1.f : 0.f);
    const float sc = (c & 1) ? d.sy : d.sx;
    const int lim = (c & 1) ? d.H - 1 : d.W - 1;
    const int v = (c & 2) ? min(lim, (int)ceilf((e + 1.5f) / sc) + 1) : max(0, (int)floorf((e - 1.5f) / sc) - 1);
    rXa = __shfl_sync(0xffffffffu, v, 0);
    rYa = __shfl_sync(0xffffffffu, v, 1);
    rXb = __shfl_sync(0xffffffffu, v, 2);
    rYb = __shfl_sync(0xffffffffu, v, 3);
  }
  const int g = (wt % NG8) * 8 + (lane & 7);
  const int Ybeg = ((wt / NG8) * 4 + (lane >> 3)) * kGenRows;
  const int Yend = min(Ybeg + kGenRows, d.H);
  if (Ybeg >= d.H) return;                          // the whole lane group leaves together
  const bool gok = g < NG;
  const float* L = logits + inst * d.mh * d.mw;
  RowSink sink{d, sk.rowsum + inst * (size_t)d.H * d.nblk,
               (!kWriteMasks && sk.bits) ? reinterpret_cast<uint16_t*>(sk.bits) + inst * (size_t)d.H * (2 * d.bit_words) : nullptr,
               sk.lattice + inst * (size_t)d.lat_rows * d.lat_words, g, lane & 7, gok, LeaderStats()};
  const int X0 = 16 * g, X1 = min(X0 + 15, d.W - 1);
  const int npx = gok ? X1 - X0 + 1 : 0;
  const bool col_in = gok && !(X1 < rXa || X0 > rXb);
  int xa = 0, xb = 0;                        // proto columns the 16 pixels can read
  if (col_in) {
    int t;
    float f0, f1;
    src_index(d.sx, X0, d.mw, xa, t, f0, f1);
    src_index(d.sx, X1, d.mw, t, xb, f0, f1);
  }
  uint8_t* M = (kWriteMasks && gok) ? masks + inst * (size_t)d.H * d.W + X0 : nullptr;
  const bool vec = (npx == 16) && ((d.W & 15) == 0);
  // whole thread tile outside the rectangle (the common case with many small instances): zeros, no logit is read
  const bool tile_out = !col_in || Yend <= rYa || Ybeg > rYb;
  if (__all_sync(0xffffffffu, tile_out || !gok)) {
    // the whole warp tile (128 px x 32 rows) lies outside the instance's rectangle: zeros in the masks (and the bit
    // rows), no summary entry to write - the array rests at zero
    if (gok) {
      for (int Y = Ybeg; Y < Yend; ++Y) {
        if (M) {
          uint8_t* Mr = M + (size_t)Y * d.W;
          if (vec) *reinterpret_cast<uint4*>(Mr) = make_uint4(0u, 0u, 0u, 0u);
          else for (int px = 0; px < npx; ++px) Mr[px] = 0;
        }
        if (sink.bits_inst) sink.bits_inst[(size_t)Y * (2 * d.bit_words) + g] = 0;
      }
    }
    return;
  }
  int py0 = -1, py1 = -1;
  float mn = 0.f, mx = 0.f;
#pragma unroll 1
  for (int Y = Ybeg; Y < Yend; ++Y) {
    unsigned ww[4] = {0u, 0u, 0u, 0u};
    if (!tile_out && Y >= rYa && Y <= rYb) {
      int y0, y1;
      float ly0, ly1;
      src_index(d.sy, Y, d.mh, y0, y1, ly0, ly1);
      const float* r0 = L + (size_t)y0 * d.mw;
      const float* r1 = L + (size_t)y1 * d.mw;
      if (y0 != py0 || y1 != py1) {            // sign range of every tap these 16 pixels can read in rows y0, y1
        mn = INFINITY; mx = -INFINITY;
        for (int x = xa; x <= xb; ++x) {
          const float u = __ldg(r0 + x), v = __ldg(r1 + x);
          mn = fminf(mn, fminf(u, v));
          mx = fmaxf(mx, fmaxf(u, v));
        }
        py0 = y0; py1 = y1;
      }
      if (mn > kTiny) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int left = npx - 4 * q;        // pixels of this word that exist
          ww[q] = left >= 4 ? 0x01010101u : left <= 0 ? 0u : (0x01010101u >> (8 * (4 - left)));
        }
      } else if (mx > 0.f) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          unsigned acc = 0;
#pragma unroll 1
          for (int k = 0; k < 4 && 4 * q + k < npx; ++k) {
            int x0, x1;
            float lx0, lx1;
            src_index(d.sx, X0 + 4 * q + k, d.mw, x0, x1, lx0, lx1);
            const float top = fmaf(__ldg(r0 + x0), lx0, __fmul_rn(__ldg(r0 + x1), lx1));
            const float bot = fmaf(__ldg(r1 + x0), lx0, __fmul_rn(__ldg(r1 + x1), lx1));
            const float o = fmaf(top, ly0, __fmul_rn(bot, ly1));
            if (o > 0.f) acc |= 1u << (8 * k);
          }
          ww[q] = acc;
        }
      }
    }
    const uint4 w = make_uint4(ww[0], ww[1], ww[2], ww[3]);
    if (M) {
      uint8_t* Mr = M + (size_t)Y * d.W;
      if (vec) *reinterpret_cast<uint4*>(Mr) = w;
      else for (int px = 0; px < npx; ++px) Mr[px] = (uint8_t)(((px < 4 ? w.x : px < 8 ? w.y : px < 12 ? w.z : w.w) >> (8 * (px & 3))) & 1u);
    }
    sink.row(pat16(w), Y, true);
  }
  if ((lane & 7) == 0) publish_leader(sink.ls, sk.stats + inst);
}

// ---------------------------------------------------------------------------------------------
// by-products from caller-provided u8 masks (va_mask_to_records)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kUpThreads)
mask_stats_kernel(Dims d, const uint8_t* __restrict__ masks, const int* __restrict__ counts, MaskSinks sk) {
  const int b = blockIdx.z, i = blockIdx.y;
  if (i >= min(counts[b], d.max_n)) return;
  const int NG = ceil_div(d.W, 16);
  const int NG8 = ceil_div(NG, 8);
  const int NY4 = ceil_div(d.H, 4);
  const int wt = blockIdx.x * (kUpThreads / 32) + (threadIdx.x >> 5);
  if (wt >= NG8 * NY4) return;
  const int lane = threadIdx.x & 31;
  const int g = (wt % NG8) * 8 + (lane & 7);
  const int Y = (wt / NG8) * 4 + (lane >> 3);
  const size_t inst = (size_t)b * d.max_n + i;
  RowSink sink{d, sk.rowsum + inst * (size_t)d.H * d.nblk, nullptr, sk.lattice + inst * (size_t)d.lat_rows * d.lat_words,
               g, lane & 7, g < NG, LeaderStats()};
  unsigned pat = 0;
  if (g < NG && Y < d.H) {
    const uint8_t* M = masks + inst * (size_t)d.H * d.W + (size_t)Y * d.W + 16 * g;
    const int npx = min(16, d.W - 16 * g);
    if (npx == 16 && (d.W & 15) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(M));
      const unsigned vv[4] = {v.x, v.y, v.z, v.w};
      unsigned ww[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // normalise non-zero bytes to 1
        unsigned t = vv[q];
        t |= t >> 4; t |= t >> 2; t |= t >> 1;
        ww[q] = t & 0x01010101u;
      }
      pat = pat16(make_uint4(ww[0], ww[1], ww[2], ww[3]));
    } else {
      for (int px = 0; px < npx; ++px) if (M[px]) pat |= 1u << px;
    }
  }
  sink.row(pat, Y, Y < d.H);
  if ((lane & 7) == 0) publish_leader(sink.ls, sk.stats + inst);
}

__global__ void init_scratch_kernel(InstStats* stats, size_t n_stats, unsigned* lattice, size_t n_lat) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = t; k < n_stats; k += stride) {
    InstStats s;
    s.area = 0; s.minx = INT_MAX; s.miny = INT_MAX; s.maxx = -1; s.maxy = -1; s.euler4 = 0; s.pad0 = 0; s.pad1 = 0;
    stats[k] = s;
  }
  for (size_t k = t; k < n_lat; k += stride) lattice[k] = 0u;
}

// ---------------------------------------------------------------------------------------------
cudaError_t launch_upsample(const Dims& d, const float* logits, const float* boxes, const int* counts, int B, uint8_t* masks,
                            const MaskSinks& sk, cudaStream_t st) {
  const bool x4 = (d.H == 4 * d.mh) && (d.W == 4 * d.mw) && (d.mw % 4 == 0);
  if (x4) {
    const int warps = ceil_div(d.W >> 4, 8) * ceil_div(d.mh, 4 * kStrip);
    dim3 grid(ceil_div(warps, kUpThreads / 32), d.max_n, B);
    if (masks) upsample4x_kernel<true><<<grid, kUpThreads, 0, st>>>(d, logits, counts, masks, sk);
    else upsample4x_kernel<false><<<grid, kUpThreads, 0, st>>>(d, logits, counts, masks, sk);
  } else {
    const int warps = ceil_div(ceil_div(d.W, 16), 8) * ceil_div(d.H, 4 * kGenRows);
    dim3 grid(ceil_div(warps, kUpThreads / 32), d.max_n, B);
    if (masks) upsample_generic_kernel<true><<<grid, kUpThreads, 0, st>>>(d, logits, boxes, counts, masks, sk);
    else upsample_generic_kernel<false><<<grid, kUpThreads, 0, st>>>(d, logits, boxes, counts, masks, sk);
  }
  return cudaGetLastError();
}

cudaError_t launch_mask_stats(const Dims& d, const uint8_t* masks, const int* counts, int B, const MaskSinks& sk,
                              cudaStream_t st) {
  const int warps = ceil_div(ceil_div(d.W, 16), 8) * ceil_div(d.H, 4);
  dim3 grid(ceil_div(warps, kUpThreads / 32), d.max_n, B);
  mask_stats_kernel<<<grid, kUpThreads, 0, st>>>(d, masks, counts, sk);
  return cudaGetLastError();
}

cudaError_t launch_init_scratch(const Dims& d, int max_batch, InstStats* stats, unsigned* lattice, cudaStream_t st) {
  const size_t ns = (size_t)max_batch * d.max_n;
  const size_t nl = ns * d.lat_rows * d.lat_words;
  init_scratch_kernel<<<256, 256, 0, st>>>(stats, ns, lattice, nl);
  return cudaGetLastError();
}

}  // namespace va
