// Fused mask assembly for sm_100a: TMA-fed tcgen05 (TMEM accumulator) prototype x coefficient
// contraction, box crop, exact-4x bilinear upsample, threshold, u8 mask store and the per-instance
// reductions for the grid stage - one persistent kernel, nothing but the prototypes is read from
// HBM and nothing but the masks (and a few hundred bytes of reductions) is written.
//
// Reference semantics: ops.process_mask (testing/old/segmenting_using_tflite/ops.py:707-737).
//
// Work item = (frame, band of prototype rows).  Inside a CTA, warp-specialised roles connected by
// mbarriers (no CTA-wide barrier inside the steady state):
//
//   warp 0      TMA producer  : [32 prototypes x 128 pixels] fp32 boxes of the pixel-contiguous [K, P]
//                               prototype matrix into a staging ring (kStagesHi stages).
//   warp 1      MMA issuer    : D[128 px, N inst] = A * B^T with kind::tf32, M=128, N=16, K=8 per
//                               instruction, both operands K-major with 128B swizzle (tcgen05 does not
//                               transpose 32-bit operands: an MN-major tf32 A returns zeros, measured with
//                               tests/micro/umma_probe.cu).  fp32-class accuracy comes from a 3-pass split
//                               A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, hi = top 19 bits (what the tensor core
//                               reads; it truncates, measured), lo = x - hi.  Accumulators: kAcc TMEM tiles.
//   warps 2-5   split+epilogue: (front) transpose the staged [k][px] box into the K-major [px][k] operand
//                               tiles A_hi / A_lo (one pixel row per thread, swizzled 16 B chunks), per-frame
//                               B tiles (coefficients hi/lo) ; (back, lagging) tcgen05.ld of finished
//                               accumulators, box crop, store of the cropped logits into row-chunk buffers.
//   warps 6-15  upsample      : per chunk of PR row pairs: 4-tap blend with torch's exact roundings,
//                               > 0, 16-byte mask stores, area / bbox / lattice reductions.
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "va_up_common.cuh"

namespace va {

constexpr int kTileM = 128;                 // pixels per MMA tile (TMEM lanes)
constexpr int kTileBytes = kTileM * kProtoK * 4;   // 16 KB
constexpr int kStagesHi = 3;                // TMA staging ring ([32 k][128 px] boxes, no swizzle)
constexpr int kStagesLo = 2;                // operand ring: one A_hi + one A_lo K-major tile per stage
constexpr int kAcc = 4;                     // TMEM accumulator ring
constexpr int kNPad = 16;                   // UMMA N (instances padded)
constexpr int kLag = 2;                     // epilogue runs kLag tiles behind the split front
constexpr int kChunkBufs = 3;
constexpr int kWarpsSplit = 4;
constexpr int kWarpsUp = 10;
constexpr int kThreads = 32 * (2 + kWarpsSplit + kWarpsUp);   // 512
constexpr int kUpThreadsTc = 32 * kWarpsUp;
constexpr int kMaxInstTc = 16;              // instances this kernel handles (N = 16)
constexpr int kTmemCols = kAcc * kNPad;     // 64: power of two >= 32

struct FusedParams {
  Dims d;
  const float* coefs;
  const float* boxes;
  const int* counts;
  uint8_t* masks;
  float* logits_dbg;
  InstStats* stats;
  unsigned* lattice;
  int B;
  int nbands;        // bands per frame
  int ppb;           // row pairs per band
  int pr;            // row pairs per chunk
  int n_items;
  int nst;           // instance stride of the chunk buffers (= max_n)
  int chunk_floats;  // floats per chunk buffer = (pr+1) * nst * mw
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128B swizzle, version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// work decomposition
// ---------------------------------------------------------------------------------------------
struct Item {
  int valid;
  int b;        // frame
  int pa, pb;   // row pairs [pa, pb): pair r blends proto rows r and min(r+1, mh-1) into dst rows 4r+2..4r+5
  int nrows;    // proto rows pa .. min(pb, mh-1)
  int npx;      // nrows * mw
  int ntiles;   // ceil(npx / 128)
  int nchunks;  // ceil((pb - pa) / pr)
};

__device__ __forceinline__ Item get_item(const FusedParams& p, int k) {
  Item it;
  const int item = blockIdx.x + k * gridDim.x;
  it.valid = item < p.n_items;
  const int b = item / p.nbands, j = item - b * p.nbands;
  it.b = b;
  it.pa = j * p.ppb;
  it.pb = min(it.pa + p.ppb, p.d.mh);
  it.nrows = min(it.pb, p.d.mh - 1) - it.pa + 1;
  it.npx = it.nrows * p.d.mw;
  it.ntiles = ceil_div(it.npx, kTileM);
  it.nchunks = ceil_div(it.pb - it.pa, p.pr);
  return it;
}

// last band-local proto row that chunk c needs
__device__ __forceinline__ int chunk_last_row(const FusedParams& p, const Item& it, int c) {
  return min((c + 1) * p.pr, it.nrows - 1);
}

struct SmemLayout {
  uint8_t* hi;        // [kStagesHi][16 KB]   TMA staging, [32 k][128 px]
  uint8_t* lo;        // [kStagesLo][A_hi 16 KB | A_lo 16 KB]   K-major SW128 operand tiles, 1024 B aligned
  uint8_t* bt;        // [2 parity][hi, lo][kNPad * 128 B]
  float* chunks;      // [kChunkBufs][chunk_floats]
  float* box;         // [4 items ring][kMaxInstTc][4]   (the epilogue lags the front by < 4 items)
  int* stat;          // [kMaxInstTc][8]  area, minx, miny, maxx, maxy
  short* latrow;      // [H]  lattice row index of dst row Y, -1 if none
  uint64_t* bars;
  uint32_t* tmem_slot;
};
enum {
  BAR_HI_FULL = 0,
  BAR_HI_EMPTY = BAR_HI_FULL + kStagesHi,
  BAR_LO_FULL = BAR_HI_EMPTY + kStagesHi,
  BAR_LO_EMPTY = BAR_LO_FULL + kStagesLo,
  BAR_ACC_FULL = BAR_LO_EMPTY + kStagesLo,
  BAR_ACC_EMPTY = BAR_ACC_FULL + kAcc,
  BAR_B_FULL = BAR_ACC_EMPTY + kAcc,
  BAR_B_EMPTY = BAR_B_FULL + 2,
  BAR_CH_FULL = BAR_B_EMPTY + 2,
  BAR_CH_EMPTY = BAR_CH_FULL + kChunkBufs,
  BAR_COUNT = BAR_CH_EMPTY + kChunkBufs
};

__host__ __device__ inline size_t fused_smem_layout(int chunk_floats, int H, SmemLayout* s, uint8_t* base) {
  size_t o = 0;
  auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) / align * align; size_t r = o; o += bytes; return r; };
  const size_t o_hi = take((size_t)kStagesHi * kTileBytes, 1024);
  const size_t o_lo = take((size_t)kStagesLo * 2 * kTileBytes, 1024);
  const size_t o_bt = take((size_t)2 * 2 * kNPad * 128, 1024);
  const size_t o_ch = take((size_t)kChunkBufs * chunk_floats * 4, 16);
  const size_t o_box = take((size_t)4 * kMaxInstTc * 4 * 4, 16);
  const size_t o_st = take((size_t)kMaxInstTc * 8 * 4, 16);
  const size_t o_lr = take((size_t)H * 2, 16);
  const size_t o_bar = take((size_t)BAR_COUNT * 8, 8);
  const size_t o_tm = take(16, 16);
  if (s) {
    s->hi = base + o_hi; s->lo = base + o_lo; s->bt = base + o_bt; s->chunks = (float*)(base + o_ch);
    s->box = (float*)(base + o_box); s->stat = (int*)(base + o_st); s->latrow = (short*)(base + o_lr);
    s->bars = (uint64_t*)(base + o_bar); s->tmem_slot = (uint32_t*)(base + o_tm);
  }
  return o;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool kWriteMasks>
__global__ void __launch_bounds__(kThreads, 1)
fused_tc_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  extern __shared__ uint8_t smem_dyn[];
  // dynamic shared memory is only guaranteed 16 B aligned: align to 1024 for the 128B-swizzle tiles
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  SmemLayout s;
  fused_smem_layout(p.chunk_floats, p.d.H, &s, base);
  const Dims& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesHi; ++i) { mbar_init(&s.bars[BAR_HI_FULL + i], 1); mbar_init(&s.bars[BAR_HI_EMPTY + i], kWarpsSplit); }
    for (int i = 0; i < kStagesLo; ++i) { mbar_init(&s.bars[BAR_LO_FULL + i], kWarpsSplit); mbar_init(&s.bars[BAR_LO_EMPTY + i], 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(&s.bars[BAR_ACC_FULL + i], 1); mbar_init(&s.bars[BAR_ACC_EMPTY + i], kWarpsSplit); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s.bars[BAR_B_FULL + i], kWarpsSplit); mbar_init(&s.bars[BAR_B_EMPTY + i], 1); }
    for (int i = 0; i < kChunkBufs; ++i) { mbar_init(&s.bars[BAR_CH_FULL + i], kWarpsSplit); mbar_init(&s.bars[BAR_CH_EMPTY + i], kWarpsUp); }
    fence_barrier_init();
  }
  for (int y = threadIdx.x; y < d.H; y += kThreads) {
    const int t = y - (d.gs >> 1);
    s.latrow[y] = (t >= 0 && t % d.gs == 0) ? (short)(t / d.gs) : (short)-1;
  }
  for (int t = threadIdx.x; t < kMaxInstTc * 8; t += kThreads) {
    const int f = t & 7;
    s.stat[t] = (f == 1 || f == 2) ? INT_MAX : (f == 3 || f == 4) ? -1 : 0;
  }
  if (warp == 1) tmem_alloc(s.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t g = 0;   // global tile counter of this CTA
      for (int k = 0;; ++k) {
        const Item it = get_item(p, k);
        if (!it.valid) break;
        const int px0 = it.pa * d.mw;
        for (int t = 0; t < it.ntiles; ++t, ++g) {
          const int st = g % kStagesHi;
          mbar_wait(&s.bars[BAR_HI_EMPTY + st], ((g / kStagesHi) & 1) ^ 1);
          mbar_arrive_expect_tx(&s.bars[BAR_HI_FULL + st], kTileBytes);
          tma_load_3d(s.hi + (size_t)st * kTileBytes, &tmap, &s.bars[BAR_HI_FULL + st], px0 + t * kTileM, 0, it.b);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kTileM, kNPad);
      uint32_t g = 0;
      for (int k = 0;; ++k) {
        const Item it = get_item(p, k);
        if (!it.valid) break;
        const int par = k & 1;
        mbar_wait(&s.bars[BAR_B_FULL + par], (k >> 1) & 1);
        const uint32_t b_hi = smem_u32(s.bt + (size_t)(par * 2 + 0) * kNPad * 128);
        const uint32_t b_lo = smem_u32(s.bt + (size_t)(par * 2 + 1) * kNPad * 128);
        for (int t = 0; t < it.ntiles; ++t, ++g) {
          const int sl = g % kStagesLo, ac = g % kAcc;
          mbar_wait(&s.bars[BAR_ACC_EMPTY + ac], ((g / kAcc) & 1) ^ 1);
          mbar_wait(&s.bars[BAR_LO_FULL + sl], (g / kStagesLo) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(s.lo + (size_t)sl * 2 * kTileBytes);
          const uint32_t a_lo = a_hi + kTileBytes;
          const uint32_t dcol = tmem_base + ac * kNPad;
          // K-major SW128 operands: rows (pixels / instances) are 128 B, 8-row groups 1024 B apart (SBO),
          // one k-step (8 tf32) = 32 B along the row.
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(dcol, make_smem_desc(a_hi + ks * 32, 16, 1024), make_smem_desc(b_hi + ks * 32, 16, 1024), idesc, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(dcol, make_smem_desc(a_hi + ks * 32, 16, 1024), make_smem_desc(b_lo + ks * 32, 16, 1024), idesc, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(dcol, make_smem_desc(a_lo + ks * 32, 16, 1024), make_smem_desc(b_hi + ks * 32, 16, 1024), idesc, 1);
          umma_commit(&s.bars[BAR_LO_EMPTY + sl]);
          umma_commit(&s.bars[BAR_ACC_FULL + ac]);
        }
        umma_commit(&s.bars[BAR_B_EMPTY + par]);
      }
    }
  } else if (warp < 2 + kWarpsSplit) {
    // =========================== split (front) + epilogue (back) ===========================
    const int sw = warp - 2;                 // 0..3
    const int st_tid = sw * 32 + lane;       // 0..127
    const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are accessible to this warp
    // front iterator
    int fk = 0, ft = 0;
    Item fit = get_item(p, 0);
    uint32_t fg = 0;
    // back iterator
    int bk = 0, bt = 0;
    Item bit = get_item(p, 0);
    uint32_t bg = 0;
    uint32_t chunk_base = 0;                 // global chunk index of chunk 0 of the back item
    int acquired = 0, completed = 0;         // chunks of the back item acquired for writing / signalled full
    int lagged = 0;

    while (fit.valid || bit.valid) {
      // ---------------- front: B tiles at the start of an item, A_lo for tile (fk, ft) ----------------
      if (fit.valid) {
        if (ft == 0) {
          const int par = fk & 1;
          mbar_wait(&s.bars[BAR_B_EMPTY + par], ((fk >> 1) & 1) ^ 1);
          const int n = min(p.counts[fit.b], min(d.max_n, kMaxInstTc));
          uint8_t* bh = s.bt + (size_t)(par * 2 + 0) * kNPad * 128;
          uint8_t* bl = s.bt + (size_t)(par * 2 + 1) * kNPad * 128;
          // row r (instance), 16 B chunk c: stored at chunk c ^ (r & 7)  (128B swizzle, K-major)
          for (int q = st_tid; q < kNPad * 8; q += 128) {
            const int r = q >> 3, c = q & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < n) v = __ldg(reinterpret_cast<const float4*>(p.coefs + ((size_t)fit.b * d.max_n + r) * d.K + 4 * c));
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
            const int off = r * 128 + ((c ^ (r & 7)) << 4);
            *reinterpret_cast<float4*>(bh + off) = h;
            *reinterpret_cast<float4*>(bl + off) = l;
          }
          for (int q = st_tid; q < kMaxInstTc * 4; q += 128) {
            const int i = q >> 2, c = q & 3;
            float v = 0.f;
            if (i < n) v = __fmul_rn(__ldg(p.boxes + ((size_t)fit.b * d.max_n + i) * 4 + c), (c & 1) ? d.hr : d.wr);
            s.box[((fk & 3) * kMaxInstTc + i) * 4 + c] = v;
          }
          fence_proxy_async();
          named_bar_sync(1, 32 * kWarpsSplit);
          if (lane == 0) mbar_arrive(&s.bars[BAR_B_FULL + par]);
        }
        const int sh = fg % kStagesHi, sl = fg % kStagesLo;
        mbar_wait(&s.bars[BAR_HI_FULL + sh], (fg / kStagesHi) & 1);
        mbar_wait(&s.bars[BAR_LO_EMPTY + sl], ((fg / kStagesLo) & 1) ^ 1);
        // transpose + split: thread = pixel row of the tile.  Reads of [k][px] are conflict-free across
        // the warp (consecutive px); each 16 B chunk c of the K-major row lands at chunk c ^ (px & 7).
        const float* src = reinterpret_cast<const float*>(s.hi + (size_t)sh * kTileBytes) + st_tid;
        uint8_t* dhi = s.lo + (size_t)sl * 2 * kTileBytes + (size_t)st_tid * 128;
        uint8_t* dlo = dhi + kTileBytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v, h, l;
          v.x = src[(4 * c + 0) * kTileM]; v.y = src[(4 * c + 1) * kTileM];
          v.z = src[(4 * c + 2) * kTileM]; v.w = src[(4 * c + 3) * kTileM];
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
          const int off = (c ^ (st_tid & 7)) << 4;
          *reinterpret_cast<float4*>(dhi + off) = h;
          *reinterpret_cast<float4*>(dlo + off) = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s.bars[BAR_LO_FULL + sl]);
          mbar_arrive(&s.bars[BAR_HI_EMPTY + sh]);
        }
        ++fg;
        if (++ft == fit.ntiles) { ft = 0; fit = get_item(p, ++fk); }
      }
      // ---------------- back: epilogue of the tile kLag steps behind ----------------
      if (lagged < kLag && fit.valid) { ++lagged; continue; }
      if (bit.valid) {
        const int ac = bg % kAcc;
        mbar_wait(&s.bars[BAR_ACC_FULL + ac], (bg / kAcc) & 1);
        __syncwarp();
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + ac * kNPad, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.bars[BAR_ACC_EMPTY + ac]);

        const int n = min(p.counts[bit.b], min(d.max_n, kMaxInstTc));
        const int px = bt * kTileM + quarter * 32 + lane;     // band-local pixel of this thread
        const bool live = px < bit.npx;
        const int row = px / d.mw, col = px - row * d.mw;     // band-local row
        // chunks touched by this tile (warp-uniform bounds): rows [row_first, row_last]
        const int tile_last_px = min((bt + 1) * kTileM, bit.npx) - 1;
        const int row_last = tile_last_px / d.mw;
        const int c_hi = min(row_last / p.pr, bit.nchunks - 1);
        while (acquired <= c_hi) {                            // acquire chunk buffers in order
          const uint32_t gc = chunk_base + acquired;
          mbar_wait(&s.bars[BAR_CH_EMPTY + gc % kChunkBufs], ((gc / kChunkBufs) & 1) ^ 1);
          ++acquired;
        }
        if (live) {
          const float fx = (float)col, fy = (float)(bit.pa + row);
          const int c1 = row / p.pr, rr = row - c1 * p.pr;
          float* dst1 = (c1 < bit.nchunks)
                            ? s.chunks + (size_t)((chunk_base + c1) % kChunkBufs) * p.chunk_floats + (size_t)rr * p.nst * d.mw + col
                            : nullptr;
          float* dst0 = (rr == 0 && c1 > 0)
                            ? s.chunks + (size_t)((chunk_base + c1 - 1) % kChunkBufs) * p.chunk_floats + (size_t)p.pr * p.nst * d.mw + col
                            : nullptr;
          const float* bx = s.box + (bk & 3) * kMaxInstTc * 4;
#pragma unroll
          for (int i = 0; i < kMaxInstTc; ++i) {
            if (i < n) {
              const bool keep = (fx >= bx[4 * i]) && (fx < bx[4 * i + 2]) && (fy >= bx[4 * i + 1]) && (fy < bx[4 * i + 3]);
              const float v = keep ? __uint_as_float(r[i]) : 0.f;
              if (dst1) dst1[(size_t)i * d.mw] = v;
              if (dst0) dst0[(size_t)i * d.mw] = v;
              if (p.logits_dbg) p.logits_dbg[(((size_t)bit.b * d.max_n + i) * d.mh + bit.pa + row) * d.mw + col] = v;
            }
          }
        }
        __syncwarp();
        // chunks completed by this tile
        const int rows_done = (bt + 1 == bit.ntiles) ? bit.nrows : ((bt + 1) * kTileM) / d.mw;   // complete rows so far
        while (completed < bit.nchunks && chunk_last_row(p, bit, completed) < rows_done) {
          const uint32_t gc = chunk_base + completed;
          if (lane == 0) mbar_arrive(&s.bars[BAR_CH_FULL + gc % kChunkBufs]);
          ++completed;
        }
        ++bg;
        if (++bt == bit.ntiles) {
          bt = 0;
          chunk_base += bit.nchunks;
          acquired = 0;
          completed = 0;
          bit = get_item(p, ++bk);
        }
      }
    }
  } else {
    // =========================== upsample + threshold + store + reductions ===========================
    const int ut = threadIdx.x - 32 * (2 + kWarpsSplit);     // 0..319
    const int NG = d.W >> 4, NG8 = ceil_div(NG, 8);
    const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    const uint4 zeros = make_uint4(0u, 0u, 0u, 0u);
    uint32_t gc = 0;
    for (int k = 0;; ++k) {
      const Item it = get_item(p, k);
      if (!it.valid) break;
      const int n = min(p.counts[it.b], min(d.max_n, kMaxInstTc));
      for (int c = 0; c < it.nchunks; ++c, ++gc) {
        const int buf = gc % kChunkBufs;
        mbar_wait(&s.bars[BAR_CH_FULL + buf], (gc / kChunkBufs) & 1);
        const float* cb = s.chunks + (size_t)buf * p.chunk_floats;
        const int r0 = it.pa + c * p.pr;                        // first pair of the chunk
        const int npairs = min(p.pr, it.pb - r0);
        const int ntasks = n * NG8 * p.pr * 8;                  // gl (8) fastest, then pair, g8, instance
        for (int q = ut; q < ntasks; q += kUpThreadsTc) {
          const int gl = q & 7;
          int rest = q >> 3;
          const int pair = rest % p.pr; rest /= p.pr;
          const int g8 = rest % NG8;
          const int i = rest / NG8;
          const int g = g8 * 8 + gl;
          if (g >= NG || pair >= npairs) continue;
          const int r = r0 + pair;
          const bool last = (r == d.mh - 1);
          const float* rowA = cb + ((size_t)pair * p.nst + i) * d.mw;
          const float* rowB = last ? rowA : rowA + (size_t)p.nst * d.mw;
          float sA[6], sB[6];
          {
            const float4 v = *reinterpret_cast<const float4*>(rowA + 4 * g);
            sA[1] = v.x; sA[2] = v.y; sA[3] = v.z; sA[4] = v.w;
            sA[0] = (g > 0) ? rowA[4 * g - 1] : v.x;
            sA[5] = (4 * g + 4 < d.mw) ? rowA[4 * g + 4] : v.w;
            const float4 u = *reinterpret_cast<const float4*>(rowB + 4 * g);
            sB[1] = u.x; sB[2] = u.y; sB[3] = u.z; sB[4] = u.w;
            sB[0] = (g > 0) ? rowB[4 * g - 1] : u.x;
            sB[5] = (4 * g + 4 < d.mw) ? rowB[4 * g + 4] : u.w;
          }
          const float mnA = fminf(fminf(fminf(sA[0], sA[1]), fminf(sA[2], sA[3])), fminf(sA[4], sA[5]));
          const float mxA = fmaxf(fmaxf(fmaxf(sA[0], sA[1]), fmaxf(sA[2], sA[3])), fmaxf(sA[4], sA[5]));
          const float mnB = fminf(fminf(fminf(sB[0], sB[1]), fminf(sB[2], sB[3])), fminf(sB[4], sB[5]));
          const float mxB = fmaxf(fmaxf(fmaxf(sB[0], sB[1]), fmaxf(sB[2], sB[3])), fmaxf(sB[4], sB[5]));
          const size_t inst = (size_t)it.b * d.max_n + i;
          uint8_t* M = kWriteMasks ? p.masks + inst * (size_t)d.H * d.W + 16 * g : nullptr;
          unsigned* lat = p.lattice + inst * (size_t)d.lat_rows * d.lat_words;
          ThreadStats ts;
          auto emit = [&](const uint4& w, int Y) {
            if (kWriteMasks) *reinterpret_cast<uint4*>(M + (size_t)Y * d.W) = w;
            ts.add_row(w, Y);
            if ((w.x | w.y | w.z | w.w) && s.latrow[Y] >= 0) lattice_row(w, Y, 16 * g, d, lat);
          };
          const bool left = (g == 0);
          const bool uni_pos = fminf(mnA, mnB) > kTiny, uni_neg = fmaxf(mxA, mxB) <= 0.f;
          const int Y0 = 4 * r + 2;
          float hA[16], hB[16];
          const bool need_h = !(uni_pos || uni_neg);
          if (need_h) { hinterp4(sA, hA, left); hinterp4(sB, hB, left); }
          if (r == 0) {       // dst rows 0,1 take h(row 0) unchanged (src y clamps to 0)
            uint4 w;
            if (mnA > kTiny) w = ones;
            else if (mxA <= 0.f) w = zeros;
            else { if (!need_h) hinterp4(sA, hA, left); w = hpack(hA); }
            emit(w, 0);
            emit(w, 1);
          }
          if (uni_pos) {
            emit(ones, Y0); emit(ones, Y0 + 1);
            if (!last) { emit(ones, Y0 + 2); emit(ones, Y0 + 3); }
          } else if (uni_neg) {
            emit(zeros, Y0); emit(zeros, Y0 + 1);
            if (!last) { emit(zeros, Y0 + 2); emit(zeros, Y0 + 3); }
          } else {
            emit(vblend(hA, hB, 0.875f, 0.125f), Y0);
            emit(vblend(hA, hB, 0.625f, 0.375f), Y0 + 1);
            if (!last) {
              emit(vblend(hA, hB, 0.375f, 0.625f), Y0 + 2);
              emit(vblend(hA, hB, 0.125f, 0.875f), Y0 + 3);
            }
          }
          ts.flush();
          if (ts.area) {
            int minx = INT_MAX, maxx = -1;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
              if (ts.orw[w4]) {
                minx = min(minx, 16 * g + 4 * w4 + ((__ffs(ts.orw[w4]) - 1) >> 3));
                maxx = max(maxx, 16 * g + 4 * w4 + ((31 - __clz(ts.orw[w4])) >> 3));
              }
            }
            int* st = s.stat + i * 8;
            atomicAdd(&st[0], (int)ts.area);
            atomicMin(&st[1], minx);
            atomicMin(&st[2], ts.miny);
            atomicMax(&st[3], maxx);
            atomicMax(&st[4], ts.maxy);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.bars[BAR_CH_EMPTY + buf]);
      }
      // ---- item done: publish the band's reductions for frame it.b ----
      named_bar_sync(2, kUpThreadsTc);
      if (ut < n) {
        int* st = s.stat + ut * 8;
        if (st[0]) {
          InstStats* dst = p.stats + (size_t)it.b * d.max_n + ut;
          atomicAdd(&dst->area, (unsigned)st[0]);
          atomicMin(&dst->minx, st[1]);
          atomicMin(&dst->miny, st[2]);
          atomicMax(&dst->maxx, st[3]);
          atomicMax(&dst->maxy, st[4]);
        }
        st[0] = 0; st[1] = INT_MAX; st[2] = INT_MAX; st[3] = -1; st[4] = -1;
      }
      named_bar_sync(2, kUpThreadsTc);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct FusedPlan {
  PFN_encodeTiled encode;
  int num_sms;
  int pr;
  int chunk_floats;
  size_t smem_bytes;
  // cached tensor map
  const float* map_ptr;
  int map_B;
  CUtensorMap map;
};

FusedPlan* fused_plan_create(const Dims& d, int device, char* err, size_t errlen) {
  if (!(d.H == 4 * d.mh && d.W == 4 * d.mw)) { snprintf(err, errlen, "tcgen05 path needs H=4*mh, W=4*mw"); return nullptr; }
  if (d.max_n > kMaxInstTc) { snprintf(err, errlen, "tcgen05 path handles max_n <= %d", kMaxInstTc); return nullptr; }
  if ((d.mw % 4) != 0 || (d.W % 16) != 0) { snprintf(err, errlen, "mw %% 4 / W %% 16"); return nullptr; }
  if (((size_t)d.mh * d.mw) % 4 != 0) { snprintf(err, errlen, "P %% 4"); return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { snprintf(err, errlen, "cudaGetDeviceProperties failed"); return nullptr; }
  if (prop.major != 10) { snprintf(err, errlen, "device is sm_%d%d, tcgen05 needs sm_100", prop.major, prop.minor); return nullptr; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled not available");
    return nullptr;
  }
  FusedPlan* pl = new FusedPlan();
  memset(pl, 0, sizeof(*pl));
  pl->encode = (PFN_encodeTiled)fn;
  pl->num_sms = prop.multiProcessorCount;
  // pairs per chunk: largest of {4, 2} whose three chunk buffers fit next to the tile rings
  const size_t limit = (size_t)prop.sharedMemPerBlockOptin - 1024;
  int pr = 0;
  for (int cand : {4, 2}) {
    const int cf = (cand + 1) * d.max_n * d.mw;
    if (fused_smem_layout(cf, d.H, nullptr, nullptr) + 1024 <= limit) { pr = cand; break; }
  }
  if (!pr) { delete pl; snprintf(err, errlen, "chunk buffers do not fit in shared memory (max_n=%d, mw=%d)", d.max_n, d.mw); return nullptr; }
  pl->pr = pr;
  pl->chunk_floats = (pr + 1) * d.max_n * d.mw;
  pl->smem_bytes = fused_smem_layout(pl->chunk_floats, d.H, nullptr, nullptr) + 1024;
  cudaError_t e = cudaFuncSetAttribute(fused_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes);
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(%zu B): %s", pl->smem_bytes, cudaGetErrorString(e)); delete pl; return nullptr; }
  return pl;
}

void fused_plan_destroy(FusedPlan* p) { delete p; }

cudaError_t launch_fused(FusedPlan* pl, const Dims& d, const float* protos, const float* coefs, const float* boxes,
                         const int* counts, int B, uint8_t* masks, float* logits_dbg, InstStats* stats, unsigned* lattice,
                         cudaStream_t st, char* err, size_t errlen) {
  const size_t P = (size_t)d.mh * d.mw;
  if (pl->map_ptr != protos || pl->map_B != B) {
    // [K, P] per frame, pixel-contiguous: dims (px, k, frame); box 128 px x 32 k x 1, no swizzle
    cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)d.K, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)P * d.K * 4};
    cuuint32_t box[3] = {kTileM, kProtoK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = pl->encode(&pl->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)protos, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d)", (int)r); return cudaErrorInvalidValue; }
    pl->map_ptr = protos;
    pl->map_B = B;
  }
  FusedParams p;
  p.d = d; p.coefs = coefs; p.boxes = boxes; p.counts = counts; p.masks = masks; p.logits_dbg = logits_dbg;
  p.stats = stats; p.lattice = lattice; p.B = B;
  p.pr = pl->pr; p.nst = d.max_n; p.chunk_floats = pl->chunk_floats;
  // bands per frame: balance the persistent grid against the one-row halo each band recomputes
  const int max_bands = (d.mh / (2 * pl->pr)) > 0 ? d.mh / (2 * pl->pr) : 1;
  int best_nb = 1;
  double best_eff = 0.0;
  for (int nb = 1; nb <= max_bands && nb <= 64; ++nb) {
    int ppb = ceil_div(ceil_div(d.mh, nb), pl->pr) * pl->pr;
    const int nbands = ceil_div(d.mh, ppb);
    const long items = (long)B * nbands;
    const long rounds = (items + pl->num_sms - 1) / pl->num_sms;
    const double eff = (double)items / (double)(rounds * pl->num_sms) * (double)ppb / (double)(ppb + 1.5);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_nb = nb; }
  }
  p.ppb = ceil_div(ceil_div(d.mh, best_nb), pl->pr) * pl->pr;
  p.nbands = ceil_div(d.mh, p.ppb);
  p.n_items = B * p.nbands;
  const int grid = p.n_items < pl->num_sms ? p.n_items : pl->num_sms;
  if (masks) fused_tc_kernel<true><<<grid, kThreads, pl->smem_bytes, st>>>(pl->map, p);
  else fused_tc_kernel<false><<<grid, kThreads, pl->smem_bytes, st>>>(pl->map, p);
  return cudaGetLastError();
}

}  // namespace va
