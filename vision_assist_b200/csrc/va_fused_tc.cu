// Fused mask assembly for sm_100a: TMA-fed tcgen05 (TMEM accumulator) prototype x coefficient
// contraction, box crop, exact-4x bilinear upsample, threshold, u8 mask store and the per-instance
// reductions for the grid stage - one persistent kernel, nothing but the prototypes is read from
// HBM and nothing but the masks (and a few hundred bytes of reductions) is written.
//
// Reference semantics: ops.process_mask (testing/old/segmenting_using_tflite/ops.py:707-737).
//
// Work item = (band of prototype rows, frame, group of 8 / 16 instances), stolen dynamically from a global
// counter.  Inside a CTA (16 warps), warp-specialised roles connected only by mbarriers:
//
//   warp 0      work stealing + TMA producer: takes the next item, reduces the group's boxes to their hull in
//               prototype rows (rows outside it are zero after crop_mask: neither loaded nor multiplied), publishes
//               {item, live chunk range, instance count} to a shared-memory ring, then loads [32 prototypes x 128
//               pixels] fp32 boxes of the pixel-contiguous [K, P] prototype matrix into a staging ring.
//   warps 1-4   split + MMA issue: one pixel (= TMEM lane) per thread: 32 ld.shared, hi = the raw word (the tensor
//               core truncates fp32 to tf32, measured), lo = x - trunc(x), both written to TENSOR MEMORY with
//               tcgen05.st (tcgen05 does not transpose 32-bit operands - an MN-major tf32 A returns zeros, measured
//               with tests/micro/umma_probe.cu - so the K-major A operand is produced by threads).  Lane 0 of warps
//               1 and 2 issue D[128 px, 32] = A * [B_hi ; B_lo]^T (kind::tf32, A from TMEM, B from smem, K = 8 x 4):
//               issuer 0 with A_hi, issuer 1 with A_lo, into separate accumulators (3xTF32-style split, fp32-class
//               accuracy); MMAs issued from different warps overlap (tests/micro/umma_2issuers.cu).
//   warps 5-8   epilogue: tcgen05.ld of the four partial products, sum, box crop with integer bounds, store of the
//               cropped logits into row-chunk buffers in shared memory (instances far from the tile rows skipped).
//   warps 9-15  upsample: per chunk of `pr` row pairs: 4-tap blend with torch's exact roundings, > 0, 16-byte mask
//               stores, area / bbox / lattice reductions; dead (chunk, instance) pairs and rows outside the hull
//               are bulk zero-filled (cp.async.bulk shared -> global).
//
// The last CTA to finish re-arms the work counter; the tail kernel is a programmatic dependent launch.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "va_up_common.cuh"

namespace va {

constexpr int kTileM = 128;                 // pixels per MMA tile (TMEM lanes)
constexpr int kTileBytes = kTileM * kProtoK * 4;   // 16 KB
constexpr int kStagesHi = 5;                // TMA staging ring ([32 k][128 px] boxes, no swizzle)
constexpr int kStagesLo = 2;                // A-operand ring in TENSOR MEMORY: A_hi (32 cols) + A_lo (32 cols) per stage
constexpr int kAcc = 4;                     // TMEM accumulator ring
constexpr int kNPad = 16;                   // instances padded
constexpr int kNMma = 2 * kNPad;            // UMMA N: [B_hi ; B_lo] stacked along N - the cost of a small MMA does not depend on N
constexpr int kChunkBufs = 4;                // maximum; the plan uses 4 when they fit in shared memory, else 3 (FusedParams::nbuf)
constexpr int kItemRing = 32;               // published item indices.  The TMA thread leads the slowest role by at most
                                            // kStagesHi + kStagesLo + kAcc tiles + kChunkBufs chunks < 64 tiles and every item has >= 2 tiles
                                            // (checked at plan creation), so a slot is never republished before it was read
constexpr int kWarpsSplit = 4;              // one warp per TMEM lane quarter: staged box -> registers -> tcgen05.st
constexpr int kWarpsEpi = 4;                // one warp per TMEM lane quarter: TMEM -> crop -> chunk buffers
constexpr int kWarpsUp = 7;
constexpr int kFirstSplitWarp = 1;          // warp 0: TMA producer
constexpr int kFirstEpiWarp = kFirstSplitWarp + kWarpsSplit;
constexpr int kFirstUpWarp = kFirstEpiWarp + kWarpsEpi;
constexpr int kThreads = 32 * (kFirstUpWarp + kWarpsUp);   // 512
constexpr int kUpThreadsTc = 32 * kWarpsUp;
constexpr int kMaxInstTc = 16;              // instances per group (accumulator columns of one pass: N = 16)
constexpr int kTmemAOff = 0;                // TMEM columns [0, 128): A ring
constexpr int kIssuers = 2;                 // MMA-issuing threads (lane 0 of split warps 0 and 1): tcgen05.mma from different warps overlap
constexpr int kAccCols = kIssuers * kNMma;  // accumulator columns per tile: [hi pass | lo pass], each [B_hi | B_lo]
constexpr int kTmemAccOff = kStagesLo * 64; // TMEM columns [128, 384): accumulators (kAcc x kAccCols)
constexpr int kTmemCols = 512;

struct FusedParams {
  Dims d;
  const float* coefs;
  const float* boxes;
  const int* counts;
  uint8_t* masks;
  float* logits_dbg;
  InstStats* stats;
  unsigned* lattice;
  uint32_t* rowsum;  // [B][max_n][H][nblk] per-(row, 128 px block) summaries (va_contour_core.h)
  uint16_t* bits16;  // [B][max_n][H][2 * bit_words] bit-packed masks, written when the u8 masks are not (else nullptr)
  int B;
  int nbands;        // bands per frame
  int ppb;           // row pairs per band
  int pr;            // row pairs per chunk (2 or 4)
  int pr_shift;      // log2(pr)
  int n_items;
  int groups;        // instance groups per frame: group q covers instances [q*gsize, (q+1)*gsize)
  int gsize;         // instances per group (= kNI of the instantiation: 8 or 16)
  int nst;           // instance stride of the chunk buffers (= gsize)
  int chunk_floats;  // floats per chunk buffer = (pr+1) * nst * mw
  int nbuf;          // chunk buffers in the ring (3 or 4)
  int zero_bytes;    // size of the shared-memory zero buffer (source of the bulk zero fills), a multiple of W
  int zero_rows;     // = zero_bytes / W
  int max_rows;      // generic scale: most dst rows owned by one prototype row pair (<= 12)
  int* work_counter;            // [0] global work-stealing counter, [1] finished CTAs (the last one resets both)
  unsigned long long* timing;   // developer diagnostic (VA_FUSED_TIMING=1): [grid][5 roles][8] cycle counters, or nullptr
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
  int i0;       // first instance of the item's group
  int n;        // instances of the group that exist in this frame (0 .. gsize)
  int pa, pb;   // row pairs [pa, pb): pair r blends proto rows r and min(r+1, mh-1) into dst rows 4r+2..4r+5
  int nrows;    // proto rows pa .. min(pb, mh-1)
  int npx;      // nrows * mw
  int ntiles;   // ceil(npx / 128)
  int nchunks;  // ceil((pb - pa) / pr)
  int band_pa, band_pb;   // the whole band; [pa, pb) is its live part (chunks that intersect the hull of the group's boxes)
};

// Work item index -> (band, frame, instance group), group fastest: the groups of one (frame, band) read the same
// prototype rows and are stolen by different CTAs at nearly the same time, so all but the first read hit L2.
// Items are handed out dynamically (work stealing through a global counter) in
// bottom-band-first order: the bands that contain the sidewalk blob cost more, so the light top bands form the
// tail of the schedule.  Out of line on purpose: called once per item per role, and inlined copies of its
// divisions would compete with the hot loops for the instruction cache.
__device__ __noinline__ Item decode_item(const FusedParams& p, int item, int c_lo, int c_hi, int n_known) {
  Item it;
  it.valid = item < p.n_items;
  const int fb = item / p.groups;
  const int q = item - fb * p.groups;
  const int jb = fb / p.B;
  const int b = fb - jb * p.B;
  const int j = p.nbands - 1 - jb;
  it.b = b;
  it.i0 = q * p.gsize;
  // instances of the group present in this frame: read once by the TMA warp, passed on through the item ring
  it.n = !it.valid ? 0 : (n_known >= 0) ? n_known : max(min(min(p.counts[b], p.d.max_n) - it.i0, p.gsize), 0);
  it.band_pa = j * p.ppb;
  it.band_pb = min(it.band_pa + p.ppb, p.d.mh);
  // live chunks [c_lo, c_hi) of the band (the TMA warp computes them from the boxes; (0, INT_MAX) = whole band)
  it.pa = min(it.band_pa + c_lo * p.pr, it.band_pb);
  it.pb = (c_hi >= p.ppb) ? it.band_pb : max(min(it.band_pa + c_hi * p.pr, it.band_pb), it.pa);
  if (it.pb > it.pa) {
    it.nrows = min(it.pb, p.d.mh - 1) - it.pa + 1;
    it.npx = it.nrows * p.d.mw;
    it.ntiles = ceil_div(it.npx, kTileM);
    it.nchunks = ceil_div(it.pb - it.pa, p.pr);
  } else {
    it.nrows = it.npx = it.ntiles = it.nchunks = 0;
  }
  return it;
}

// last band-local proto row that chunk c needs
__device__ __forceinline__ int chunk_last_row(const FusedParams& p, const Item& it, int c) {
  return min((c + 1) * p.pr, it.nrows - 1);
}

// Shared-memory map as byte OFFSETS from the 1024 B-aligned base; all hot accesses go through
// explicit ld.shared / st.shared on 32-bit shared addresses (no generic-space LD/ST).
struct SmemMap {
  uint32_t hi;        // [kStagesHi][16 KB]   TMA staging, [32 k][128 px]
  uint32_t bt;        // [2 parity][kNMma rows x 128 B]: rows [0,16) = coefficients hi, rows [16,32) = lo
  uint32_t chunks;    // [kChunkBufs][chunk_floats] f32
  uint32_t box;       // [2 item parity][kMaxInstTc][4] f32 (epilogue warps)
  uint32_t ubox;      // [kWarpsUp][kMaxInstTc][4] f32 (one private copy per upsample warp: no cross-warp barrier)
  uint32_t latpair;   // [mh] i16: the lattice dst row inside rows 4r+2..4r+5 of pair r, -1 if none
  uint32_t zeros;     // [(4*pr+2) * W] bytes of 0: source of the bulk zero-fill stores
  uint32_t zero_bytes;
  uint32_t stat;      // [kWarpsUp][kMaxInstTc][8] i32: area, minx, miny, maxx, maxy (private per upsample warp)
  uint32_t latrow;    // [H] i16; generic scale: [mh + 1] first dst row of every prototype row pair (ytab)
  uint32_t bars;      // [BAR_COUNT] u64
  uint32_t tmem_slot;
  uint32_t items;     // [kItemRing][4] i32
  uint32_t total;
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
  BAR_ITEM = BAR_CH_EMPTY + kChunkBufs,      // item ring: the TMA thread publishes the stolen item indices
  BAR_COUNT = BAR_ITEM + kItemRing
};

__host__ __device__ inline SmemMap fused_smem_map(int chunk_floats, int nbuf, int H, int mh, int zero_bytes) {
  SmemMap m;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes, uint32_t align) { o = (o + align - 1) / align * align; uint32_t r = o; o += bytes; return r; };
  m.hi = take(kStagesHi * kTileBytes, 1024);
  m.bt = take(2 * kNMma * 128, 1024);
  m.chunks = take((uint32_t)nbuf * chunk_floats * 4, 16);
  m.box = take(2 * kMaxInstTc * 4 * 4, 16);
  m.ubox = take(kWarpsUp * kMaxInstTc * 4 * 4, 16);
  m.latpair = take((uint32_t)(mh + 8) * 2, 16);
  m.zero_bytes = (uint32_t)zero_bytes;
  m.zeros = take((uint32_t)zero_bytes, 128);
  m.stat = take(kWarpsUp * kMaxInstTc * 8 * 4, 16);
  m.latrow = take((uint32_t)H * 2, 16);
  m.bars = take(BAR_COUNT * 8, 8);
  m.tmem_slot = take(16, 16);
  m.items = take(kItemRing * 16, 16);   // {item index, first live chunk, end of live chunks, instances}
  m.total = o;
  return m;
}

__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s16(uint32_t a) { short v; asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(a)); return (int)v; }
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_s16(uint32_t a, short v) { asm volatile("st.shared.s16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void atoms_add(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_min(uint32_t a, int v) { asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_max(uint32_t a, int v) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// barrier helpers on 32-bit shared addresses
__device__ __forceinline__ void bar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bar_commit(uint32_t a) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
template <int kN>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[kN]);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// cycle accounting of the mbarrier waits (only when p.timing != nullptr)
struct RoleTimer {
  unsigned long long* out;
  long long t_start, wait[6];
  __device__ __forceinline__ void begin(unsigned long long* o) { out = o; if (out) { t_start = clock64(); for (int i = 0; i < 6; ++i) wait[i] = 0; } }
  __device__ __forceinline__ void end() { if (out) { out[0] = (unsigned long long)(clock64() - t_start); for (int i = 0; i < 6; ++i) out[1 + i] = (unsigned long long)wait[i]; } }
};
#define TIMED_WAIT(tm, slot, bar, parity)                           \
  do {                                                               \
    const long long t0__ = (kDiag && (tm).out) ? clock64() : 0;      \
    bar_wait((bar), (parity));                                       \
    if (kDiag && (tm).out) (tm).wait[(slot)] += clock64() - t0__;    \
  } while (0)

template <int kN>
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[kN]);
template <>
__device__ __forceinline__ void tmem_ld_nowait<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_nowait<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// A box with a NaN coordinate keeps no pixel in the reference (every comparison of crop_mask, ops.py:688-704, is
// false).  All roles replace it by the empty box (+inf, +inf, -inf, -inf), for which every "outside" test of this
// kernel is true.  `v` is coordinate c (0: x1, 1: y1, 2: x2, 3: y2) of an instance whose four coordinates are held by
// four consecutive, 4-aligned lanes; all 32 lanes must call.
__device__ __forceinline__ float sanitize_box_coord(float v, int c) {
  const unsigned nan_lanes = __ballot_sync(0xffffffffu, v != v);
  const bool any = (nan_lanes >> ((threadIdx.x & 31) & ~3)) & 0xfu;
  return any ? ((c < 2) ? INFINITY : -INFINITY) : v;
}

__device__ __forceinline__ float trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// ---------------------------------------------------------------------------------------------
// the kernel.  kNI = instances handled per accumulator read (8 or 16 TMEM columns)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A operand read from tensor memory (one tf32 per 32-bit column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kGeneric = false: exact 4x geometry (H = 4 mh, W = 4 mw), constant blend weights.  kGeneric = true: any up-sampling
// scale (ATen's source index / lambda arithmetic per row and per pixel), dst rows assigned to prototype row pairs by a
// shared-memory table.
template <bool kWriteMasks, int kNI, bool kDiag, int kNBuf, bool kGeneric>
__global__ void __launch_bounds__(kThreads, 1)
fused_tc_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  // dynamic shared memory is only guaranteed 16 B aligned: round the shared address up to 1024
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const SmemMap sm = fused_smem_map(p.chunk_floats, kNBuf, p.d.H, p.d.mh, p.zero_bytes);
  const Dims& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = sbase + sm.bars;
  auto BAR = [&](int idx) { return bars + 8u * (uint32_t)idx; };
  // k-th item of this CTA as published by the TMA thread
  auto next_item = [&](int k) -> Item {
    bar_wait(BAR(BAR_ITEM + (k % kItemRing)), (k / kItemRing) & 1);
    const uint32_t e = sbase + sm.items + 16 * (k % kItemRing);
    return decode_item(p, lds_s32(e), lds_s32(e + 4), lds_s32(e + 8), lds_s32(e + 12));
  };

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesHi; ++i) { bar_init(BAR(BAR_HI_FULL + i), 1); bar_init(BAR(BAR_HI_EMPTY + i), kWarpsSplit); }
    for (int i = 0; i < kStagesLo; ++i) { bar_init(BAR(BAR_LO_FULL + i), kWarpsSplit); bar_init(BAR(BAR_LO_EMPTY + i), kIssuers); }
    for (int i = 0; i < kAcc; ++i) { bar_init(BAR(BAR_ACC_FULL + i), kIssuers); bar_init(BAR(BAR_ACC_EMPTY + i), kWarpsEpi); }
    for (int i = 0; i < 2; ++i) { bar_init(BAR(BAR_B_FULL + i), kWarpsSplit); bar_init(BAR(BAR_B_EMPTY + i), kIssuers); }
    for (int i = 0; i < kChunkBufs; ++i) { bar_init(BAR(BAR_CH_FULL + i), kWarpsEpi); bar_init(BAR(BAR_CH_EMPTY + i), kWarpsUp); }
    for (int i = 0; i < kItemRing; ++i) bar_init(BAR(BAR_ITEM + i), 1);
    fence_barrier_init();
  }
  for (int r = threadIdx.x; r < d.mh; r += kThreads) {   // lattice (cell-centre) dst row inside rows 4r+2 .. 4r+5
    const int half = d.gs >> 1;
    int ly = -1;
    for (int j = 0; j < 4; ++j) {
      const int t = 4 * r + 2 + j - half;
      if (t >= 0 && t % d.gs == 0 && 4 * r + 2 + j < d.H) ly = 4 * r + 2 + j;
    }
    sts_s16(sbase + sm.latpair + 2 * r, (short)ly);
  }
  if (kGeneric) {
    // ytab[r] = first dst row whose upper source row is r (ATen source index, monotone in Y, steps of at most 1 for
    // up-sampling scales); ytab[mh] = H.  Pair r owns dst rows ytab[r] .. ytab[r+1]-1.
    for (int Y = threadIdx.x; Y < d.H; Y += kThreads) {
      int y0, y1, q0, q1;
      float l0, l1;
      src_index(d.sy, Y, d.mh, y0, y1, l0, l1);
      if (Y == 0) { sts_s16(sbase + sm.latrow, 0); }
      else {
        src_index(d.sy, Y - 1, d.mh, q0, q1, l0, l1);
        if (q0 != y0) sts_s16(sbase + sm.latrow + 2 * y0, (short)Y);
      }
    }
    if (threadIdx.x == 0) sts_s16(sbase + sm.latrow + 2 * d.mh, (short)d.H);
  }
  for (uint32_t t = threadIdx.x * 16u; t < sm.zero_bytes; t += kThreads * 16u) sts_v4(sbase + sm.zeros + t, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_proxy_async();
  for (int t = threadIdx.x; t < kWarpsUp * kMaxInstTc * 8; t += kThreads) {
    const int f = t & 7;
    sts_s32(sbase + sm.stat + 4 * t, (f == 1 || f == 2) ? INT_MAX : (f == 3 || f == 4) ? -1 : 0);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + sm.tmem_slot), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds_u32(sbase + sm.tmem_slot);
  // Programmatic dependent launch: the tail kernel (launched with programmatic stream serialization) may become
  // resident on SMs whose CTA of this kernel has already exited; it waits in griddepcontrol.wait for the whole grid.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // =========================== work stealing + TMA producer ===========================
    // The whole warp fetches and sizes the next item; lane 0 then issues its tile loads.  (Fetching item k+1 while
    // the tiles of item k are being issued was measured slower: a CTA then holds a reserved item while others idle
    // at the end of the launch.)
    RoleTimer tm; tm.begin((kDiag && p.timing && lane == 0) ? p.timing + ((size_t)blockIdx.x * 5 + 1) * 8 : nullptr);
    uint32_t g = 0;
    for (int k = 0;; ++k) {
      Item it;
      int item = 0;
      do {                                                   // instance groups with no instance in this frame are dropped here
        if (lane == 0) item = atomicAdd(p.work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        it = decode_item(p, item, 0, INT_MAX, -1);
      } while (it.valid && it.n == 0);
      // Hull of the group's boxes in proto rows: chunks of the band that lie above / below every box produce only
      // zeros (crop_mask) - they are neither loaded nor multiplied; the upsample warps zero-fill their rows.
      int c_lo = 0, c_hi = it.nchunks;
      if (it.valid && !(kDiag && p.logits_dbg)) {
        float y1 = INFINITY, y2 = -INFINITY;
        if (lane < it.n) {
          const float* bx = p.boxes + ((size_t)it.b * d.max_n + it.i0 + lane) * 4;
          const float4 q = __ldg(reinterpret_cast<const float4*>(bx));
          y1 = __fmul_rn(q.y, d.hr);
          y2 = __fmul_rn(q.w, d.hr);
          if (q.x != q.x || q.y != q.y || q.z != q.z || q.w != q.w) { y1 = INFINITY; y2 = -INFINITY; }   // NaN box = empty box
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          y1 = fminf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
          y2 = fmaxf(y2, __shfl_xor_sync(0xffffffffu, y2, o));
        }
        // same test as the upsample warps' per-instance chunk test, against the hull
        while (c_lo < c_hi && (float)min(it.pa + (c_lo + 1) * p.pr, it.pb) < y1) ++c_lo;
        while (c_hi > c_lo && (float)(it.pa + (c_hi - 1) * p.pr) >= y2) --c_hi;
        it = decode_item(p, item, c_lo, c_hi, it.n);
      }
      if (lane == 0) {
        const uint32_t e = sbase + sm.items + 16 * (k % kItemRing);
        sts_s32(e, item); sts_s32(e + 4, c_lo); sts_s32(e + 8, c_hi); sts_s32(e + 12, it.n);
        bar_arrive(BAR(BAR_ITEM + (k % kItemRing)));          // release: the entry is visible to the waiting roles
      }
      if (!it.valid) break;
      if (lane == 0) {
        const int px0 = it.pa * d.mw;
#pragma unroll 1
        for (int t = 0; t < it.ntiles; ++t) {
          const int st = (g + t) % kStagesHi;
          TIMED_WAIT(tm, 0, BAR(BAR_HI_EMPTY + st), (((g + t) / kStagesHi) & 1) ^ 1);
          bar_expect_tx(BAR(BAR_HI_FULL + st), kTileBytes);
          tma_load_3d_a(sbase + sm.hi + st * kTileBytes, &tmap, BAR(BAR_HI_FULL + st), px0 + t * kTileM, 0, it.b);
        }
      }
      g += it.ntiles;
      __syncwarp();
    }
    tm.end();
  } else if (warp < kFirstEpiWarp) {
    // =========================== 3xTF32 split: staged [k][px] box -> registers -> tensor memory ===========================
    const int quarter = warp & 3;                    // TMEM lanes [32*quarter, +32) belong to this warp
    const int st_tid = (warp - kFirstSplitWarp) * 32 + lane;
    const int px = quarter * 32 + lane;              // tile pixel (= TMEM lane) of this thread
    const int issuer = (warp - kFirstSplitWarp) < kIssuers ? (warp - kFirstSplitWarp) : -1;
    const uint32_t idesc = make_idesc(kTileM, kNMma);
    RoleTimer tm; tm.begin((kDiag && p.timing && st_tid == 0) ? p.timing + ((size_t)blockIdx.x * 5 + 2) * 8 : nullptr);
    uint32_t g = 0;
    int kk = 0;                                      // items that have tiles (parity of the B tile / box double buffers)
    for (int k = 0;; ++k) {
      const Item it = next_item(k);
      if (!it.valid) break;
      if (it.ntiles == 0) continue;                  // no live chunk: only the upsample warps act (zero fill)
      {   // per-frame B tiles: coefficients hi / lo, K-major, 16 B chunk c of row r stored at chunk c ^ (r & 7)
        const int par = kk & 1;
        TIMED_WAIT(tm, 0, BAR(BAR_B_EMPTY + par), ((kk >> 1) & 1) ^ 1);
        const int n = it.n;
        const uint32_t bh = sbase + sm.bt + par * kNMma * 128;   // rows [0,16): hi
        const uint32_t bl = bh + kNPad * 128;                      // rows [16,32): lo (16 = 2 swizzle periods: same XOR pattern)
        const int r = st_tid >> 3, c = st_tid & 7;   // 128 threads = 16 rows x 8 chunks
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) v = __ldg(reinterpret_cast<const float4*>(p.coefs + ((size_t)it.b * d.max_n + it.i0 + r) * d.K + 4 * c));
        float4 h, l;
        h.x = trunc_tf32(v.x); l.x = v.x - h.x;
        h.y = trunc_tf32(v.y); l.y = v.y - h.y;
        h.z = trunc_tf32(v.z); l.z = v.z - h.z;
        h.w = trunc_tf32(v.w); l.w = v.w - h.w;
        const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
        sts_v4(bh + off, h);
        sts_v4(bl + off, l);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) bar_arrive(BAR(BAR_B_FULL + par));
      }
      for (int t = 0; t < it.ntiles; ++t, ++g) {
        const int sh = g % kStagesHi, sl = g % kStagesLo;
        TIMED_WAIT(tm, 1, BAR(BAR_HI_FULL + sh), (g / kStagesHi) & 1);
        const uint32_t src = sbase + sm.hi + sh * kTileBytes + px * 4;
        uint32_t hi[kProtoK], lo[kProtoK];
#pragma unroll
        for (int q = 0; q < kProtoK; ++q) hi[q] = __float_as_uint(lds_f32(src + q * kTileM * 4));   // conflict-free: consecutive px
        __syncwarp();
        if (lane == 0) bar_arrive(BAR(BAR_HI_EMPTY + sh));      // staging slot can be refilled
        // the tensor core reads the top 19 bits of each word (truncation, measured): the raw value IS the hi
        // operand; lo = x - trunc(x) is exact in fp32.
#pragma unroll
        for (int q = 0; q < kProtoK; ++q) lo[q] = __float_as_uint(__uint_as_float(hi[q]) - __uint_as_float(hi[q] & 0xffffe000u));
        TIMED_WAIT(tm, 2, BAR(BAR_LO_EMPTY + sl), ((g / kStagesLo) & 1) ^ 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + kTmemAOff + sl * 64;
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(BAR(BAR_LO_FULL + sl));
        // ---- MMA issue.  tcgen05.mma instructions issued by DIFFERENT warps overlap (measured: ~70 cycles per
        // small MMA per issuing thread, 2 issuers -> 2x the rate), so the two passes of a tile are issued by
        // lane 0 of split warps 0 and 1 into separate accumulators; the epilogue adds them.
        //   issuer 0: D0[:, 0:16 | 16:32] = A_hi * [B_hi | B_lo]^T      issuer 1: D1 = A_lo * [B_hi | B_lo]^T
        if (issuer >= 0 && lane == 0) {
          const int par = kk & 1;
          if (t == 0) TIMED_WAIT(tm, 3, BAR(BAR_B_FULL + par), (kk >> 1) & 1);
          const int ac = g % kAcc;
          TIMED_WAIT(tm, 4, BAR(BAR_LO_FULL + sl), (g / kStagesLo) & 1);       // all four lane quarters of A are in TMEM
          TIMED_WAIT(tm, 5, BAR(BAR_ACC_EMPTY + ac), ((g / kAcc) & 1) ^ 1);
          tc_fence_after();
          const uint32_t b_all = sbase + sm.bt + par * kNMma * 128;
          const uint32_t a_op = tmem_base + kTmemAOff + sl * 64 + issuer * 32;     // A in tensor memory: lane = pixel, column = k
          const uint32_t dcol = tmem_base + kTmemAccOff + ac * kAccCols + issuer * kNMma;
          // B K-major SW128 in shared memory: rows are 128 B, 8-row groups 1024 B apart (SBO), one k-step
          // (8 tf32) = 32 B along the row = 8 TMEM columns of A.
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(dcol, a_op + ks * 8, make_smem_desc(b_all + ks * 32, 16, 1024), idesc, ks > 0);
          bar_commit(BAR(BAR_LO_EMPTY + sl));
          bar_commit(BAR(BAR_ACC_FULL + ac));
          if (t + 1 == it.ntiles) bar_commit(BAR(BAR_B_EMPTY + par));
        }
        __syncwarp();
      }
      ++kk;
    }
    tm.end();
  } else if (warp < kFirstUpWarp) {
    // =========================== epilogue: TMEM -> crop -> chunk buffers ===========================
    const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are accessible to this warp
    const int ep_px = quarter * 32 + lane;   // tile pixel this thread reads back
    const int ep_tid = (warp - kFirstEpiWarp) * 32 + lane;
    const int tile_rows = kTileM / d.mw, tile_cols = kTileM - tile_rows * d.mw;   // (row, col) advance per tile
    const uint32_t inst_stride = (uint32_t)d.mw * 4;
    const uint32_t row_stride = (uint32_t)p.nst * inst_stride;
    const uint32_t chunks = sbase + sm.chunks;
    RoleTimer tm; tm.begin((kDiag && p.timing && ep_tid == 0) ? p.timing + ((size_t)blockIdx.x * 5 + 3) * 8 : nullptr);
    uint32_t g = 0, chunk_base = 0;
    int kk = 0;
    for (int k = 0;; ++k) {
      const Item it = next_item(k);
      if (!it.valid) break;
      if (it.ntiles == 0) continue;
      const int n = it.n;
      // scaled boxes of this frame (double-buffered by item parity; the 4 epilogue warps stay within one item)
      const uint32_t bx = sbase + sm.box + (kk & 1) * kMaxInstTc * 16;
      ++kk;
      if (ep_tid < kMaxInstTc * 4) {
        const int i = ep_tid >> 2, c = ep_tid & 3;
        float v = 0.f;
        if (i < n) v = __fmul_rn(__ldg(p.boxes + ((size_t)it.b * d.max_n + it.i0 + i) * 4 + c), (c & 1) ? d.hr : d.wr);
        sts_f32(bx + (i * 4 + c) * 4, sanitize_box_coord(v, c));
      }
      named_bar_sync(1, 32 * kWarpsEpi);
      // crop_mask (ops.py:688-704) keeps proto pixel (col,row) iff col >= x1 && col < x2 && row >= y1 && row < y2 with
      // FLOAT box bounds (x1.. scaled as ops.py:725-732).  col/row are integers, so this is exactly
      // col in [ceil(x1), ceil(x2)) and row in [ceil(y1), ceil(y2)): integer bounds per instance, kept in
      // registers for the whole item; one unsigned compare per axis.  Instances >= n have an empty box.
      int cl[kNI], rl[kNI];
      unsigned cwid[kNI], rwid[kNI];
#pragma unroll
      for (int i = 0; i < kNI; ++i) {
        const float4 q = lds_v4(bx + 16 * i);
        const int x1 = (int)ceilf(fminf(fmaxf(q.x, -1e6f), 1e6f)), x2 = (int)ceilf(fminf(fmaxf(q.z, -1e6f), 1e6f));
        const int y1 = (int)ceilf(fminf(fmaxf(q.y, -1e6f), 1e6f)), y2 = (int)ceilf(fminf(fmaxf(q.w, -1e6f), 1e6f));
        cl[i] = x1; cwid[i] = (unsigned)max(x2 - x1, 0);
        rl[i] = (i < n) ? y1 : (INT_MAX >> 1); rwid[i] = (i < n) ? (unsigned)max(y2 - y1, 0) : 0u;
      }
      // lane i also keeps the row range of instance i: one ballot per tile gives the instances worth storing
      int my_rl = INT_MAX >> 1, my_rend = INT_MAX >> 1;
#pragma unroll
      for (int i = 0; i < kNI; ++i)
        if (lane == i) { my_rl = rl[i]; my_rend = rl[i] + (int)rwid[i]; }
      int brow = ep_px / d.mw, bcol = ep_px - brow * d.mw;   // band-local (row, col) of this thread's pixel
      int done_rows = 0, done_cols = 0;                        // complete rows / extra pixels after the current tile
      int acquired = 0, completed = 0;
      for (int t = 0; t < it.ntiles; ++t, ++g) {
        const int ac = g % kAcc;
        TIMED_WAIT(tm, 0, BAR(BAR_ACC_FULL + ac), (g / kAcc) & 1);
        __syncwarp();
        tc_fence_after();
        uint32_t r[kNI], r2[kNI];
        const long long t_ld0 = (kDiag && tm.out) ? clock64() : 0;
        {
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + kTmemAccOff + ac * kAccCols;
          uint32_t r3[kNI], r4[kNI];
          tmem_ld_nowait<kNI>(ta, r);                       // A_hi * B_hi
          tmem_ld_nowait<kNI>(ta + kNPad, r2);              // A_hi * B_lo
          tmem_ld_nowait<kNI>(ta + kNMma, r3);              // A_lo * B_hi
          tmem_ld_nowait<kNI>(ta + kNMma + kNPad, r4);      // A_lo * B_lo
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < kNI; ++i)
            r[i] = __float_as_uint((__uint_as_float(r[i]) + __uint_as_float(r3[i])) + (__uint_as_float(r2[i]) + __uint_as_float(r4[i])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(BAR(BAR_ACC_EMPTY + ac));
        if (kDiag && tm.out) tm.wait[2] += clock64() - t_ld0;

        const int tile_ra = it.pa + done_rows;                 // first proto row this tile touches
        done_rows += tile_rows; done_cols += tile_cols;
        if (done_cols >= d.mw) { done_cols -= d.mw; ++done_rows; }
        const bool last_tile = (t + 1 == it.ntiles);
        const int rows_done = last_tile ? it.nrows : done_rows;
        const int row_last = last_tile ? it.nrows - 1 : (done_cols == 0 ? done_rows - 1 : done_rows);
        const int c_hi = min(row_last >> p.pr_shift, it.nchunks - 1);
        const int tile_rb = it.pa + row_last;                  // last proto row this tile touches
        // A stored row rho is read by the upsample tasks of pairs rho-1 and rho, and a task whose two rows are
        // outside the instance's box rows (r+1 < y1 or r >= y2) never reads its chunk rows.  So when every row of
        // this tile is at least 2 above or 1 below the box (or the instance does not exist), nothing will read what
        // would be stored: such instances are skipped (warp-uniform mask).
        const unsigned store_mask = (kDiag && p.logits_dbg) ? 0xffffffffu
                                    : __ballot_sync(0xffffffffu, !(tile_rb + 2 <= my_rl || tile_ra - 1 >= my_rend));
#pragma unroll 1
        while (acquired <= c_hi) {            // acquire the chunk buffers this tile writes, in order
          const uint32_t gc = chunk_base + acquired;
          TIMED_WAIT(tm, 1, BAR(BAR_CH_EMPTY + gc % kNBuf), ((gc / kNBuf) & 1) ^ 1);
          ++acquired;
        }
        if (brow < it.nrows) {
          const int grow = it.pa + brow;
          const int c1 = brow >> p.pr_shift, rr = brow - (c1 << p.pr_shift);
          const bool has1 = c1 < it.nchunks, has0 = (rr == 0 && c1 > 0);
          const uint32_t dst1 = chunks + ((chunk_base + c1) % kNBuf) * (uint32_t)p.chunk_floats * 4 + rr * row_stride + bcol * 4;
          const uint32_t dst0 = chunks + ((chunk_base + c1 + kNBuf - 1) % kNBuf) * (uint32_t)p.chunk_floats * 4 + p.pr * row_stride + bcol * 4;
          float* dbg = (kDiag && p.logits_dbg) ? p.logits_dbg + (((size_t)it.b * d.max_n + it.i0) * d.mh + grow) * d.mw + bcol : nullptr;
          const size_t dbg_stride = (size_t)d.mh * d.mw;
#pragma unroll
          for (int i = 0; i < kNI; ++i) {
            if (!((store_mask >> i) & 1u)) continue;
            const bool keep = ((unsigned)(bcol - cl[i]) < cwid[i]) && ((unsigned)(grow - rl[i]) < rwid[i]);
            const float v = keep ? __uint_as_float(r[i]) : 0.f;
            if (has1) sts_f32(dst1 + i * inst_stride, v);
            if (has0) sts_f32(dst0 + i * inst_stride, v);
            if (kDiag && dbg && i < n) dbg[i * dbg_stride] = v;
          }
        }
        __syncwarp();
#pragma unroll 1
        while (completed < it.nchunks && chunk_last_row(p, it, completed) < rows_done) {   // chunks completed by this tile
          const uint32_t gc = chunk_base + completed;
          if (lane == 0) bar_arrive(BAR(BAR_CH_FULL + gc % kNBuf));
          ++completed;
        }
        brow += tile_rows; bcol += tile_cols;
        if (bcol >= d.mw) { bcol -= d.mw; ++brow; }
      }
      chunk_base += it.nchunks;
    }
    tm.end();
  } else {
    // =========================== upsample + threshold + store + reductions ===========================
    const int uw = warp - kFirstUpWarp;                      // 0..kWarpsUp-1
    const int ut = threadIdx.x - 32 * kFirstUpWarp;
    const int NG = d.W >> 4, NG8 = ceil_div(NG, 8);
    // lane -> (8 column groups) x (4 slots); slot -> (pair within the chunk, sub-block of 8 groups)
    const int gl = lane & 7, slot = lane >> 3;
    const int subs = 4 >> p.pr_shift;                        // pr = 4: 1, pr = 2: 2
    const int pair = slot & (p.pr - 1), sub = slot >> p.pr_shift;
    const int ng8w = ceil_div(NG8, subs);                    // warp tasks per instance
    const int ng8w_mod = ng8w % kWarpsUp;
    const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    const uint4 zeros = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t inst_stride = (uint32_t)d.mw * 4;
    const uint32_t row_stride = (uint32_t)p.nst * inst_stride;
    // first dst row of prototype row pair r (pair r owns dst rows YT(r) .. YT(r+1)-1; YT(mh) = H)
    auto YT = [&](int r) -> int {
      if (kGeneric) return lds_s16(sbase + sm.latrow + 2 * r);
      return (r == 0) ? 0 : min(4 * r + 2, d.H);
    };
    // bulk zero fill of dst rows [Ya, Yb] of one instance, in pieces of at most zero_rows rows
    auto zero_rows_bulk = [&](uint8_t* inst_base, int Ya, int Yb) {
      for (int y = Ya; y <= Yb; y += p.zero_rows) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(inst_base + (size_t)y * d.W),
                     "r"(sbase + sm.zeros), "r"((uint32_t)(min(p.zero_rows, Yb + 1 - y) * d.W))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    RoleTimer tm; tm.begin((kDiag && p.timing && ut == 0) ? p.timing + ((size_t)blockIdx.x * 5 + 4) * 8 : nullptr);
    uint32_t gc = 0;
    int base = 0;                        // (live task count) % kWarpsUp
    int zw = 0;                          // warp that issues the next bulk zero-fill
    for (int k = 0;; ++k) {
      const Item it = next_item(k);
      if (!it.valid) break;
      const int n = it.n;
      // scaled boxes of this frame for the outside-the-box tests: a private copy per warp, so the upsample warps
      // never wait for each other
      const uint32_t ubox = sbase + sm.ubox + uw * kMaxInstTc * 16;
      const uint32_t wstat = sbase + sm.stat + uw * kMaxInstTc * 32;
      for (int q4 = lane; q4 < kMaxInstTc * 4; q4 += 32) {
        const int i = q4 >> 2, c = q4 & 3;
        float v = 0.f;
        if (i < n) v = __fmul_rn(__ldg(p.boxes + ((size_t)it.b * d.max_n + it.i0 + i) * 4 + c), (c & 1) ? d.hr : d.wr);
        sts_f32(ubox + q4 * 4, sanitize_box_coord(v, c));
      }
      __syncwarp();
      // Live chunk range [ci_lo, ci_hi) of every instance (lane = instance): a chunk whose proto rows r0 .. r0+npairs
      // all lie outside the instance's box rows was zeroed by crop_mask, and because the test is monotonic in r0
      // the dead chunks are a prefix and a suffix of the item.  The chunk loop below only visits live (chunk,
      // instance) pairs; dead ones get one bulk zero-fill of the chunk's dst rows (issued when the chunk comes up, so
      // that the copies are spread over the item and do not queue up in front of the prototype loads).
      int ci_lo = 0, ci_hi = 0;
      if (lane < n) {
        const float y1 = lds_f32(ubox + 16 * lane + 4), y2 = lds_f32(ubox + 16 * lane + 12);
        ci_hi = it.nchunks;
        while (ci_lo < ci_hi && (float)min(it.pa + (ci_lo + 1) * p.pr, it.pb) < y1) ++ci_lo;
        while (ci_hi > ci_lo && (float)(it.pa + (ci_hi - 1) * p.pr) >= y2) --ci_hi;
      }
      if (kWriteMasks) {
        // Parts of the band the hull of the boxes excluded (no chunk exists for them): zero for every instance.
        // Pieces of <= zero_rows rows, one per lane; the instances go round the upsample warps.
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          const int A = part ? it.pb : it.band_pa, Bp = part ? it.band_pb : it.pa;
          if (A >= Bp) continue;
          const int Ya = YT(A);
          const int Yb = YT(Bp) - 1;
#pragma unroll 1
          for (int i = 0; i < n; ++i) {
            if (zw == uw) {
              uint8_t* base = p.masks + ((size_t)it.b * d.max_n + it.i0 + i) * (size_t)d.H * d.W;
              for (int y = Ya + lane * p.zero_rows; y <= Yb; y += 32 * p.zero_rows) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)y * d.W),
                             "r"(sbase + sm.zeros), "r"((uint32_t)(min(p.zero_rows, Yb + 1 - y) * d.W))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
            zw = (zw + 1 == kWarpsUp) ? 0 : zw + 1;
          }
        }
      }
      for (int c = 0; c < it.nchunks; ++c, ++gc) {
        const int buf = gc % kNBuf;
        TIMED_WAIT(tm, 0, BAR(BAR_CH_FULL + buf), (gc / kNBuf) & 1);
        const uint32_t cb = sbase + sm.chunks + buf * (uint32_t)p.chunk_floats * 4;
        const int r0 = it.pa + c * p.pr;                        // first pair of the chunk
        const int npairs = min(p.pr, it.pb - r0);
        const int r = r0 + pair;
        const bool last = (r == d.mh - 1);
        // dst rows of this lane's pair: Yfirst .. Yfirst + nrows - 1 (exact 4x: 4, 6 for pair 0, 2 for the last pair)
        const int Yfirst = (pair < npairs) ? YT(r) : 0;
        const int nrows = (pair < npairs) ? YT(r + 1) - Yfirst : 0;
        const int laty = (!kGeneric && pair < npairs) ? lds_s16(sbase + sm.latpair + 2 * r) : -1;
        // Warp tasks: (live instance, block of 8*subs column groups), numbered consecutively and dealt round-robin:
        // task wq = live_ordinal * ng8w + g8w belongs to warp wq % kWarpsUp.  The numbering runs on across chunks
        // and items, so the warps that get one task more than the others rotate.
        unsigned live = __ballot_sync(0xffffffffu, lane < n && ci_lo <= c && c < ci_hi);
        if (kWriteMasks) {
          // dead instances of this chunk: instance i is zero-filled by warp (i + c) % kWarpsUp (bits i, i+7, ...)
          int kq = (uw - c) % kWarpsUp;
          if (kq < 0) kq += kWarpsUp;
          unsigned dead = ~live & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u)) & (0x10204081u << kq);
          if (lane == 0) {
            const int Ya = YT(r0);
            const int Yb = YT(r0 + npairs) - 1;
            while (dead) {
              const int i = __ffs(dead) - 1;
              dead &= dead - 1;
              zero_rows_bulk(p.masks + ((size_t)it.b * d.max_n + it.i0 + i) * (size_t)d.H * d.W, Ya, Yb);
            }
          }
          __syncwarp();
        }
#pragma unroll 1
        while (live) {
          const int i = __ffs(live) - 1;
          live &= live - 1;
          const float4 q = lds_v4(ubox + 16 * i);     // x1, y1, x2, y2 at proto resolution
          int g8w = uw - base;
          if (g8w < 0) g8w += kWarpsUp;
          base += ng8w_mod;
          if (base >= kWarpsUp) base -= kWarpsUp;
#pragma unroll 1
          for (; g8w < ng8w; g8w += kWarpsUp) {
          const int g = (g8w * subs + sub) * 8 + gl;
          const bool active = (g < NG) && (pair < npairs);
          const size_t inst = (size_t)it.b * d.max_n + it.i0 + i;
          RowPats<kGeneric> rp;                         // this lane's 16-pixel patterns of its dst rows (slot = row - Yfirst)
          bool touched = false;                         // rp may hold a set pixel
          if (active) {
            uint8_t* M = kWriteMasks ? p.masks + inst * (size_t)d.H * d.W + 16 * g : nullptr;
            unsigned* lat = p.lattice + inst * (size_t)d.lat_rows * d.lat_words;
            const uint32_t rowA = cb + pair * row_stride + i * inst_stride;
            const uint32_t rowB = last ? rowA : rowA + row_stride;
            if (!kGeneric) {
              // every proto pixel this task reads (rows r, r+1, cols 4g-1 .. 4g+4) is outside the instance's box:
              // crop_mask zeroed them, the masks are 0 - store and skip everything else
              const bool outside = ((float)(r + 1) < q.y) || ((float)r >= q.w) || ((float)(4 * g + 4) < q.x) || ((float)(4 * g - 1) >= q.z);
              if (outside) {
                if (kWriteMasks) {
#pragma unroll 1
                  for (int j = 0; j < nrows; ++j) *reinterpret_cast<uint4*>(M + (size_t)(Yfirst + j) * d.W) = zeros;
                }
              } else {
                const uint32_t pA = rowA + g * 16, pB = rowB + g * 16;
                float sA[6], sB[6];
                {
                  const float4 v = lds_v4(pA);
                  sA[1] = v.x; sA[2] = v.y; sA[3] = v.z; sA[4] = v.w;
                  sA[0] = (g > 0) ? lds_f32(pA - 4) : v.x;
                  sA[5] = (4 * g + 4 < d.mw) ? lds_f32(pA + 16) : v.w;
                  const float4 u = lds_v4(pB);
                  sB[1] = u.x; sB[2] = u.y; sB[3] = u.z; sB[4] = u.w;
                  sB[0] = (g > 0) ? lds_f32(pB - 4) : u.x;
                  sB[5] = (4 * g + 4 < d.mw) ? lds_f32(pB + 16) : u.w;
                }
                const float mnA = fminf(fminf(fminf(sA[0], sA[1]), fminf(sA[2], sA[3])), fminf(sA[4], sA[5]));
                const float mxA = fmaxf(fmaxf(fmaxf(sA[0], sA[1]), fmaxf(sA[2], sA[3])), fmaxf(sA[4], sA[5]));
                const float mnB = fminf(fminf(fminf(sB[0], sB[1]), fminf(sB[2], sB[3])), fminf(sB[4], sB[5]));
                const float mxB = fmaxf(fmaxf(fmaxf(sB[0], sB[1]), fmaxf(sB[2], sB[3])), fmaxf(sB[4], sB[5]));
                const bool left = (g == 0);
                const bool uniA_pos = mnA > kTiny, uniA_neg = mxA <= 0.f;
                const bool uni_pos = uniA_pos && (mnB > kTiny), uni_neg = uniA_neg && (mxB <= 0.f);
                if (uni_neg || uni_pos) {
                  if (kWriteMasks) {
                    const uint4 w = uni_pos ? ones : zeros;
#pragma unroll 1
                    for (int j = 0; j < nrows; ++j) *reinterpret_cast<uint4*>(M + (size_t)(Yfirst + j) * d.W) = w;
                  }
                  if (uni_pos) {
                    rp.fill(nrows);
                    touched = true;
                    if (laty >= 0 && laty < Yfirst + nrows) lattice_row(ones, laty, 16 * g, d, lat);
                  }
                } else {
                  // mixed signs: the exact 4-tap blend.  One compact, rolled loop (code size matters: the roles
                  // share the instruction cache).
                  touched = true;
                  float hA[16], hB[16];
                  hinterp4(sA, hA, left);
                  hinterp4(sB, hB, left);
                  const int jbeg = (r == 0) ? -2 : 0;     // pair 0 also owns dst rows 0,1: h(row 0) unchanged (src y clamps to 0)
#pragma unroll 1
                  for (int j = jbeg; j < nrows + jbeg; ++j) {
                    uint4 w;
                    const int Y = Yfirst + j - jbeg;
                    if (j < 0) {
                      w = hpack(hA);
                    } else {
                      const float l1 = 0.125f + 0.25f * (float)j;     // .125 .375 .625 .875 (exact)
                      w = vblend(hA, hB, 1.0f - l1, l1);
                    }
                    if (kWriteMasks) *reinterpret_cast<uint4*>(M + (size_t)Y * d.W) = w;
                    rp.set(j - jbeg, pat16(w));
                    if (Y == laty && (w.x | w.y | w.z | w.w)) lattice_row(w, Y, 16 * g, d, lat);
                  }
                }
              }
            } else {
              // ---- generic scale: the 16 pixels read proto columns xa .. xb of rows r, r+1 ----
              const int X0 = 16 * g;
              int xa, xb, tq;
              float f0, f1;
              src_index(d.sx, X0, d.mw, xa, tq, f0, f1);
              src_index(d.sx, X0 + 15, d.mw, tq, xb, f0, f1);
              const bool outside = ((float)(r + 1) < q.y) || ((float)r >= q.w) || ((float)xb < q.x) || ((float)xa >= q.z);
              if (outside) {
                if (kWriteMasks) {
#pragma unroll 1
                  for (int j = 0; j < nrows; ++j) *reinterpret_cast<uint4*>(M + (size_t)(Yfirst + j) * d.W) = zeros;
                }
              } else {
                float mn = INFINITY, mx = -INFINITY;
#pragma unroll 1
                for (int x = xa; x <= xb; ++x) {
                  const float a = lds_f32(rowA + 4 * x), bq = lds_f32(rowB + 4 * x);
                  mn = fminf(mn, fminf(a, bq));
                  mx = fmaxf(mx, fmaxf(a, bq));
                }
                const bool uni_pos = mn > kTiny, uni_neg = mx <= 0.f;
                const int half = d.gs >> 1;
                if (uni_neg || uni_pos) {
                  if (kWriteMasks) {
                    const uint4 w = uni_pos ? ones : zeros;
#pragma unroll 1
                    for (int j = 0; j < nrows; ++j) *reinterpret_cast<uint4*>(M + (size_t)(Yfirst + j) * d.W) = w;
                  }
                  if (uni_pos) {
                    rp.fill(nrows);
                    touched = true;
#pragma unroll 1
                    for (int j = 0; j < nrows; ++j) {
                      const int tl = Yfirst + j - half;
                      if (tl >= 0 && tl % d.gs == 0) lattice_row(ones, Yfirst + j, X0, d, lat);
                    }
                  }
                } else {
                  // horizontal pass for both proto rows (ATen: top = fma(a, l0x, fl(b * l1x))), then one vertical
                  // blend per dst row of the pair (out = fma(top, l0y, fl(bot * l1y)))
                  touched = true;
                  float hA[16], hB[16];
#pragma unroll
                  for (int px = 0; px < 16; ++px) {
                    int x0, x1;
                    float lx0, lx1;
                    src_index(d.sx, X0 + px, d.mw, x0, x1, lx0, lx1);
                    hA[px] = fmaf(lds_f32(rowA + 4 * x0), lx0, __fmul_rn(lds_f32(rowA + 4 * x1), lx1));
                    hB[px] = fmaf(lds_f32(rowB + 4 * x0), lx0, __fmul_rn(lds_f32(rowB + 4 * x1), lx1));
                  }
#pragma unroll 1
                  for (int j = 0; j < nrows; ++j) {
                    const int Y = Yfirst + j;
                    int y0, y1;
                    float ly0, ly1;
                    src_index(d.sy, Y, d.mh, y0, y1, ly0, ly1);
                    const uint4 w = vblend(hA, hB, ly0, ly1);
                    if (kWriteMasks) *reinterpret_cast<uint4*>(M + (size_t)Y * d.W) = w;
                    rp.set(j, pat16(w));
                    const int tl = Y - half;
                    if ((w.x | w.y | w.z | w.w) && tl >= 0 && tl % d.gs == 0) lattice_row(w, Y, X0, d, lat);
                  }
                }
              }
            }
          }
          // Per-row by-products for the contour step: the 8 lanes of a slot hold one 128-pixel block of the same dst
          // rows -> one 4-byte summary per (row, block), written by the group leader, which also carries the
          // instance's area / bbox.  Most tasks lie entirely inside or outside the mask: two votes recognise blocks
          // that are all ones / all zeros and their summaries are constants; only tasks that cross the outline
          // gather the 128-bit block patterns (four shuffles per row, the rows' chains interleaved).
          __syncwarp();
          if (p.rowsum != nullptr && !kWriteMasks) {
            const int nrow = active ? nrows : 0;
#pragma unroll 1
            for (int sidx = 0; sidx < nrow; ++sidx)
              p.bits16[(inst * d.H + Yfirst + sidx) * (size_t)(2 * d.bit_words) + g] = (uint16_t)rp.get(sidx);
          }
          // a task with no set pixel at all (most of them: outside the box or below the threshold) leaves the summaries
          // in their resting state - one vote instead of the classification below
          if (p.rowsum != nullptr && __any_sync(0xffffffffu, touched)) {   // nullptr: developer A/B switch (VA_NO_ROWSUM=1), records are then meaningless
            const int nrow = active ? nrows : 0;
            const bool is_zero = rp.is_zero();
            const bool is_full = active && rp.is_full(nrow);
            const unsigned zm = __ballot_sync(0xffffffffu, is_zero), fm = __ballot_sync(0xffffffffu, is_full || !active);
            // per 8-lane group: every lane zero, or every lane full (lanes past the row end count as full: they only
            // exist in the last block of a ragged row, where the exact path below is taken instead)
            const unsigned gz = zm & (zm >> 4), gf = fm & (fm >> 4);
            const unsigned gz2 = gz & (gz >> 2), gf2 = gf & (gf >> 2);
            const unsigned gzall = gz2 & (gz2 >> 1) & 0x01010101u, gfall = gf2 & (gf2 >> 1) & 0x01010101u;   // bit 8q: group q
            const bool ragged = (NG & 7) != 0;
            uint32_t* rs_inst = p.rowsum + inst * (size_t)d.H * d.nblk;
            LeaderStats ls;
            if (!ragged && ((gzall | gfall) == 0x01010101u)) {
              if (gl == 0 && nrow > 0 && ((gfall >> (lane & 24)) & 1u)) {     // all-zero blocks: nothing to store (resting state)
                const int blk = g >> 3;
                const uint32_t e = cc::rowsum_pack(cc::kRowBlock, 0, cc::kRowBlock - 1);
#pragma unroll 1
                for (int sidx = 0; sidx < nrow; ++sidx) rs_inst[(size_t)(Yfirst + sidx) * d.nblk + blk] = e;
                ls.area = (unsigned)(cc::kRowBlock * nrow);
                ls.minx = cc::kRowBlock * blk; ls.maxx = cc::kRowBlock * blk + cc::kRowBlock - 1;
                ls.miny = Yfirst; ls.maxy = Yfirst + nrow - 1;
              }
            } else {
              // rolled on purpose (code size: the roles share the instruction cache); exact 4x: pair 0 of the frame
              // owns 6 rows, generic scale: up to max_rows (both bounds warp-uniform)
              const int srows = kGeneric ? p.max_rows : ((r0 == 0) ? 6 : 4);
#pragma unroll 1
              for (int sidx = 0; sidx < srows; ++sidx)
                emit_row_summary(rp.get(sidx), gl, sidx < nrow, Yfirst + sidx, g >> 3, rs_inst, d.nblk, ls);
            }
            if (gl == 0 && ls.area) {
              const uint32_t st = wstat + i * 32;
              atoms_add(st, (int)ls.area);
              atoms_min(st + 4, ls.minx);
              atoms_min(st + 8, ls.miny);
              atoms_max(st + 12, ls.maxx);
              atoms_max(st + 16, ls.maxy);
            }
          }
          }   // tasks of instance i
        }     // instances
        __syncwarp();
        if (lane == 0) bar_arrive(BAR(BAR_CH_EMPTY + buf));
      }
      // ---- item done: publish this warp's reductions for frame it.b ----
      __syncwarp();
      if (lane < n) {
        const uint32_t st = wstat + lane * 32;
        const int area = lds_s32(st);
        if (area) {
          InstStats* dst = p.stats + (size_t)it.b * d.max_n + it.i0 + lane;
          atomicAdd(&dst->area, (unsigned)area);
          atomicMin(&dst->minx, lds_s32(st + 4));
          atomicMin(&dst->miny, lds_s32(st + 8));
          atomicMax(&dst->maxx, lds_s32(st + 12));
          atomicMax(&dst->maxy, lds_s32(st + 16));
          sts_s32(st, 0); sts_s32(st + 4, INT_MAX); sts_s32(st + 8, INT_MAX); sts_s32(st + 12, -1); sts_s32(st + 16, -1);
        }
      }
      __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tm.end();
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  // The last CTA to finish re-arms the work counter for the next launch (every CTA has drawn its final, invalid
  // item by now), so no memset node is needed between launches - also under CUDA-graph replay.
  if (threadIdx.x == 32) {
    __threadfence();
    if (atomicAdd(p.work_counter + 1, 1) == (int)gridDim.x - 1) {
      p.work_counter[0] = 0;
      p.work_counter[1] = 0;
    }
  }
  if (warp == 0) {
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
  int nbuf;
  int ni;            // instances per group = accumulator columns read back per tile (8 or 16)
  int groups;        // instance groups per frame = ceil(max_n / ni)
  int chunk_floats;
  bool generic;      // any up-sampling scale (row table + per-pixel source indices) instead of exact 4x
  int max_rows;      // generic: most dst rows owned by one prototype row pair
  int zero_rows;     // rows of the shared-memory zero buffer
  size_t smem_bytes;
  unsigned long long* timing;   // device buffer when VA_FUSED_TIMING=1
  int* work_counter;
  // cached tensor map
  const float* map_ptr;
  int map_B;
  CUtensorMap map;
};

FusedPlan* fused_plan_create(const Dims& d, int device, char* err, size_t errlen) {
  const bool exact4 = (d.H == 4 * d.mh && d.W == 4 * d.mw);
  int max_rows = 6;
  if (!exact4) {
    // generic scale: up-sampling in both directions, whole 16-pixel pieces, at most 8 dst rows per prototype row pair
    if (d.H < d.mh || d.W < d.mw || (d.W % 16) != 0) { snprintf(err, errlen, "tcgen05 path needs H=4*mh, W=4*mw, or an up-sampling scale with W %% 16 == 0"); return nullptr; }
    if (d.H > 32767) { snprintf(err, errlen, "tcgen05 path: H <= 32767"); return nullptr; }
    max_rows = 0;
    int prev = -1, first = 0;
    for (int Y = 0; Y <= d.H; ++Y) {                    // the same fp32 arithmetic the kernel uses for its row table
      int y0 = d.mh;
      if (Y < d.H) {
        float src = fmaf(d.sy, (float)Y + 0.5f, -0.5f);
        src = src < 0.f ? 0.f : src;
        y0 = (int)src < d.mh - 1 ? (int)src : d.mh - 1;
      }
      if (y0 != prev) {
        if (prev >= 0 && Y - first > max_rows) max_rows = Y - first;
        if (prev >= 0 && y0 != prev + 1) { snprintf(err, errlen, "tcgen05 path: a prototype row owns no dst row"); return nullptr; }
        prev = y0; first = Y;
      }
    }
    if (max_rows > 12) { snprintf(err, errlen, "tcgen05 path handles vertical scales up to 8x (at most 12 dst rows per prototype row, got %d)", max_rows); return nullptr; }
  }
  if (d.max_n > kMaxInst) { snprintf(err, errlen, "tcgen05 path handles max_n <= %d", kMaxInst); return nullptr; }
  if ((d.mw % 4) != 0 || (d.W % 16) != 0) { snprintf(err, errlen, "mw %% 4 / W %% 16"); return nullptr; }
  if (((size_t)d.mh * d.mw) % 4 != 0) { snprintf(err, errlen, "P %% 4"); return nullptr; }
  if (d.mw * 5 < 2 * kTileM + 1) { snprintf(err, errlen, "tcgen05 path needs mw >= 52 (items of >= 2 tiles)"); return nullptr; }
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
  // Instances are processed in groups of `ni` (8 or 16 accumulator columns); a frame with more instances becomes
  // several work items per band that re-read the same prototype rows (from L2).  Pick the group size, the row pairs
  // per chunk and the number of chunk buffers that fit next to the tile rings.  Groups of 16 halve the prototype
  // re-reads and splits per mask byte, but their instantiation sits at the 128-register cap and spills since the
  // per-row by-products were added: groups of 8 are tried first (measured at 640^2: n = 16 0.342 -> 0.288 ms,
  // n = 32 0.285 -> 0.254 ms; cfg2 0.60 -> 0.51 ms at equal band counts).
  const size_t limit = (size_t)prop.sharedMemPerBlockOptin - 1024;
  int pr = 0, nbuf = 0, ni = 0;
  int force_ni = 0;
  if (const char* e = getenv("VA_FUSED_GSIZE")) force_ni = atoi(e);   // tuning aid: 8 or 16
  // rows of the zero buffer: a whole chunk's dst rows when that is small (exact 4x: 4 pr + 2), else pieces of <= 16 KB
  auto zero_rows_for = [&](int cand) {
    const int want = exact4 ? 4 * cand + 2 : cand * max_rows;
    const int fit = (16 * 1024) / d.W > 0 ? (16 * 1024) / d.W : 1;
    return (exact4 || want <= fit) ? want : fit;
  };
  int zrows = 0;
  const int order[2] = {8, 16};
  for (int gi = 0; gi < 2; ++gi) {
    const int g = order[gi];
    if (g == 16 && d.max_n <= 8) continue;
    if (force_ni && g != force_ni && d.max_n > 8) continue;
    for (int cand : {4, 2}) {
      for (int nb : {4, 3}) {
        const int cf = (cand + 1) * g * d.mw;
        if ((size_t)fused_smem_map(cf, nb, d.H, d.mh, zero_rows_for(cand) * d.W).total + 1024 <= limit) { pr = cand; nbuf = nb; ni = g; zrows = zero_rows_for(cand); break; }
      }
      if (pr) break;
    }
    if (pr) break;
  }
  if (!pr) { delete pl; snprintf(err, errlen, "chunk buffers do not fit in shared memory (max_n=%d, mw=%d)", d.max_n, d.mw); return nullptr; }
  pl->pr = pr;
  pl->nbuf = nbuf;
  pl->ni = ni;
  pl->groups = ceil_div(d.max_n, ni);
  pl->chunk_floats = (pr + 1) * ni * d.mw;
  pl->generic = !exact4;
  pl->max_rows = max_rows;
  pl->zero_rows = zrows;
  pl->smem_bytes = (size_t)fused_smem_map(pl->chunk_floats, nbuf, d.H, d.mh, zrows * d.W).total + 1024;
  cudaError_t e = cudaSuccess;
#define VA_ATTR(WM, NI, DG, NB) \
  if (e == cudaSuccess) e = pl->generic ? cudaFuncSetAttribute((const void*)fused_tc_kernel<WM, NI, DG, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes) \
                                        : cudaFuncSetAttribute((const void*)fused_tc_kernel<WM, NI, DG, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes)
#define VA_ATTR4(NI, NB) VA_ATTR(true, NI, false, NB); VA_ATTR(false, NI, false, NB); VA_ATTR(true, NI, true, NB); VA_ATTR(false, NI, true, NB)
  if (pl->nbuf == 4) { if (pl->ni == 8) { VA_ATTR4(8, 4); } else { VA_ATTR4(16, 4); } }
  else               { if (pl->ni == 8) { VA_ATTR4(8, 3); } else { VA_ATTR4(16, 3); } }
#undef VA_ATTR4
#undef VA_ATTR
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(%zu B): %s", pl->smem_bytes, cudaGetErrorString(e)); delete pl; return nullptr; }
  if (cudaMalloc(&pl->work_counter, 2 * sizeof(int)) != cudaSuccess || cudaMemset(pl->work_counter, 0, 2 * sizeof(int)) != cudaSuccess) {
    snprintf(err, errlen, "cudaMalloc(work counter) failed");
    delete pl;
    return nullptr;
  }
  const char* tenv = getenv("VA_FUSED_TIMING");
  if (tenv && tenv[0] == '1') cudaMalloc(&pl->timing, (size_t)pl->num_sms * 5 * 8 * sizeof(unsigned long long));
  return pl;
}

void fused_plan_destroy(FusedPlan* p) {
  if (p && p->timing) cudaFree(p->timing);
  if (p && p->work_counter) cudaFree(p->work_counter);
  delete p;
}

cudaError_t launch_fused(FusedPlan* pl, const Dims& d, const float* protos, const float* coefs, const float* boxes,
                         const int* counts, int B, uint8_t* masks, float* logits_dbg, const MaskSinks& sinks,
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
  p.stats = sinks.stats; p.lattice = sinks.lattice; p.B = B;
  p.rowsum = sinks.rowsum;
  static const bool no_rowsum = getenv("VA_NO_ROWSUM") != nullptr;     // developer A/B switch: cost of the per-row by-products
  if (no_rowsum) p.rowsum = nullptr;
  p.bits16 = reinterpret_cast<uint16_t*>(sinks.bits);
  if (!masks && !sinks.bits) { snprintf(err, errlen, "grid-only launch without a bit-mask buffer"); return cudaErrorInvalidValue; }
  p.timing = pl->timing;
  p.work_counter = pl->work_counter;
  p.pr = pl->pr; p.pr_shift = pl->pr == 4 ? 2 : 1; p.nst = pl->ni; p.gsize = pl->ni; p.groups = pl->groups; p.chunk_floats = pl->chunk_floats; p.nbuf = pl->nbuf;
  p.zero_rows = pl->zero_rows; p.zero_bytes = pl->zero_rows * d.W; p.max_rows = pl->max_rows;
  // Bands per frame: items are stolen dynamically, so what matters is enough items per CTA for a short tail
  // (>= ~8) against the one-row halo every band recomputes and re-reads (1/ppb).
  const int max_bands = (d.mh / (2 * pl->pr)) > 0 ? d.mh / (2 * pl->pr) : 1;
  int nb = 1;
  // generic scale (store-dominated, 6-12 dst rows per prototype row): many short bands - measured at cfg2: 8 bands
  // (3.5 items per SM) 0.80 ms, 20 bands 0.48 ms, 40 bands 0.45 ms
  const int band_cap = pl->generic ? ((d.mh / pl->pr) > 0 ? d.mh / pl->pr : 1) : (max_bands < 8 ? max_bands : 8);
  const long want_items = (pl->generic ? 32L : 8L) * pl->num_sms;
  while (nb < band_cap && (long)B * nb * pl->groups < want_items) ++nb;
  // tiny batches (single-frame latency): more, shorter bands - down to one chunk of pr row pairs - until every SM has an item
  const int max_bands_tiny = (d.mh / pl->pr) > 0 ? d.mh / pl->pr : 1;
  while (nb < max_bands_tiny && (long)B * nb * pl->groups < (long)pl->num_sms) ++nb;
  if (const char* e = getenv("VA_FUSED_NBANDS")) { const int v = atoi(e); if (v >= 1 && v <= max_bands_tiny) nb = v; }   // tuning aid
  p.ppb = ceil_div(ceil_div(d.mh, nb), pl->pr) * pl->pr;
  p.nbands = ceil_div(d.mh, p.ppb);
  p.n_items = B * p.nbands * p.groups;
  const int grid = p.n_items < pl->num_sms ? p.n_items : pl->num_sms;
  const bool diag = (logits_dbg != nullptr) || (pl->timing != nullptr);   // debug logits / role timing: separate instantiation
#define VA_LAUNCH(WM, NI, DG, NB) do { if (pl->generic) fused_tc_kernel<WM, NI, DG, NB, true><<<grid, kThreads, pl->smem_bytes, st>>>(pl->map, p); \
                                       else fused_tc_kernel<WM, NI, DG, NB, false><<<grid, kThreads, pl->smem_bytes, st>>>(pl->map, p); } while (0)
#define VA_LAUNCH_NB(WM, NI, DG) do { if (pl->nbuf == 4) VA_LAUNCH(WM, NI, DG, 4); else VA_LAUNCH(WM, NI, DG, 3); } while (0)
  if (pl->ni == 8) {
    if (diag) { if (masks) VA_LAUNCH_NB(true, 8, true); else VA_LAUNCH_NB(false, 8, true); }
    else      { if (masks) VA_LAUNCH_NB(true, 8, false); else VA_LAUNCH_NB(false, 8, false); }
  } else {
    if (diag) { if (masks) VA_LAUNCH_NB(true, 16, true); else VA_LAUNCH_NB(false, 16, true); }
    else      { if (masks) VA_LAUNCH_NB(true, 16, false); else VA_LAUNCH_NB(false, 16, false); }
  }
#undef VA_LAUNCH_NB
#undef VA_LAUNCH
  if (pl->timing) {   // developer diagnostic: blocking read-back, per-role wait / busy cycles averaged over CTAs
    cudaStreamSynchronize(st);
    static unsigned long long h[256 * 5 * 8];
    cudaMemcpy(h, pl->timing, (size_t)grid * 5 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const char* roles[5] = {"(unused)", "tma", "split+issue", "epilogue", "upsample"};
    const char* waits[5][6] = {{"", "", "", "", "", ""},
                               {"hi_empty", "", "", "", "", ""},
                               {"b_empty", "hi_full", "lo_empty", "b_full", "lo_full", "acc_empty"},
                               {"acc_full", "ch_empty", "tmem_ld", "", "", ""},
                               {"ch_full", "", "", "", "", ""}};
    fprintf(stderr, "[va timing] grid=%d items=%d nbands=%d ppb=%d pr=%d nbuf=%d gsize=%d groups=%d\n", grid, p.n_items, p.nbands, p.ppb, p.pr, p.nbuf, p.gsize, p.groups);
    for (int r = 1; r < 5; ++r) {
      double tot = 0, mx = 0, w[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < grid; ++c) {
        const double t = (double)h[(c * 5 + r) * 8];
        tot += t;
        if (t > mx) mx = t;
        for (int j = 0; j < 6; ++j) w[j] += (double)h[(c * 5 + r) * 8 + 1 + j];
      }
      double ws = 0;
      fprintf(stderr, "[va timing] %-11s avg %8.0f max %8.0f cyc |", roles[r], tot / grid, mx);
      for (int j = 0; j < 6; ++j) {
        ws += w[j];
        if (waits[r][j][0]) fprintf(stderr, " %s %4.1f%%", waits[r][j], 100 * w[j] / tot);
      }
      fprintf(stderr, " | busy %4.1f%%\n", 100 * (tot - ws) / tot);
    }
  }
  return cudaGetLastError();
}

}  // namespace va
