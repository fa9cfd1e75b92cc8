// Standalone probe (GPU box only): cost model of small tcgen05.mma kind::tf32 instructions.
// Measures cycles for chains of MMAs: dependent (same accumulator) vs independent accumulators, N sweep.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff); d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16; d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// smem: A 16 KB @0 (K-major 128x32), A2 16 KB @16K, B 32 KB @32K (up to 256 rows), barrier @64K
__global__ void __launch_bounds__(128, 1) timing(int N, int n_mma, int n_acc, int reps, int two_a, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sA2 = base + 16384, sB = base + 32768, bar = base + 65536, slot = base + 65536 + 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4 * i), "f"(0.001f * (i % 97)));
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    long long best = 1LL << 60, best_issue = 1LL << 60;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
      for (int m = 0; m < n_mma; ++m) {
        const int ks = m & 3;
        const uint32_t a = (two_a && (m & 4)) ? sA2 : sA;
        const uint64_t da = make_desc(a + ks * 32, 16, 1024), db = make_desc(sB + ks * 32, 16, 1024);
        const uint32_t d = tmem + (uint32_t)((m % n_acc) * N);
        const uint32_t acc = (m >= n_acc) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      const long long t1 = clock64();
      mbar_wait(bar, phase); phase ^= 1;
      const long long t2 = clock64();
      if (t2 - t0 < best) best = t2 - t0;
      if (t1 - t0 < best_issue) best_issue = t1 - t0;
    }
    out[0] = best; out[1] = best_issue;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  struct C { int N, n_mma, n_acc, two_a; } cs[] = {
      {16, 1, 1, 0}, {16, 4, 1, 0}, {16, 12, 1, 0}, {16, 12, 3, 0}, {16, 12, 12, 0}, {16, 24, 6, 0}, {16, 24, 2, 0}, {16, 48, 12, 0},
      {32, 12, 1, 0}, {64, 12, 1, 0}, {128, 12, 1, 0}, {256, 12, 1, 0}, {256, 1, 1, 0}, {256, 4, 1, 0}, {64, 12, 3, 0}, {8, 12, 1, 0}};
  for (auto& c : cs) {
    timing<<<1, 128, 70000>>>(c.N, c.n_mma, c.n_acc, 20, c.two_a, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=128 N=%3d n_mma=%2d n_acc=%2d : total %6lld cyc (%.1f / mma), issue %5lld cyc\n", c.N, c.n_mma, c.n_acc, h[0], (double)h[0] / c.n_mma, h[1]);
  }
  return 0;
}
