// Standalone probe (GPU box only): do tcgen05.mma instructions issued from two different warps overlap?
// TS mode, M=128, N=32, kind::tf32.  Each issuing thread runs a chain of n_mma MMAs into its own accumulator.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
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
__global__ void __launch_bounds__(128, 1) probe(int n_issuers, int n_mma, int reps, int N, long long* tout) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sB = base, bar = base + 32768, slot = base + 32768 + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192; i += 128) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB + 4 * i), "f"(0.01f * (i % 13)));
  if (tid == 0) {
    for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * w));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  __shared__ long long s_t[4];
  long long best = 1LL << 60;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    if (lane == 0 && warp < n_issuers) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const long long t0 = clock64();
      for (int m = 0; m < n_mma; ++m) {
        const int ks = m & 3;
        const uint64_t db = make_desc(sB + ks * 32, 16, 1024);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + warp * 64), "r"(tmem + 256 + ks * 8), "l"(db), "r"(idesc), "r"((uint32_t)(m > 0)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * warp) : "memory");
      mbar_wait(bar + 8 * warp, rep & 1);
      s_t[warp] = clock64() - t0;
    }
    __syncthreads();
    if (tid == 0) { long long mx = 0; for (int w = 0; w < n_issuers; ++w) mx = s_t[w] > mx ? s_t[w] : mx; if (mx < best) best = mx; }
  }
  if (tid == 0) tout[0] = best;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8); long long h;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  int cfgs[][3] = {{1, 8, 32}, {2, 8, 32}, {4, 8, 32}, {1, 16, 32}, {2, 16, 32}, {1, 8, 128}, {2, 8, 128}, {1, 32, 32}, {4, 32, 32}};
  for (auto& c : cfgs) {
    probe<<<1, 128, 40000>>>(c[0], c[1], 20, c[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("issuers=%d n_mma=%2d each N=%3d : %6lld cyc total (%.1f cyc per MMA overall)\n", c[0], c[1], c[2], h, (double)h / (c[0] * c[1]));
  }
  return 0;
}
