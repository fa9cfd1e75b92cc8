// Standalone probe (GPU box only): tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY
// (written with tcgen05.st from registers), B in shared memory (K-major SW128).
// Checks numerics against a host reference and times chains of MMAs (TS mode vs SS mode).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
               "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// A [128][32] row-major in global; B image (K-major SW128, 16 rows) in global; out [128][16]; timing out
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ Bimg, float* __restrict__ out, int n_mma, int reps,
                                                long long* tout) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sB = base, bar = base + 4096, slot = base + 4096 + 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 512; i += 128) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB + 4 * i), "f"(Bimg[i]));
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  // A row of this thread -> TMEM lane tid, columns [32, 64)   (D accumulator at columns [0, 16))
  uint32_t r[32];
  for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(A[tid * 32 + k]);
  tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 32, r);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    long long best = 1LL << 60, best_issue = 1LL << 60;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
      for (int m = 0; m < n_mma; ++m) {
        const int ks = m & 3;
        const uint64_t db = make_desc(sB + ks * 32, 16, 1024);
        const uint32_t acc = (m > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 32 + ks * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      const long long t1 = clock64();
      mbar_wait(bar, phase); phase ^= 1;
      const long long t2 = clock64();
      if (t2 - t0 < best) best = t2 - t0;
      if (t1 - t0 < best_issue) best_issue = t1 - t0;
    }
    tout[0] = best; tout[1] = best_issue;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t o[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
                 "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[tid * 16 + i] = __uint_as_float(o[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}
static float Aval(int m, int k) { return (float)((m * 3 + k * 5) % 17 - 8) * 0.25f; }
static float Bval(int n, int k) { return (float)((n * 7 + k * 3) % 13 - 6) * 0.5f; }
int main() {
  const int M = 128, N = 16, K = 32;
  static float A[M][K], B[N][K], ref[M][N], imgB[512], out[M][N];
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m][k] = Aval(m, k);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n][k] = Bval(n, k);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m][k] * B[n][k]; ref[m][n] = (float)s; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) imgB[n * 32 + (((k / 4) ^ (n & 7)) * 4) + (k % 4)] = B[n][k];
  float *dA, *dB, *dout; long long* dt;
  cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dB, sizeof(imgB)); cudaMalloc(&dout, sizeof(out)); cudaMalloc(&dt, 16);
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice); cudaMemcpy(dB, imgB, sizeof(imgB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  int cfgs[][2] = {{4, 1}, {1, 20}, {4, 20}, {12, 20}, {48, 20}};
  for (auto& c : cfgs) {
    probe<<<1, 128, 16384>>>(dA, dB, dout, c[0], c[1], dt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2]; cudaMemcpy(h, dt, 16, cudaMemcpyDeviceToHost); cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
    if (c[0] == 4 && c[1] == 1) {
      double maxerr = 0; for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) maxerr = fmax(maxerr, fabs(out[m][n] - ref[m][n]));
      printf("TS numerics (A via tcgen05.st): maxerr %.6g  out[0][0..3]=%g %g %g %g ref=%g %g %g %g out[100][9]=%g ref=%g\n", maxerr, out[0][0], out[0][1], out[0][2],
             out[0][3], ref[0][0], ref[0][1], ref[0][2], ref[0][3], out[100][9], ref[100][9]);
    }
    printf("TS M=128 N=16 n_mma=%2d : total %6lld cyc (%.1f / mma), issue %5lld cyc\n", c[0], h[0], (double)h[0] / c[0], h[1]);
  }
  return 0;
}
