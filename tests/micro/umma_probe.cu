// Standalone probe (GPU box only): one CTA, one 128x16x32 tf32 tile through tcgen05.mma with
// different operand layouts, result dumped and compared with a host reference.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Args { int a_mn; uint32_t a_lbo, a_sbo, a_kstep; int use_tma; int n_passes; };

// smem: A tile 16 KB @0, B tile 2 KB @16384, barrier, tmem slot
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ Aimg /*16KB smem image*/,
                                                const float* __restrict__ Bimg /*2KB smem image*/, float* __restrict__ out /*[128][16]*/, Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  float* sA = (float*)base;
  float* sB = (float*)(base + 16384);
  uint64_t* bar = (uint64_t*)(base + 16384 + 2048);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!a.use_tma) for (int i = tid; i < 4096; i += 128) sA[i] = Aimg[i];
  for (int i = tid; i < 512; i += 128) sB[i] = Bimg[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    if (a.use_tma) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar2)), "r"(16384u) : "memory");
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                         smem_u32(base + j * 4096)), "l"((uint64_t)&tmap), "r"(smem_u32(bar2)), "r"(j * 32), "r"(0), "r"(0) : "memory");
      mbar_wait(bar2, 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.a_mn ? 1 : 0) << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    for (int pass = 0; pass < a.n_passes; ++pass)
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t da = make_desc(smem_u32(sA) + ks * a.a_kstep, a.a_lbo, a.a_sbo);
        const uint64_t db = make_desc(smem_u32(sB) + ks * 32, 16, 1024);
        const uint32_t acc = (pass > 0 || ks > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  __syncwarp();
  mbar_wait(bar, 0);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[tid * 16 + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static float Aval(int m, int k) { return (float)((m * 3 + k * 5) % 17 - 8) * 0.25f; }
static float Bval(int n, int k) { return (float)((n * 7 + k * 3) % 13 - 6) * 0.5f; }

int main() {
  const int M = 128, N = 16, K = 32;
  static float A[M][K], B[N][K], ref[M][N];
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m][k] = Aval(m, k);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n][k] = Bval(n, k);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m][k] * B[n][k]; ref[m][n] = (float)s; }
  // smem images
  static float imgA_K[4096], imgA_MN[4096], imgB[512];
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {          // K-major SW128: row m (128 B), chunk (k/4) ^ (m&7)
    int off = m * 32 + (((k / 4) ^ (m & 7)) * 4) + (k % 4);
    imgA_K[off] = A[m][k];
    // MN-major SW128: block m/32 (4 KB) ; row k (128 B) ; chunk ((m%32)/4) ^ (k&7)
    int off2 = (m / 32) * 1024 + k * 32 + ((((m % 32) / 4) ^ (k & 7)) * 4) + (m % 4);
    imgA_MN[off2] = A[m][k];
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) imgB[n * 32 + (((k / 4) ^ (n & 7)) * 4) + (k % 4)] = B[n][k];
  // global A^T [K][P] for the TMA variant (P = 4096 pixels, tile = first 128)
  const int P = 4096;
  float* hAT = (float*)calloc((size_t)K * P, 4);
  for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) hAT[(size_t)k * P + m] = A[m][k];
  float *dA_K, *dA_MN, *dB, *dAT, *dout;
  CK(cudaMalloc(&dA_K, 16384)); CK(cudaMalloc(&dA_MN, 16384)); CK(cudaMalloc(&dB, 2048)); CK(cudaMalloc(&dAT, (size_t)K * P * 4)); CK(cudaMalloc(&dout, M * N * 4));
  CK(cudaMemcpy(dA_K, imgA_K, 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dA_MN, imgA_MN, 16384, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, imgB, 2048, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAT, hAT, (size_t)K * P * 4, cudaMemcpyHostToDevice));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)K, 1}; cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)P * K * 4};
  cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
  CUresult r = ((PFN_encodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dAT, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  struct V { const char* name; Args a; const float* img; } vs[] = {
      {"V1 A K-major SW128 (lbo16,sbo1024,kstep32)", {0, 16, 1024, 32, 0, 1}, dA_K},
      {"V2 A MN-major SW128 (lbo4096,sbo1024,kstep1024)", {1, 4096, 1024, 1024, 0, 1}, dA_MN},
      {"V3 A MN-major SW128 (lbo1024,sbo4096,kstep1024)", {1, 1024, 4096, 1024, 0, 1}, dA_MN},
      {"V4 A MN-major via TMA (lbo4096,sbo1024)", {1, 4096, 1024, 1024, 1, 1}, dA_MN},
      {"V5 A MN-major via TMA (lbo1024,sbo4096)", {1, 1024, 4096, 1024, 1, 1}, dA_MN},
      {"V6 V1 x2 passes (accumulate, expect 2x)", {0, 16, 1024, 32, 0, 2}, dA_K},
  };
  static float out[M][N];
  for (auto& v : vs) {
    CK(cudaMemset(dout, 0xff, M * N * 4));
    probe<<<1, 128, 32768>>>(map, v.img, dB, dout, v.a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: kernel error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(out, dout, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int nz = 0;
    const float scale = (float)v.a.n_passes;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { maxerr = fmax(maxerr, fabs(out[m][n] - scale * ref[m][n])); nz += out[m][n] != 0; }
    printf("%s: maxerr %.6g nonzero %d/%d  out[0][0..3]=%g %g %g %g ref=%g %g %g %g  out[37][5]=%g ref=%g out[100][9]=%g ref=%g\n", v.name, maxerr, nz, M * N,
           out[0][0], out[0][1], out[0][2], out[0][3], ref[0][0], ref[0][1], ref[0][2], ref[0][3], out[37][5], ref[37][5], out[100][9], ref[100][9]);
  }
  // truncation probe: A = 1 + 2^-12 (below tf32 resolution), B = 1 on k=0 only
  {
    for (int i = 0; i < 4096; ++i) imgA_K[i] = 0.f;
    for (int i = 0; i < 512; ++i) imgB[i] = 0.f;
    const float vals[4] = {1.0f + 1.0f / 4096, 1.0f + 3.0f / 4096 /*above half ulp(2^-10)=2^-11*/, 1.0f + 1.0f / 2048, -(1.0f + 3.0f / 4096)};
    for (int m = 0; m < 4; ++m) imgA_K[m * 32 + (((0 / 4) ^ (m & 7)) * 4)] = vals[m];
    imgB[0] = 1.0f;
    CK(cudaMemcpy(dA_K, imgA_K, 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, imgB, 2048, cudaMemcpyHostToDevice));
    Args a = {0, 16, 1024, 32, 0, 1};
    probe<<<1, 128, 32768>>>(map, dA_K, dB, dout, a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout, M * N * 4, cudaMemcpyDeviceToHost));
    printf("truncation probe: in %.9g %.9g %.9g %.9g -> out %.9g %.9g %.9g %.9g (truncate => 1, 1, 1.00048828, -1)\n", vals[0], vals[1], vals[2], vals[3],
           out[0][0], out[1][0], out[2][0], out[3][0]);
  }
  return 0;
}
