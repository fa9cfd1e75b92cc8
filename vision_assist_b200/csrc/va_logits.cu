// CUDA-core (FFMA) prototype x coefficient contraction + box crop at prototype resolution.
//
//   logits[b][i][p] = crop_i(p) * sum_k coefs[b][i][k] * protos[b][k][p]
//
// follows ops.process_mask (testing/old/segmenting_using_tflite/ops.py:724-734): fp32 matmul,
// boxes scaled by fl32(mw/iw), fl32(mh/ih) (:725-732), crop_mask's half-open float box (:688-704).
// This is the debug / `logits_out` path and the VA_CFG_NO_TENSOR_CORE contraction; the production
// contraction is the tcgen05 kernel in va_fused_tc.cu.  HBM-bound: 4*K bytes in per pixel.
#include "va_common.cuh"

namespace va {

constexpr int kLogitsThreads = 128;
constexpr int kPxPerThread = 2;

__global__ void __launch_bounds__(kLogitsThreads)
logits_kernel(Dims d, const float* __restrict__ protos, const float* __restrict__ coefs,
              const float* __restrict__ boxes, const int* __restrict__ counts, float* __restrict__ logits) {
  __shared__ float s_coefT[kProtoK][kMaxInst];   // [k][i]: the n values for one k are contiguous
  __shared__ float s_box[kMaxInst][4];
  __shared__ unsigned s_live;                    // instances whose box rows intersect the rows of this CTA's pixels

  const int b = blockIdx.y;
  const int n = min(counts[b], d.max_n);
  const int P = d.mh * d.mw;
  if (n <= 0) return;

  for (int t = threadIdx.x; t < kProtoK * kMaxInst; t += kLogitsThreads) {
    const int i = t / kProtoK, k = t % kProtoK;
    s_coefT[k][i] = (i < n) ? coefs[((size_t)b * d.max_n + i) * d.K + k] : 0.f;
  }
  for (int t = threadIdx.x; t < n * 4; t += kLogitsThreads) {
    const int i = t >> 2, c = t & 3;
    const float v = boxes[((size_t)b * d.max_n + i) * 4 + c];
    s_box[i][c] = __fmul_rn(v, (c & 1) ? d.hr : d.wr);       // x1,x2 * wr ; y1,y2 * hr
  }
  if (threadIdx.x == 0) s_live = 0u;
  __syncthreads();
  if (threadIdx.x < kMaxInst) {
    // crop_mask keeps row r iff r >= y1 && r < y2: an instance none of whose kept rows is among this CTA's rows
    // gets zeros without the 32 multiply-adds per pixel (a NaN bound fails both comparisons -> not live: zeros)
    const int i = threadIdx.x;
    const int first_px = blockIdx.x * kLogitsThreads * kPxPerThread;
    const int last_px = min(first_px + kLogitsThreads * kPxPerThread, P) - 1;
    const float ra = (float)(first_px / d.mw), rb = (float)(last_px / d.mw);
    if (i < n && rb >= s_box[i][1] && ra < s_box[i][3]) atomicOr(&s_live, 1u << i);
  }
  __syncthreads();
  const unsigned live = s_live;

  const int p0 = (blockIdx.x * kLogitsThreads + threadIdx.x) * kPxPerThread;
  if (p0 >= P) return;
  const bool two = (p0 + 1 < P);

  const float* pp = protos + (size_t)b * d.K * P + p0;
  float2 v[kProtoK];
#pragma unroll
  for (int k = 0; k < kProtoK; ++k) {
    if (two) v[k] = __ldg(reinterpret_cast<const float2*>(pp + (size_t)k * P));
    else v[k] = make_float2(__ldg(pp + (size_t)k * P), 0.f);
  }
  const int py0 = p0 / d.mw, px0 = p0 - py0 * d.mw;
  const int p1 = p0 + 1;
  const int py1 = p1 / d.mw, px1 = p1 - py1 * d.mw;
  const float fx0 = (float)px0, fy0 = (float)py0, fx1 = (float)px1, fy1 = (float)py1;

  for (int i0 = 0; i0 < n; i0 += 8) {
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
    if ((live >> i0) & 0xffu) {                      // block-uniform
#pragma unroll
    for (int k = 0; k < kProtoK; ++k) {
      const float4 c0 = *reinterpret_cast<const float4*>(&s_coefT[k][i0]);
      const float4 c1 = *reinterpret_cast<const float4*>(&s_coefT[k][i0 + 4]);
      const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[j] = fmaf(cc[j], v[k].x, a0[j]);
        a1[j] = fmaf(cc[j], v[k].y, a1[j]);
      }
    }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j;
      if (i >= n) break;
      const float x1 = s_box[i][0], y1 = s_box[i][1], x2 = s_box[i][2], y2 = s_box[i][3];
      const bool k0 = (fx0 >= x1) && (fx0 < x2) && (fy0 >= y1) && (fy0 < y2);
      const bool k1 = (fx1 >= x1) && (fx1 < x2) && (fy1 >= y1) && (fy1 < y2);
      float* o = logits + ((size_t)b * d.max_n + i) * P + p0;
      if (two) *reinterpret_cast<float2*>(o) = make_float2(k0 ? a0[j] : 0.f, k1 ? a1[j] : 0.f);
      else *o = k0 ? a0[j] : 0.f;
    }
  }
}

cudaError_t launch_logits(const Dims& d, const float* protos, const float* coefs, const float* boxes,
                          const int* counts, int B, float* logits, cudaStream_t st) {
  const int P = d.mh * d.mw;
  dim3 grid(ceil_div(P, kLogitsThreads * kPxPerThread), B);
  logits_kernel<<<grid, kLogitsThreads, 0, st>>>(d, protos, coefs, boxes, counts, logits);
  return cudaGetLastError();
}

}  // namespace va
