// C ABI of libva_sm100.so (include/vision_assist_b200.h): context, record layout, launch plumbing.
#include <cuda_fp16.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "va_common.cuh"
#include "va_contour_core.h"

using namespace va;

constexpr int kProfMax = 512;

struct va_ctx {
  va_config cfg;
  Dims d;
  va_layout layout;
  Scratch scratch;
  int num_sms;
  int logits_chunk;            // frames of logits scratch (CUDA-core path)
  FusedPlan* plan;             // tcgen05 path, nullptr when unavailable / disabled
  int last_launches;
  char err[512];
  // optional per-call device timing (va_profile_enable): events on the caller's stream
  int prof_on, prof_count;     // prof_on = sampling stride (0 = off): every prof_on-th call is timed
  int prof_calls;
  cudaEvent_t prof_ev[kProfMax][3];
  // host-buffer pipeline (va_run_fused_host), created lazily
  bool host_ready;
  int host_chunk;
  cudaStream_t s_in, s_compute, s_out;
  cudaEvent_t ev_in[2], ev_done[2], ev_out[2];
  float* d_protos[2];
  uint16_t* d_protos_h[2];     // fp16 staging of va_run_fused_host_f16 (allocated on first use)
  float* d_coefs[2];
  float* d_boxes[2];
  int* d_counts[2];
  uint8_t* d_records[2];
  uint8_t* d_masks[2];
};

static char g_create_err[512] = "";

static void set_err(va_ctx* c, const char* fmt, ...) {
  char* dst = c ? c->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
}

#define VA_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return VA_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

static int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Entry points run on the context's device and put the caller's current device back on every exit path (a process
// that drives several GPUs from one thread must not find its current device switched by a library call).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err != cudaSuccess) { prev = -1; err = cudaSetDevice(device); }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define VA_ON_DEVICE(ctx)                                                                                   \
  DeviceGuard guard__((ctx)->cfg.device);                                                                   \
  if (guard__.err != cudaSuccess) { set_err((ctx), "cudaSetDevice(%d) failed: %s", (ctx)->cfg.device, cudaGetErrorString(guard__.err)); return VA_ERR_CUDA; }

static bool compute_dims(const va_config& c, Dims& d, va_layout& L, char* why, size_t n) {
  if (c.K != kProtoK) { snprintf(why, n, "K must be %d (got %d)", kProtoK, c.K); return false; }
  if (c.max_n < 1 || c.max_n > kMaxInst) { snprintf(why, n, "max_n must be in [1,%d]", kMaxInst); return false; }
  if (c.H < 8 || c.W < 16 || c.mh < 2 || c.mw < 4 || (c.mw % 4) != 0) {
    snprintf(why, n, "unsupported geometry H=%d W=%d mh=%d mw=%d (mw must be a multiple of 4)", c.H, c.W, c.mh, c.mw);
    return false;
  }
  if (c.gs < 4 || c.gs > c.H || c.gs > c.W) { snprintf(why, n, "gs must be in [4, min(H,W)]"); return false; }
  if (c.max_batch < 1 || c.max_batch > 65535) { snprintf(why, n, "max_batch must be in [1,65535]"); return false; }
  memset(&d, 0, sizeof(d));
  d.H = c.H; d.W = c.W; d.mh = c.mh; d.mw = c.mw; d.K = c.K; d.max_n = c.max_n; d.gs = c.gs;
  d.flags = c.flags;
  if (const char* e = getenv("VA_TAIL_TIMING")) { if (e[0] == '1') d.flags |= 1 << 30; if (e[0] == '2') d.flags |= 1 << 28; if (e[0] == '3') d.flags |= 1 << 27; }
  if (const char* e = getenv("VA_TAIL_BLOCK")) d.flags |= (atoi(e) & 0xfff) << 8;                 // which frame VA_TAIL_TIMING=1 clocks   // developer diagnostics (va_tail.cu)
  if (const char* e = getenv("VA_TAIL_ROLES")) { if (e[0] == '3') d.flags |= 1 << 29; }    // tuning aid (va_tail.cu)
  const int half = c.gs / 2;
  d.lat_rows = ceil_div(c.H - half, c.gs);
  d.lat_cols = ceil_div(c.W - half, c.gs);
  d.lat_words = ceil_div(d.lat_cols, 32);
  d.nblk = ceil_div(c.W, cc::kRowBlock);
  d.bit_words = ceil_div(c.W, 32);
  d.plane_rows = ceil_div(c.H, c.gs);
  int s = (int)((long long)c.H * 7 / 8);                   // int(H * 0.875), FrameProcessor.py:126
  s = s + (c.gs - s % c.gs) % c.gs;                        // :127
  d.band_start = s;
  const int nband = (s < c.H) ? ceil_div(c.H - s, c.gs) : 0;
  d.rmax = ceil_div(c.H, c.gs) + nband + 2;
  d.cmax = ceil_div(c.W, c.gs);
  d.cwords = ceil_div(d.cmax, 32);
  if (d.cwords > 64) { snprintf(why, n, "W / gs must not exceed 2048 columns"); return false; }
  d.pmax = (d.cmax + 1) / 2 + 1;
  int o = 0;
  L.off_header = o; o += 64;
  d.off_row_y = o; o += 4 * d.rmax;
  d.off_row_attr = o; o += 4 * d.rmax;
  o = align_up(o, 8);
  d.off_penalty = o; o += 8 * d.rmax * d.cmax;
  d.off_peaks = o; o += 8 * d.pmax;
  d.off_occ = o; o += d.rmax * d.cmax;
  o = align_up(o, 4);
  d.off_goals = o; o += 8 * d.pmax;
  d.off_lookup = o; o += 4 * 2 * d.rmax;
  d.record_bytes = align_up(o, 16);
  d.wr = (float)((double)c.mw / (double)c.W);              // python float ratio cast to fp32 by torch
  d.hr = (float)((double)c.mh / (double)c.H);
  d.sx = (float)c.mw / (float)c.W;                         // ATen: static_cast<float>(in) / out
  d.sy = (float)c.mh / (float)c.H;
  memset(&L, 0, sizeof(L));
  L.record_bytes = d.record_bytes; L.rmax = d.rmax; L.cmax = d.cmax; L.pmax = d.pmax;
  L.off_header = 0; L.off_row_y = d.off_row_y; L.off_row_attr = d.off_row_attr; L.off_penalty = d.off_penalty;
  L.off_peaks = d.off_peaks; L.off_occ = d.off_occ; L.lat_rows = d.lat_rows; L.lat_cols = d.lat_cols;
  L.off_goals = d.off_goals; L.off_lookup = d.off_lookup; L.lookup_rows = 2 * d.rmax;
  L.algorithmic_bytes_per_frame_n1 = 4 * c.K * c.mh * c.mw + 4 * c.K + 16 + c.H * c.W + d.rmax * d.cmax * 9 + 64;
  return true;
}

extern "C" int va_abi_version(void) { return VA_ABI_VERSION; }

extern "C" int va_create(va_ctx** out, const va_config* cfg) {
  if (!out || !cfg) { set_err(nullptr, "va_create: null argument"); return VA_ERR_INVALID; }
  *out = nullptr;
  va_ctx* c = new (std::nothrow) va_ctx();
  if (!c) { set_err(nullptr, "out of host memory"); return VA_ERR_INVALID; }
  memset(c, 0, sizeof(*c));
  c->cfg = *cfg;
  char why[256];
  if (!compute_dims(*cfg, c->d, c->layout, why, sizeof(why))) {
    set_err(nullptr, "va_create: %s", why);
    delete c;
    return VA_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    set_err(nullptr, "va_create: no usable CUDA device %d (%s); this library has no CPU fallback", cfg->device,
            e == cudaSuccess ? "device ordinal out of range" : cudaGetErrorString(e));
    delete c;
    return VA_ERR_CUDA;
  }
#define VA_CREATE_CUDA(call)                                                      \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      set_err(nullptr, "va_create: %s failed: %s", #call, cudaGetErrorString(e__)); \
      va_destroy(c);                                                              \
      return VA_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)
  DeviceGuard guard__(cfg->device);
  VA_CREATE_CUDA(guard__.err);
  const Dims& d = c->d;
  const size_t ns = (size_t)cfg->max_batch * d.max_n;
  VA_CREATE_CUDA(cudaMalloc(&c->scratch.stats, ns * sizeof(InstStats)));
  VA_CREATE_CUDA(cudaMalloc(&c->scratch.lattice, ns * d.lat_rows * d.lat_words * sizeof(unsigned)));
  VA_CREATE_CUDA(launch_init_scratch(d, cfg->max_batch, c->scratch.stats, c->scratch.lattice, 0));
  {
    // contour step (tail kernel, va_contour_core.h): per-row summaries and the general path's global slabs
    int sms = 0;
    VA_CREATE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    c->num_sms = sms;
    c->d.num_sms = sms;
    VA_CREATE_CUDA(cudaMalloc(&c->scratch.rowsum, ns * d.H * d.nblk * sizeof(uint32_t)));
    VA_CREATE_CUDA(cudaMemset(c->scratch.rowsum, 0, ns * d.H * d.nblk * sizeof(uint32_t)));
    // run capacity: a 4x-upsampled mask row has at most W / 8 runs (one sign change per proto cell)
    int cap = (int)(((size_t)d.H * d.W) / 8);
    if (cap < 4096) cap = 4096;
    c->scratch.cc_cap = cap;
    // one slab per frame of a batch, at most two per SM (a frame's tail CTA takes the slab's lock only when the batch
    // has more frames than slabs)
    c->scratch.nslab = cfg->max_batch < 2 * sms ? cfg->max_batch : 2 * sms;
    c->scratch.cc_slab_bytes = (contour_slab_bytes(d, cap) + 255) & ~(size_t)255;
    VA_CREATE_CUDA(cudaMalloc(&c->scratch.cc_slab, c->scratch.cc_slab_bytes * c->scratch.nslab));
    VA_CREATE_CUDA(cudaMalloc(&c->scratch.slab_lock, sizeof(int) * c->scratch.nslab));
    VA_CREATE_CUDA(cudaMemset(c->scratch.slab_lock, 0, sizeof(int) * c->scratch.nslab));
  }
  // logits scratch for the CUDA-core path: keep one chunk (<= ~48 MB) so that it stays L2-resident
  const size_t per_frame = (size_t)d.max_n * d.mh * d.mw * sizeof(float);
  int chunk = (int)((48u << 20) / per_frame);
  if (chunk < 1) chunk = 1;
  if (chunk > cfg->max_batch) chunk = cfg->max_batch;
  c->logits_chunk = chunk;
  VA_CREATE_CUDA(cudaMalloc(&c->scratch.logits, per_frame * chunk));
  {
    // quotient table for the penalty ratios (PenaltyCalculator.py:98-110): positions and run lengths are small
    // integers in cell units, and (double)m / (double)den computed here is the correctly rounded quotient the
    // reference's pixel-unit division yields (same rational) - the kernel looks it up instead of dividing
    const int N = (d.cmax > 2 * d.rmax) ? d.cmax : 2 * d.rmax;
    std::vector<double> tab((size_t)(N + 1) * (N + 1));
    for (int m = 0; m <= N; ++m)
      for (int den = 0; den <= N; ++den) tab[(size_t)m * (N + 1) + den] = den ? (double)m / (double)den : 0.5;
    double* dev = nullptr;
    VA_CREATE_CUDA(cudaMalloc(&dev, tab.size() * sizeof(double)));
    c->d.ratio = dev;
    c->d.ratio_n = N;
    VA_CREATE_CUDA(cudaMemcpy(dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  c->plan = nullptr;
  if (!(cfg->flags & VA_CFG_NO_TENSOR_CORE)) {
    char perr[256] = "";
    c->plan = fused_plan_create(d, cfg->device, perr, sizeof(perr));
    if (!c->plan) snprintf(c->err, sizeof(c->err), "tensor-core plan unavailable: %s", perr);
  }
  VA_CREATE_CUDA(cudaDeviceSynchronize());
  *out = c;
  return VA_OK;
}

// releases whatever part of the host pipeline exists (also after a half-finished host_pipeline_init)
static void host_pipeline_destroy(va_ctx* c) {
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->d_protos[i]); cudaFree(c->d_coefs[i]); cudaFree(c->d_boxes[i]); cudaFree(c->d_counts[i]);
    cudaFree(c->d_records[i]); cudaFree(c->d_masks[i]);
    cudaFree(c->d_protos_h[i]); c->d_protos_h[i] = nullptr;
    c->d_protos[i] = c->d_coefs[i] = c->d_boxes[i] = nullptr; c->d_counts[i] = nullptr; c->d_records[i] = c->d_masks[i] = nullptr;
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    c->ev_in[i] = c->ev_done[i] = c->ev_out[i] = nullptr;
  }
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_compute) cudaStreamDestroy(c->s_compute);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  c->s_in = c->s_compute = c->s_out = nullptr;
  c->host_ready = false;
}

extern "C" void va_destroy(va_ctx* c) {
  if (!c) return;
  DeviceGuard guard__(c->cfg.device);
  host_pipeline_destroy(c);
  if (c->plan) fused_plan_destroy(c->plan);
  cudaFree(c->scratch.stats);
  cudaFree(c->scratch.lattice);
  cudaFree(c->scratch.logits);
  cudaFree(c->scratch.rowsum);
  cudaFree(c->scratch.bits);
  cudaFree(c->scratch.cc_slab);
  cudaFree(c->scratch.slab_lock);
  cudaFree(const_cast<double*>(c->d.ratio));
  if (c->prof_ev[0][0])
    for (int i = 0; i < kProfMax; ++i)
      for (int j = 0; j < 3; ++j) cudaEventDestroy(c->prof_ev[i][j]);
  delete c;
}

extern "C" const char* va_last_error(const va_ctx* c) { return c ? c->err : g_create_err; }

extern "C" int va_get_layout(const va_ctx* c, va_layout* out) {
  if (!c || !out) return VA_ERR_INVALID;
  *out = c->layout;
  return VA_OK;
}

extern "C" int va_profile_enable(va_ctx* c, int on) {
  if (!c) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  if (on && !c->prof_ev[0][0]) {
    for (int i = 0; i < kProfMax; ++i)
      for (int j = 0; j < 3; ++j) VA_CUDA(c, cudaEventCreate(&c->prof_ev[i][j]));
  }
  c->prof_on = on > 0 ? on : 0;
  c->prof_count = 0;
  c->prof_calls = 0;
  return VA_OK;
}

extern "C" int va_profile_read(va_ctx* c, float* assemble_ms, float* tail_ms, int32_t* calls) {
  if (!c || !assemble_ms || !tail_ms || !calls) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  float a = 0.f, t = 0.f;
  for (int i = 0; i < c->prof_count; ++i) {
    float ms = 0.f;
    VA_CUDA(c, cudaEventSynchronize(c->prof_ev[i][2]));
    VA_CUDA(c, cudaEventElapsedTime(&ms, c->prof_ev[i][0], c->prof_ev[i][1]));
    a += ms;
    VA_CUDA(c, cudaEventElapsedTime(&ms, c->prof_ev[i][1], c->prof_ev[i][2]));
    t += ms;
  }
  *assemble_ms = a; *tail_ms = t; *calls = c->prof_count;
  c->prof_count = 0;
  return VA_OK;
}

extern "C" int va_nms(va_ctx* c, const float* pred, int32_t A, const va_nms_params* prm, int32_t B, float* coefs_out,
                      float* boxes_out, float* conf_out, int32_t* cls_out, int32_t* counts_out, void* stream) {
  if (!c) return VA_ERR_INVALID;
  if (!pred || !prm || !coefs_out || !boxes_out || !counts_out) { set_err(c, "va_nms: null pointer"); return VA_ERR_INVALID; }
  if (B < 0 || B > c->cfg.max_batch) { set_err(c, "batch %d exceeds max_batch %d", B, c->cfg.max_batch); return VA_ERR_CAPACITY; }
  if (A < 1 || prm->nc < 1 || prm->max_det < 1 || prm->max_det > c->cfg.max_n) {
    set_err(c, "va_nms: need A >= 1, nc >= 1, 1 <= max_det <= max_n (%d)", c->cfg.max_n);
    return VA_ERR_INVALID;
  }
  if (B == 0) return VA_OK;
  VA_ON_DEVICE(c);
  const float off = prm->agnostic ? 0.f : (float)prm->max_wh;
  constexpr int kMaxNms = 30000;      // ops.non_max_suppression's max_nms default (ops.py:225); predict does not change it
  VA_CUDA(c, launch_nms(pred, A, prm->nc, c->cfg.K, prm->conf_thres, prm->iou_thres, off, prm->max_det, c->cfg.max_n, kMaxNms, B,
                        coefs_out, boxes_out, conf_out, cls_out, counts_out, (cudaStream_t)stream));
  c->last_launches = 1;
  return VA_OK;
}

extern "C" int va_scale_boxes(va_ctx* c, const float* boxes, const int32_t* counts, int32_t B, int32_t img1_h, int32_t img1_w,
                              int32_t img0_h, int32_t img0_w, float* boxes_out, void* stream) {
  if (!c) return VA_ERR_INVALID;
  if (!boxes || !counts || !boxes_out) { set_err(c, "va_scale_boxes: null pointer"); return VA_ERR_INVALID; }
  if (B < 0 || B > c->cfg.max_batch) { set_err(c, "batch %d exceeds max_batch %d", B, c->cfg.max_batch); return VA_ERR_CAPACITY; }
  if (img1_h < 1 || img1_w < 1 || img0_h < 1 || img0_w < 1) { set_err(c, "va_scale_boxes: image shapes must be positive"); return VA_ERR_INVALID; }
  if (B == 0) return VA_OK;
  VA_ON_DEVICE(c);
  // ops.py:158-163 in Python arithmetic: doubles, round() = round half to even
  const double gain = std::min((double)img1_h / img0_h, (double)img1_w / img0_w);
  const double pad_x = std::nearbyint((img1_w - img0_w * gain) / 2 - 0.1), pad_y = std::nearbyint((img1_h - img0_h * gain) / 2 - 0.1);
  VA_CUDA(c, launch_scale_boxes(boxes, counts, c->cfg.max_n, B, (float)pad_x, (float)pad_y, (float)gain, (float)img0_w, (float)img0_h,
                                boxes_out, (cudaStream_t)stream));
  c->last_launches = 1;
  return VA_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU record sink: peer-mapped buffers + system-scope flags
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == VA_IPC_HANDLE_BYTES, "IPC handle size");

__global__ void signal_kernel(int* flag, int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void wait_flags_kernel(const int* flags, int n, int value) {
  for (int i = threadIdx.x; i < n; i += (int)blockDim.x) {
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (v < value) __nanosleep(200);
    } while (v < value);
  }
  __threadfence_system();
}

extern "C" int va_peer_alloc(va_ctx* c, uint64_t bytes, void** dptr, uint8_t handle[VA_IPC_HANDLE_BYTES]) {
  if (!c || !dptr || !handle || bytes == 0) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  void* p = nullptr;
  VA_CUDA(c, cudaMalloc(&p, bytes));
  VA_CUDA(c, cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_err(c, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return VA_ERR_CUDA; }
  memcpy(handle, &h, sizeof(h));
  VA_CUDA(c, cudaDeviceSynchronize());
  *dptr = p;
  return VA_OK;
}
extern "C" int va_peer_open(va_ctx* c, const uint8_t handle[VA_IPC_HANDLE_BYTES], void** dptr) {
  if (!c || !dptr || !handle) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  VA_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr = p;
  return VA_OK;
}
extern "C" int va_peer_close(va_ctx* c, void* dptr) {
  if (!c || !dptr) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  VA_CUDA(c, cudaIpcCloseMemHandle(dptr));
  return VA_OK;
}
extern "C" int va_peer_free(va_ctx* c, void* dptr) {
  if (!c || !dptr) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  VA_CUDA(c, cudaFree(dptr));
  return VA_OK;
}
extern "C" int va_peer_put(va_ctx* c, void* dst, const void* src, uint64_t bytes, void* stream) {
  if (!c || !dst || !src) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  VA_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return VA_OK;
}
extern "C" int va_signal(va_ctx* c, int32_t* flag, int32_t value, void* stream) {
  if (!c || !flag) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
  VA_CUDA(c, cudaGetLastError());
  return VA_OK;
}
extern "C" int va_wait_flags(va_ctx* c, const int32_t* flags, int32_t n, int32_t value, void* stream) {
  if (!c || !flags || n < 1) return VA_ERR_INVALID;
  VA_ON_DEVICE(c);
  wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, value);
  VA_CUDA(c, cudaGetLastError());
  return VA_OK;
}

extern "C" int va_last_launch_count(const va_ctx* c) { return c ? c->last_launches : 0; }
extern "C" int va_uses_tensor_core(const va_ctx* c) { return (c && c->plan) ? 1 : 0; }

static int check_batch(va_ctx* c, int B, const void* a, const void* b, const void* cc, const void* dd) {
  if (!c) return VA_ERR_INVALID;
  if (B == 0) return VA_OK;
  if (!a || !b || !cc || !dd) { set_err(c, "null input pointer"); return VA_ERR_INVALID; }
  if (B < 0 || B > c->cfg.max_batch) { set_err(c, "batch %d exceeds max_batch %d", B, c->cfg.max_batch); return VA_ERR_CAPACITY; }
  return VA_OK;
}

// mask assembly for frames [0, B): tcgen05 path when available, else logits + upsample in
// L2-sized chunks.  Leaves stats / lattice filled for the tail.
static int assemble(va_ctx* c, const float* protos, const float* coefs, const float* boxes, const int* counts, int B,
                    uint8_t* masks, float* logits_out, cudaStream_t st) {
  const Dims& d = c->d;
  if (!masks && !c->scratch.bits) {
    // grid-only mode: the contour step still needs the pixels of non-trivial masks - the kernels write them bit-packed
    // (1/8 of the mask bytes) into context scratch instead
    const size_t ns = (size_t)c->cfg.max_batch * d.max_n;
    VA_CUDA(c, cudaMalloc(&c->scratch.bits, ns * d.H * d.bit_words * sizeof(uint32_t)));
  }
  MaskSinks sinks;
  sinks.stats = c->scratch.stats; sinks.lattice = c->scratch.lattice; sinks.rowsum = c->scratch.rowsum;
  sinks.bits = masks ? nullptr : c->scratch.bits;
  if (c->plan) {
    char perr[256] = "";
    cudaError_t e = launch_fused(c->plan, d, protos, coefs, boxes, counts, B, masks, logits_out, sinks, st, perr, sizeof(perr));
    if (e != cudaSuccess) { set_err(c, "fused kernel launch failed: %s %s", cudaGetErrorString(e), perr); return VA_ERR_CUDA; }
    c->last_launches += 1;
    return VA_OK;
  }
  const size_t P = (size_t)d.mh * d.mw;
  const size_t fr_protos = (size_t)d.K * P, fr_coefs = (size_t)d.max_n * d.K, fr_boxes = (size_t)d.max_n * 4;
  const size_t fr_logits = (size_t)d.max_n * P, fr_masks = (size_t)d.max_n * d.H * d.W;
  // L2-sized sub-batches of equal size (32 frames with room for 14: 11 + 11 + 10 rather than 14 + 14 + 4)
  const int nsub = (B + c->logits_chunk - 1) / c->logits_chunk;
  const int step = logits_out ? B : (B + nsub - 1) / nsub;
  for (int b0 = 0; b0 < B; b0 += step) {
    const int nb = (B - b0 < step) ? B - b0 : step;
    float* lg = logits_out ? logits_out + b0 * fr_logits : c->scratch.logits;
    VA_CUDA(c, launch_logits(d, protos + b0 * fr_protos, coefs + b0 * fr_coefs, boxes + b0 * fr_boxes, counts + b0, nb, lg, st));
    MaskSinks sub = sinks;
    sub.stats += (size_t)b0 * d.max_n;
    sub.lattice += (size_t)b0 * d.max_n * d.lat_rows * d.lat_words;
    sub.rowsum += (size_t)b0 * d.max_n * d.H * d.nblk;
    if (sub.bits) sub.bits += (size_t)b0 * d.max_n * d.H * d.bit_words;
    VA_CUDA(c, launch_upsample(d, lg, boxes + b0 * fr_boxes, counts + b0, nb, masks ? masks + b0 * fr_masks : nullptr, sub, st));
    c->last_launches += 2;
  }
  return VA_OK;
}

extern "C" int va_assemble_masks(va_ctx* c, const float* protos, const float* coefs, const float* boxes,
                                 const int32_t* counts, int32_t B, uint8_t* masks_out, float* logits_out, void* stream) {
  int rc = check_batch(c, B, protos, coefs, boxes, counts);
  if (rc != VA_OK) return rc;
  if (B == 0) return VA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VA_ON_DEVICE(c);
  c->last_launches = 0;
  rc = assemble(c, protos, coefs, boxes, counts, B, masks_out, logits_out, st);
  if (rc != VA_OK) return rc;
  // masks-only call: the reductions are not consumed by a tail, reset them
  VA_CUDA(c, launch_init_scratch(c->d, B, c->scratch.stats, c->scratch.lattice, st));
  VA_CUDA(c, cudaMemsetAsync(c->scratch.rowsum, 0, (size_t)B * c->d.max_n * c->d.H * c->d.nblk * sizeof(uint32_t), st));
  c->last_launches += 1;
  return VA_OK;
}

extern "C" int va_run_fused(va_ctx* c, const float* protos, const float* coefs, const float* boxes,
                            const int32_t* counts, int32_t B, uint8_t* masks_out, uint8_t* records_out, void* stream) {
  int rc = check_batch(c, B, protos, coefs, boxes, counts);
  if (rc != VA_OK) return rc;
  if (B == 0) return VA_OK;
  if (!records_out) { set_err(c, "records_out is null"); return VA_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  VA_ON_DEVICE(c);
  c->last_launches = 0;
  const bool prof = c->prof_on && (c->prof_calls++ % c->prof_on) == 0 && c->prof_count < kProfMax;
  cudaEvent_t* ev = prof ? c->prof_ev[c->prof_count] : nullptr;
  if (prof) VA_CUDA(c, cudaEventRecord(ev[0], st));
  rc = assemble(c, protos, coefs, boxes, counts, B, masks_out, nullptr, st);
  if (rc != VA_OK) return rc;
  if (prof) VA_CUDA(c, cudaEventRecord(ev[1], st));
  VA_CUDA(c, launch_tail(c->d, counts, B, c->scratch, masks_out, nullptr, nullptr, records_out, st));
  if (prof) { VA_CUDA(c, cudaEventRecord(ev[2], st)); c->prof_count++; }
  c->last_launches += 1;
  return VA_OK;
}

extern "C" int va_mask_to_records(va_ctx* c, const uint8_t* masks, const int32_t* counts, int32_t B,
                                  const int32_t* rects, const int32_t* sel, uint8_t* records_out, void* stream) {
  if (!c) return VA_ERR_INVALID;
  if (!masks || !counts || !records_out) { set_err(c, "null pointer"); return VA_ERR_INVALID; }
  if (B < 0 || B > c->cfg.max_batch) { set_err(c, "batch %d exceeds max_batch %d", B, c->cfg.max_batch); return VA_ERR_CAPACITY; }
  if (B == 0) return VA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VA_ON_DEVICE(c);
  c->last_launches = 0;
  MaskSinks sinks;
  sinks.stats = c->scratch.stats; sinks.lattice = c->scratch.lattice; sinks.rowsum = c->scratch.rowsum; sinks.bits = nullptr;
  VA_CUDA(c, launch_mask_stats(c->d, masks, counts, B, sinks, st));
  VA_CUDA(c, launch_tail(c->d, counts, B, c->scratch, masks, rects, sel, records_out, st));
  c->last_launches = 2;
  return VA_OK;
}

extern "C" int va_grid_to_penalty_peaks(va_ctx* c, const va_grid_input* hdr, const int32_t* row_y,
                                        const int32_t* row_attr, const uint8_t* occ, const int32_t* plane_y,
                                        const uint8_t* plane_occ, int32_t B, uint8_t* records_out, void* stream) {
  if (!c) return VA_ERR_INVALID;
  if (!hdr || !row_y || !row_attr || !occ || !records_out) { set_err(c, "null pointer"); return VA_ERR_INVALID; }
  if (B < 0 || B > c->cfg.max_batch) { set_err(c, "batch %d exceeds max_batch %d", B, c->cfg.max_batch); return VA_ERR_CAPACITY; }
  if (B == 0) return VA_OK;
  VA_ON_DEVICE(c);
  VA_CUDA(c, launch_grid_mode(c->d, hdr, row_y, row_attr, occ, plane_y, plane_occ, B, records_out, (cudaStream_t)stream));
  c->last_launches = 1;
  return VA_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer pipeline
// ---------------------------------------------------------------------------------------------
static int host_pipeline_init_impl(va_ctx* c);
static int host_pipeline_init(va_ctx* c) {
  if (c->host_ready) return VA_OK;
  const int rc = host_pipeline_init_impl(c);
  if (rc != VA_OK) host_pipeline_destroy(c);       // nothing of a half-built pipeline survives
  return rc;
}
static int host_pipeline_init_impl(va_ctx* c) {
  const Dims& d = c->d;
  const size_t P = (size_t)d.mh * d.mw;
  const size_t fr_in = (size_t)d.K * P * 4;
  int chunk = (int)((96u << 20) / fr_in);          // ~96 MB of prototypes per chunk
  if (chunk < 1) chunk = 1;
  if (chunk > c->cfg.max_batch) chunk = c->cfg.max_batch;
  c->host_chunk = chunk;
  VA_CUDA(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  VA_CUDA(c, cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
  VA_CUDA(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    VA_CUDA(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    VA_CUDA(c, cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    VA_CUDA(c, cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    VA_CUDA(c, cudaMalloc(&c->d_protos[i], fr_in * chunk));
    VA_CUDA(c, cudaMalloc(&c->d_coefs[i], (size_t)d.max_n * d.K * 4 * chunk));
    VA_CUDA(c, cudaMalloc(&c->d_boxes[i], (size_t)d.max_n * 16 * chunk));
    VA_CUDA(c, cudaMalloc(&c->d_counts[i], sizeof(int) * chunk));
    VA_CUDA(c, cudaMalloc(&c->d_records[i], (size_t)d.record_bytes * chunk));
    c->d_masks[i] = nullptr;
    c->d_protos_h[i] = nullptr;
  }
  c->host_ready = true;
  return VA_OK;
}

// protos.float() (ops.py:724) of fp16 prototypes: exact widening, 8 values per thread
__global__ void half_to_float_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, size_t n8) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n8; t += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[t];
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), e = __half22float2(h[3]);
    dst[2 * t] = make_float4(a.x, a.y, b.x, b.y);
    dst[2 * t + 1] = make_float4(c.x, c.y, e.x, e.y);
  }
}

static int run_fused_host_impl(va_ctx* c, const void* h_protos_any, bool f16, const float* h_coefs, const float* h_boxes,
                               const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out);

extern "C" int va_run_fused_host(va_ctx* c, const float* h_protos, const float* h_coefs, const float* h_boxes,
                                 const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out) {
  return run_fused_host_impl(c, h_protos, false, h_coefs, h_boxes, h_counts, B, h_masks_out, h_records_out);
}

extern "C" int va_run_fused_host_f16(va_ctx* c, const uint16_t* h_protos_f16, const float* h_coefs, const float* h_boxes,
                                     const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out) {
  return run_fused_host_impl(c, h_protos_f16, true, h_coefs, h_boxes, h_counts, B, h_masks_out, h_records_out);
}

static int run_fused_host_impl(va_ctx* c, const void* h_protos_any, bool f16, const float* h_coefs, const float* h_boxes,
                               const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out) {
  const float* h_protos = static_cast<const float*>(h_protos_any);
  const uint16_t* h_protos_h = static_cast<const uint16_t*>(h_protos_any);
  int rc = check_batch(c, B, h_protos, h_coefs, h_boxes, h_counts);
  if (rc != VA_OK) return rc;
  if (!h_records_out) { set_err(c, "h_records_out is null"); return VA_ERR_INVALID; }
  if (B == 0) return VA_OK;
  VA_ON_DEVICE(c);
  rc = host_pipeline_init(c);
  if (rc != VA_OK) return rc;
  const Dims& d = c->d;
  const size_t P = (size_t)d.mh * d.mw;
  const size_t fr_protos = (size_t)d.K * P, fr_coefs = (size_t)d.max_n * d.K, fr_boxes = (size_t)d.max_n * 4;
  const size_t fr_masks = (size_t)d.max_n * d.H * d.W;
  const int chunk = c->host_chunk;
  if (h_masks_out) {
    for (int i = 0; i < 2; ++i)
      if (!c->d_masks[i]) VA_CUDA(c, cudaMalloc(&c->d_masks[i], fr_masks * chunk));
  }
  if (f16) {
    if ((fr_protos & 7) != 0) { set_err(c, "fp16 prototypes need K*mh*mw to be a multiple of 8"); return VA_ERR_INVALID; }
    for (int i = 0; i < 2; ++i)
      if (!c->d_protos_h[i]) VA_CUDA(c, cudaMalloc(&c->d_protos_h[i], fr_protos * 2 * chunk));
  }
  int launches = 0;
  int k = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++k) {
    const int nb = (B - b0 < chunk) ? B - b0 : chunk;
    const int s = k & 1;
    if (k >= 2) VA_CUDA(c, cudaStreamWaitEvent(c->s_in, c->ev_done[s], 0));     // input slot consumed
    if (f16)
      VA_CUDA(c, cudaMemcpyAsync(c->d_protos_h[s], h_protos_h + b0 * fr_protos, fr_protos * 2 * nb, cudaMemcpyHostToDevice, c->s_in));
    else
      VA_CUDA(c, cudaMemcpyAsync(c->d_protos[s], h_protos + b0 * fr_protos, fr_protos * 4 * nb, cudaMemcpyHostToDevice, c->s_in));
    VA_CUDA(c, cudaMemcpyAsync(c->d_coefs[s], h_coefs + b0 * fr_coefs, fr_coefs * 4 * nb, cudaMemcpyHostToDevice, c->s_in));
    VA_CUDA(c, cudaMemcpyAsync(c->d_boxes[s], h_boxes + b0 * fr_boxes, fr_boxes * 4 * nb, cudaMemcpyHostToDevice, c->s_in));
    VA_CUDA(c, cudaMemcpyAsync(c->d_counts[s], h_counts + b0, sizeof(int) * nb, cudaMemcpyHostToDevice, c->s_in));
    VA_CUDA(c, cudaEventRecord(c->ev_in[s], c->s_in));
    VA_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_in[s], 0));
    if (k >= 2) VA_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_out[s], 0)); // output slot drained
    if (f16) {
      const size_t n8 = fr_protos * nb / 8;
      half_to_float_kernel<<<c->d.num_sms * 8, 256, 0, c->s_compute>>>(reinterpret_cast<const uint4*>(c->d_protos_h[s]),
                                                                       reinterpret_cast<float4*>(c->d_protos[s]), n8);
      VA_CUDA(c, cudaGetLastError());
      ++launches;
    }
    c->last_launches = 0;
    rc = assemble(c, c->d_protos[s], c->d_coefs[s], c->d_boxes[s], c->d_counts[s], nb,
                  h_masks_out ? c->d_masks[s] : nullptr, nullptr, c->s_compute);
    if (rc != VA_OK) {               // copies of earlier chunks may still reference the caller's buffers: drain first
      cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_compute); cudaStreamSynchronize(c->s_out);
      return rc;
    }
    VA_CUDA(c, launch_tail(d, c->d_counts[s], nb, c->scratch, h_masks_out ? c->d_masks[s] : nullptr, nullptr, nullptr,
                           c->d_records[s], c->s_compute));
    launches += c->last_launches + 1;
    VA_CUDA(c, cudaEventRecord(c->ev_done[s], c->s_compute));
    VA_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_done[s], 0));
    VA_CUDA(c, cudaMemcpyAsync(h_records_out + (size_t)b0 * d.record_bytes, c->d_records[s], (size_t)d.record_bytes * nb,
                               cudaMemcpyDeviceToHost, c->s_out));
    if (h_masks_out)
      VA_CUDA(c, cudaMemcpyAsync(h_masks_out + b0 * fr_masks, c->d_masks[s], fr_masks * nb, cudaMemcpyDeviceToHost, c->s_out));
    VA_CUDA(c, cudaEventRecord(c->ev_out[s], c->s_out));
  }
  VA_CUDA(c, cudaStreamSynchronize(c->s_out));
  VA_CUDA(c, cudaStreamSynchronize(c->s_compute));
  c->last_launches = launches;
  return VA_OK;
}
