// Shared definitions for the sm_100a kernels of libva_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vision_assist_b200.h"

namespace va {

constexpr int kMaxInst = 32;   // instance capacity (va_config.max_n upper bound)
constexpr int kProtoK = 32;    // prototypes

// Per-(frame, instance) reduction of the binary mask, produced by the mask kernels and consumed
// (then reset to this state) by the tail kernel.
struct InstStats {
  unsigned int area;   // pixels set
  int minx, miny;      // init INT_MAX
  int maxx, maxy;      // init -1
  int euler4;          // 4 * Euler number partial sums (optional)
  int pad0, pad1;
};
static_assert(sizeof(InstStats) == 32, "InstStats is 32 B");

// Everything a kernel needs to know about the problem geometry (passed by value).
struct Dims {
  int H, W, mh, mw, K, max_n, gs;
  int lat_rows, lat_cols, lat_words;   // cell-centre lattice: point (gs*lx+gs/2, gs*ly+gs/2)
  int nblk;                            // 128-pixel blocks per row (per-row summaries, va_contour_core.h)
  int bit_words;                       // 32-pixel words per row of the bit-packed masks (grid-only mode)
  int plane_rows;                      // ceil(H/gs): rows of the grid_lookup plane
  int rmax, cmax, pmax, cwords;        // record capacity; cwords = ceil(cmax/32)
  int record_bytes, off_row_y, off_row_attr, off_penalty, off_peaks, off_occ, off_goals, off_lookup;
  int band_start;                      // FrameProcessor.py:126-127 starting_y
  int flags;                           // VA_CFG_*
  int num_sms;                         // of the context's device
  float wr, hr;                        // fl32(mw/W), fl32(mh/H): box scale, ops.py:725-732
  float sx, sy;                        // fl32(mw)/W, fl32(mh)/H: bilinear scales (ATen area_pixel_compute_scale)
  const double* ratio;                 // [ratio_n + 1][ratio_n + 1]: ratio[m][den] = (double)m / (double)den (host-built, exact)
  int ratio_n;
};

struct Scratch {
  InstStats* stats;        // [max_batch][max_n]
  unsigned int* lattice;   // [max_batch][max_n][lat_rows][lat_words]
  float* logits;           // [max_batch][max_n][mh][mw] (CUDA-core path only)
  uint32_t* rowsum;        // [max_batch][max_n][H][nblk] per-(row, 128 px block) summaries of the masks
  uint32_t* bits;          // [max_batch][max_n][H][bit_words] bit-packed masks, written INSTEAD of the u8 masks in
                           // grid-only mode (allocated on first use)
  unsigned char* cc_slab;  // [nslab][cc_slab_bytes] global scratch of the contour step's general path (tail kernel)
  size_t cc_slab_bytes;
  int cc_cap;              // run capacity per instance
  int nslab;
  int* slab_lock;          // [nslab]
};

// Where the mask kernels leave their by-products (besides the u8 masks).
struct MaskSinks {
  InstStats* stats;
  unsigned int* lattice;
  uint32_t* rowsum;
  uint32_t* bits;          // nullptr when the u8 masks are written
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Python floor division for possibly negative numerators (b > 0).
__host__ __device__ inline int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------------
// launchers (one per translation unit)
// ---------------------------------------------------------------------------------------------
// CUDA-core contraction + crop: logits[b][i][mh][mw] (ops.py:724-734)
cudaError_t launch_logits(const Dims& d, const float* protos, const float* coefs, const float* boxes,
                          const int* counts, int B, float* logits, cudaStream_t st);
// bilinear upsample + threshold (+ stats / lattice) from proto-resolution logits (ops.py:736-737)
cudaError_t launch_upsample(const Dims& d, const float* logits, const float* boxes, const int* counts, int B, uint8_t* masks,
                            const MaskSinks& sinks, cudaStream_t st);
// stats / lattice / row summaries from caller-provided binary masks
cudaError_t launch_mask_stats(const Dims& d, const uint8_t* masks, const int* counts, int B, const MaskSinks& sinks,
                              cudaStream_t st);
size_t contour_slab_bytes(const Dims& d, int cap);
cudaError_t launch_init_scratch(const Dims& d, int max_batch, InstStats* stats, unsigned int* lattice,
                                cudaStream_t st);
// contour step (per instance the polygon the reference keeps: doubled contourArea, bounding box, lattice samples of
// its fillPoly raster) -> selection -> grid -> penalties -> peaks -> record; resets stats / lattice / row summaries.
// masks may be nullptr (then scratch.bits holds the pixels).
cudaError_t launch_tail(const Dims& d, const int* counts, int B, const Scratch& sc, const uint8_t* masks, const int* rects,
                        const int* sel, uint8_t* records, cudaStream_t st);
cudaError_t launch_grid_mode(const Dims& d, const va_grid_input* hdr, const int* row_y, const int* row_attr,
                             const uint8_t* occ, const int* plane_y, const uint8_t* plane_occ, int B,
                             uint8_t* records, cudaStream_t st);
size_t tail_smem_bytes(const Dims& d);

// candidate filter + NMS on raw head output (va_nms.cu); counts_out[b] < 0: more than the 512 candidates it holds
cudaError_t launch_nms(const float* pred, int A, int nc, int nm, float conf_thres, float iou_thres, float class_offset,
                       int max_det, int max_n, int max_nms, int B, float* coefs_out, float* boxes_out, float* conf_out, int* cls_out,
                       int* counts_out, cudaStream_t st);
cudaError_t launch_scale_boxes(const float* boxes, const int* counts, int max_n, int B, float pad_x, float pad_y, float gain,
                               float w0, float h0, float* out, cudaStream_t st);

// tcgen05 / TMA fused kernel (va_fused_tc.cu)
struct FusedPlan;  // opaque, owned by the context
FusedPlan* fused_plan_create(const Dims& d, int device, char* err, size_t errlen);
void fused_plan_destroy(FusedPlan* p);
cudaError_t launch_fused(FusedPlan* p, const Dims& d, const float* protos, const float* coefs, const float* boxes,
                         const int* counts, int B, uint8_t* masks, float* logits_dbg, const MaskSinks& sinks,
                         cudaStream_t st, char* err, size_t errlen);

}  // namespace va
