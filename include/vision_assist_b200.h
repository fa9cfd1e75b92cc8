/*
 * vision_assist_b200.h - C ABI of libva_sm100.so
 *
 * B200-native (sm_100a) implementation of Vision Assist's per-frame data-parallel stage:
 *
 *   YOLOv8-seg mask assembly   ultralytics.utils.ops.process_mask, vendored in the reference at
 *                              testing/old/segmenting_using_tflite/ops.py:707-737 (+ crop_mask :688-704)
 *   mask -> occupancy grid     FrameProcessor._extract_grid_information     FrameProcessor.py:50-171
 *   penalty map                PenaltyCalculator                            PenaltyCalculator.py:26-142
 *                              (driver FrameProcessor._calculate_penalties  FrameProcessor.py:173-182)
 *   top-edge (protrusion) scan ProtrusionDetector.__call__/_find_peak       ProtrusionDetector.py:38-158,419-535
 *
 * The reference has no FFI layer (it is pure Python); these entry points are what a ctypes
 * binding inside the reference's FrameProcessor / PenaltyCalculator / ProtrusionDetector would
 * call (INTEGRATION.md shows that binding).  Conventions:
 *
 *   - plain C, no exceptions: every call returns VA_OK (0) or a negative VA_ERR_* code and
 *     va_last_error() describes the failure;
 *   - device entry points take raw DEVICE pointers (e.g. torch.Tensor.data_ptr()) and the
 *     caller's CUDA stream (cudaStream_t passed as void*; NULL = legacy default stream); they
 *     enqueue work and return without synchronising;  *_host entry points take HOST pointers,
 *     do their own H2D/D2H copies and return after the results are in host memory;
 *   - the caller owns every buffer; the library owns only the context's scratch memory;
 *   - a context is bound to one device and is not thread-safe (one context per stream);
 *   - there is NO CPU fallback: without a CUDA device va_create fails with VA_ERR_CUDA.
 */
#ifndef VISION_ASSIST_B200_H
#define VISION_ASSIST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VA_ABI_VERSION 2
#if defined(__GNUC__)
#define VA_API __attribute__((visibility("default")))
#else
#define VA_API
#endif

enum {
  VA_OK = 0,
  VA_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  VA_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  VA_ERR_CAPACITY = -3,  /* batch or instance count above what the context was created for */
  VA_ERR_UNSUPPORTED = -4
};

/* va_config.flags */
enum {
  VA_CFG_CHECK_SIMPLE = 1,  /* accepted and ignored since ABI 2: the contour step always classifies the masks
                               (VA_FLAG_NON_SIMPLE) and is exact on all of them */
  VA_CFG_NO_TENSOR_CORE = 2 /* debug: force the CUDA-core (FFMA) contraction instead of tcgen05 */
};

/* va_frame_header.flags */
enum {
  VA_FLAG_EMPTY = 1,       /* no grid: reference returns [] (FrameProcessor.py:99-101, :328-332) */
  VA_FLAG_CENTRE_OOB = 2,  /* reference raises IndexError at FrameProcessor.py:97 (cell centre outside frame) */
  VA_FLAG_LIST_OOB = 4,    /* reference raises IndexError at FrameProcessor.py:163 (negative list index) */
  VA_FLAG_NON_SIMPLE = 8,  /* informational: the selected mask is not a single hole-free 8-connected blob; the record is
                              still exact - it is built from the polygon the reference keeps (external contour with
                              the most points of masks2segments, ops.py:837-859, filled: FrameProcessor.py:85-86) */
  VA_FLAG_OVERFLOW = 16,   /* record capacity exceeded (cannot happen for frames the context was sized for), or more
                              than H*W/8 pixel runs in one mask (cannot happen for 4x-upsampled masks) */
  VA_FLAG_NO_POLYGON = 32  /* the selected instance's mask is empty: masks2segments yields a polygon without points
                              and the reference raises cv2.error in cv2.fillPoly (FrameProcessor.py:86); set with
                              VA_FLAG_EMPTY */
};

typedef struct va_ctx va_ctx;

typedef struct va_config {
  int32_t device;    /* CUDA device ordinal */
  int32_t H, W;      /* frame (= network input) height / width in pixels */
  int32_t mh, mw;    /* prototype map height / width */
  int32_t K;         /* number of prototypes (32) */
  int32_t max_n;     /* instance capacity per frame (<= 32) */
  int32_t gs;        /* grid cell size in pixels (config.py:1 grid_size = 20) */
  int32_t max_batch; /* frames per call capacity */
  int32_t flags;     /* VA_CFG_* */
} va_config;

/* Per-frame result record: one contiguous blob of layout.record_bytes bytes.
 *   header   : va_frame_header (64 B)
 *   row_y    : int32 [rmax]        pixel y of list row k        (Grid.coords.y)
 *   row_attr : int32 [rmax]        Grid.row attribute of row k  (differs from k under the
 *                                  duplicate-row / gap-compression / negative-index quirks)
 *   penalty  : double[rmax][cmax]  Grid.penalty, NaN where the cell is empty (None in the reference)
 *   peaks    : int32 [pmax][2]     (x, y) of ProtrusionDetector's returned Coordinates
 *   occ      : uint8 [rmax][cmax]  bit0 = not Grid.empty, bit1 = Grid.artificial
 *   goals    : int32 [pmax][2]     (list row, column) of the path end cell of every peak: the non-empty list
 *                                  cell closest to it (utils.py:6-32 as called at FrameProcessor.py:238-239)
 *   lookup   : int32 [lookup_rows] for y = ly*gs: record row that owns grid_lookup at y, -1 if none.  This is the
 *                                  A* graph of FrameProcessor._create_graph (:184-207) in implicit form: a non-empty
 *                                  list cell (x, y) has the edge to (x +- gs, y) iff that column exists, and to
 *                                  (x, y +- gs) iff lookup[(y +- gs) / gs] >= 0 (also towards empty cells, :203)
 * Rows [0, n_rows) are FrameProcessor.grids in list order (np_grids = occ & 1); rows
 * [n_rows, n_rows + n_orphans) are rows that are only reachable through grid_lookup.
 * Cell (k, c) has Grid.coords = (x0 + c*gs, row_y[k]) and Grid.col = c. */
typedef struct va_frame_header {
  int32_t flags;      /* VA_FLAG_* */
  int32_t sel;        /* selected instance: largest cv2.contourArea of the kept polygons, first maximum
                         (FrameProcessor.py:72-73); -1 if the frame has no instance */
  int32_t x0, y0;     /* snapped bbox origin (FrameProcessor.py:79-80) */
  int32_t n_cols;     /* C */
  int32_t n_rows;     /* R = len(FrameProcessor.grids) */
  int32_t n_orphans;
  int32_t n_peaks;
  int32_t area;       /* pixels set in the selected instance's mask */
  int32_t n_mask_rows;
  int32_t minx, miny, maxx, maxy; /* cv2.boundingRect of the kept polygon (FrameProcessor.py:76) */
  int32_t contour_area2; /* 2 * cv2.contourArea of the kept polygon (exact integer) */
  int32_t start_cell; /* (list row << 16) | column of the path start cell, -1 if none: the non-empty list cell
                         closest to (W/2, H) - utils.get_closest_grid_to_point as called at FrameProcessor.py:236 */
} va_frame_header;

typedef struct va_layout {
  int32_t record_bytes;
  int32_t rmax, cmax, pmax;
  int32_t off_header, off_row_y, off_row_attr, off_penalty, off_peaks, off_occ;
  int32_t lat_rows, lat_cols; /* cell-centre lattice sampled from the masks */
  int32_t algorithmic_bytes_per_frame_n1; /* SURVEY 8(d) with n = 1 masks written; informational */
  int32_t off_goals, off_lookup, lookup_rows;
} va_layout;

/* Grid-mode input header for va_grid_to_penalty_peaks (PenaltyCalculator / ProtrusionDetector
 * drop-ins and the reference's *_grids.npy fixtures).  All coordinates are multiples of gs. */
typedef struct va_grid_input {
  int32_t x0;       /* Grid.coords.x of column 0 */
  int32_t n_cols;   /* C <= cmax */
  int32_t n_rows;   /* list rows R <= rmax */
  int32_t n_plane;  /* number of grid_lookup rows given (0 = derive the lookup from the list rows,
                       later rows overriding earlier ones with the same y) */
  int32_t use_easy; /* 1 = _pre_compute_easy_segments(np_grids, grids) as FrameProcessor does;
                       0 = empty np_grids (run_on_main.py fixture path: pure traversal) */
  int32_t reserved[3];
} va_grid_input;

VA_API int va_abi_version(void);

/* Create / destroy a context.  Replaces: FrameProcessor.__init__ (FrameProcessor.py:29-42) state. */
VA_API int va_create(va_ctx** out, const va_config* cfg);
VA_API void va_destroy(va_ctx* ctx);
VA_API const char* va_last_error(const va_ctx* ctx); /* ctx may be NULL: error of the last failed va_create */
VA_API int va_get_layout(const va_ctx* ctx, va_layout* out);

/* Mask assembly only.  Replaces ops.process_mask(protos[b], coefs[b,:n], boxes[b,:n], (H,W),
 * upsample=True) for every frame b (ops.py:707-737).
 *   protos [B][K][mh][mw] f32, coefs [B][max_n][K] f32, boxes [B][max_n][4] f32 xyxy in (H,W)
 *   pixels, counts [B] i32 (instances per frame, <= max_n).
 *   masks_out  [B][max_n][H][W] u8 in {0,1}, or NULL
 *   logits_out [B][max_n][mh][mw] f32 (cropped proto-resolution logits, ops.py:734), or NULL */
VA_API int va_assemble_masks(va_ctx* ctx, const float* protos, const float* coefs, const float* boxes,
                      const int32_t* counts, int32_t B, uint8_t* masks_out, float* logits_out,
                      void* stream);

/* Whole path in one call: mask assembly -> kept polygon per instance -> grid -> penalties -> peaks.
 * Replaces model.predict()'s process_mask + Results.masks.xy (masks2segments) + FrameProcessor.
 * _extract_grid_information + _calculate_penalties + ProtrusionDetector.__call__ (FrameProcessor.py:322-341).
 *   masks_out may be NULL ("grid-only" mode: the u8 masks never touch HBM; a bit-packed copy, 1/8 of the bytes,
 *   goes to context scratch for the contour step).
 *   records_out [B][record_bytes] */
VA_API int va_run_fused(va_ctx* ctx, const float* protos, const float* coefs, const float* boxes,
                 const int32_t* counts, int32_t B, uint8_t* masks_out, uint8_t* records_out,
                 void* stream);

/* Same as va_run_fused with HOST buffers: copies inputs H2D in chunks overlapped with compute,
 * copies records (and masks when h_masks_out != NULL) back, returns when they are in host memory.
 * Pinned host memory is recommended (pageable works, slower). */
VA_API int va_run_fused_host(va_ctx* ctx, const float* h_protos, const float* h_coefs, const float* h_boxes,
                      const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out);

/* va_run_fused_host with fp16 prototypes in host memory (a model run with half = True emits them; the reference then
 * computes on protos.float(), ops.py:724).  Half the host -> device bytes of the fp32 entry point - the path is
 * PCIe-bound - and the same arithmetic: the prototypes are widened exactly on the device, everything after that is
 * va_run_fused.  h_protos_f16 [B][K][mh][mw] IEEE binary16; K*mh*mw must be a multiple of 8. */
VA_API int va_run_fused_host_f16(va_ctx* ctx, const uint16_t* h_protos_f16, const float* h_coefs, const float* h_boxes,
                          const int32_t* h_counts, int32_t B, uint8_t* h_masks_out, uint8_t* h_records_out);

/* Binary masks -> records.  Replaces FrameProcessor._extract_grid_information (+ penalties, peaks)
 * when the masks come from elsewhere (e.g. cv2.fillPoly of a polygon model output).
 *   masks [B][max_n][H][W] u8 (non-zero = inside), counts [B]
 *   rects [B][4] i32 = cv2.boundingRect (x, y, w, h) override of the selected instance's pixel
 *   bbox, or NULL; sel [B] i32 = instance to use, or NULL (largest area). */
VA_API int va_mask_to_records(va_ctx* ctx, const uint8_t* masks, const int32_t* counts, int32_t B,
                       const int32_t* rects, const int32_t* sel, uint8_t* records_out, void* stream);

/* Occupancy grids -> penalties + peaks.  Replaces PenaltyCalculator._pre_compute_easy_segments +
 * calculate_penalty for every cell (PenaltyCalculator.py:26-142) and ProtrusionDetector.__call__.
 *   hdr [B]; row_y/row_attr [B][rmax] i32; occ [B][rmax][cmax] u8 (bit0 non-empty, bit1 artificial);
 *   plane_y [B][rmax] i32 and plane_occ [B][rmax][cmax] u8 (bit0 non-empty) when n_plane > 0, else NULL.
 *   records_out [B][record_bytes] (same record as va_run_fused). */
VA_API int va_grid_to_penalty_peaks(va_ctx* ctx, const va_grid_input* hdr, const int32_t* row_y,
                             const int32_t* row_attr, const uint8_t* occ, const int32_t* plane_y,
                             const uint8_t* plane_occ, int32_t B, uint8_t* records_out, void* stream);

/* Front of the path: confidence filter + NMS on the raw segmentation-head output.
 * Replaces ops.non_max_suppression(pred, conf_thres, iou_thres, nc=nc, max_det=max_det) (vendored ultralytics,
 * testing/old/segmenting_using_tflite/ops.py:214-363, with torchvision.ops.nms) as model.predict(frame, conf=0.5)
 * runs it (FrameProcessor.py:322): best class only, class-offset boxes unless agnostic, stable descending scores.
 *   pred [B][4 + nc + K][A] f32: rows cx, cy, w, h, nc class confidences, K mask coefficients; A anchors
 *   outputs in the layout va_run_fused takes: coefs_out [B][max_n][K], boxes_out [B][max_n][4] (xyxy, input pixels),
 *   counts_out [B]; conf_out [B][max_n] f32 and cls_out [B][max_n] i32 may be NULL.  Slots >= counts are zero.
 *   max_det <= max_n <= 32 survivors are kept: the mask path carries at most 32 instances per frame (the reference's
 *   max_det default is 300; a frame with more than max_n survivors keeps the max_n best, like max_det = max_n).
 *   Any number of candidates: they are visited in tiles of 512 in order of (score descending, anchor ascending), each
 *   tile thinned by the survivors of the earlier ones, until max_det survive or the best 30000 (ops.py max_nms) are
 *   used up - the reference's result, not an approximation.  counts_out[b] is never negative. */
typedef struct va_nms_params {
  float conf_thres;   /* 0.5 in FrameProcessor.py:322 */
  float iou_thres;    /* ultralytics predict default 0.7 */
  int32_t nc;         /* number of classes */
  int32_t max_det;    /* <= max_n */
  int32_t agnostic;   /* 1 = no class offset */
  int32_t max_wh;     /* class offset in pixels, ops.py default 7680 */
} va_nms_params;
VA_API int va_nms(va_ctx* ctx, const float* pred, int32_t A, const va_nms_params* prm, int32_t B, float* coefs_out,
                  float* boxes_out, float* conf_out, int32_t* cls_out, int32_t* counts_out, void* stream);

/* ops.scale_boxes(img1_shape, boxes, img0_shape) + clip_boxes (ops.py:139-174, :367-385; padding = True, xyxy): the
 * kept boxes of va_nms, moved from the letterboxed model input (img1) back to the original frame (img0) as
 * SegmentationPredictor.postprocess does for Results.boxes AFTER process_mask has used the unscaled ones.
 *   boxes [B][max_n][4] f32, counts [B] -> boxes_out [B][max_n][4] (may alias boxes); slots >= counts are zero. */
VA_API int va_scale_boxes(va_ctx* ctx, const float* boxes, const int32_t* counts, int32_t B, int32_t img1_h, int32_t img1_w,
                          int32_t img0_h, int32_t img0_w, float* boxes_out, void* stream);

/* Multi-GPU record sink (frames are sharded across the GPUs of one box, only the per-frame records are
 * gathered).  Instead of a collective, every rank's tail kernel stores its records straight into the gathering
 * rank's memory over NVLink: records_out of va_run_fused may be a PEER pointer obtained here.
 *   va_peer_alloc   on the gathering rank: device buffer + a 64-byte handle to send to the other processes
 *   va_peer_open    on the other ranks: map that buffer (peer access is enabled by the mapping); *dptr is valid in
 *                   kernels of the context's device
 *   va_peer_put     enqueue a device-to-device copy of finished records into the (peer) buffer: runs on a copy engine,
 *                   no SM is involved (the other way to fill the buffer is to pass a peer pointer as records_out -
 *                   fewer steps, but the tail kernel's small stores then cross NVLink one by one)
 *   va_signal       enqueue "flag = value" (system-scope release) after everything queued before it on the stream -
 *                   the per-step "records of step k have landed" mark; flag may be a peer pointer
 *   va_wait_flags   enqueue a wait until flags[i] >= value for all i < n (system-scope acquire) on the stream */
#define VA_IPC_HANDLE_BYTES 64
VA_API int va_peer_alloc(va_ctx* ctx, uint64_t bytes, void** dptr, uint8_t handle[VA_IPC_HANDLE_BYTES]);
VA_API int va_peer_open(va_ctx* ctx, const uint8_t handle[VA_IPC_HANDLE_BYTES], void** dptr);
VA_API int va_peer_close(va_ctx* ctx, void* dptr);
VA_API int va_peer_free(va_ctx* ctx, void* dptr);
VA_API int va_peer_put(va_ctx* ctx, void* dst, const void* src, uint64_t bytes, void* stream);
VA_API int va_signal(va_ctx* ctx, int32_t* flag, int32_t value, void* stream);
VA_API int va_wait_flags(va_ctx* ctx, const int32_t* flags, int32_t n, int32_t value, void* stream);

/* Introspection for benchmarks: number of kernels launched by the last call, and which
 * contraction path the context uses (1 = tcgen05/TMEM, 0 = CUDA-core FFMA). */
VA_API int va_last_launch_count(const va_ctx* ctx);

/* Optional device timing of va_run_fused: with on = N > 0 every N-th call records CUDA events on the
 * caller's stream around the mask-assembly kernel(s) and the tail kernel (at most 512 timed calls
 * between reads; the events cost a few microseconds per timed call).  va_profile_read synchronises on
 * them, returns the summed milliseconds and the number of timed calls, and resets. */
VA_API int va_profile_enable(va_ctx* ctx, int32_t on);
VA_API int va_profile_read(va_ctx* ctx, float* assemble_ms, float* tail_ms, int32_t* calls);
VA_API int va_uses_tensor_core(const va_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VISION_ASSIST_B200_H */
