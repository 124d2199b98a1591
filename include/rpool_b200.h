/*
 * rpool_b200.h -- C ABI of the B200-native FPN multi-level RoIAlign library
 * (librpool_b200.so, built from chainer-maskrcnn_b200/csrc/ for sm_100a).
 *
 * This is the drop-in boundary for the reference's RoI feature-pooling path.
 * Citations are relative to the katotetsuro/chainer-maskrcnn tree:
 *
 *   rpool_forward / rpool_backward replace
 *     - the CuPy kernels roi_align_2d_fwd / roi_align_2d_bwd
 *       (chainer_maskrcnn/functions/roi_align/roi_align_2d.py:100-144, :196-279),
 *     - the pybind11 module caffe2_roi_align.forward
 *       (.../caffe2_operation/caffe2_roi_align.cpp:231-248), and
 *     - the heads' per-RoI dispatch loops that call them once per RoI
 *       (chainer_maskrcnn/model/head/fpn_roi_mask_head.py:57-63,74-78,90-95;
 *        fpn_roi_keypoint_head.py:59-71,83-87,99-104)
 *     with ONE launch per direction over the whole pyramid and all heads (the
 *     backward pass of a two-size problem runs one launch per pooled size).
 *   rpool_assign_levels replaces map_rois_to_fpn_levels
 *       (chainer_maskrcnn/model/rpn/multilevel_region_proposal_network.py:16-31)
 *       plus the clip at chainer_maskrcnn/model/maskrcnn.py:141.
 *   roi_format RPOOL_ROI_YX folds in _roi_align_2d_yx's column permutation
 *       (chainer_maskrcnn/functions/roi_align_2d_yx.py:5).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, a POD problem descriptor; no
 *     allocation inside the library (caller owns outputs and workspace);
 *   - every entry point that launches work takes a cudaStream_t passed as
 *     void* and is asynchronous: no hidden device synchronisation;
 *   - return value 0 = RPOOL_OK, otherwise an rpool_status; the message of the
 *     last failure on the calling thread is at rpool_last_error();
 *   - there is no CPU fallback: without a CUDA device every launch fails.
 */
#ifndef RPOOL_B200_H_
#define RPOOL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPOOL_VERSION 200          /* 0.2.0 */
#define RPOOL_MAX_LEVELS 8
#define RPOOL_MAX_HEADS 2

#if defined(__GNUC__)
#define RPOOL_API __attribute__((visibility("default")))
#else
#define RPOOL_API
#endif

typedef enum rpool_status {
    RPOOL_OK = 0,
    RPOOL_ERR_INVALID = 1,      /* bad argument (null pointer, size, enum)       */
    RPOOL_ERR_UNSUPPORTED = 2,  /* valid but not implemented combination         */
    RPOOL_ERR_WORKSPACE = 3,    /* workspace missing or too small                */
    RPOOL_ERR_CUDA = 4          /* a CUDA runtime call failed (see last_error)   */
} rpool_status;

typedef enum rpool_layout {
    RPOOL_NHWC = 0,  /* channels-last: features (N,H,W,C), pooled (R,PH,PW,C)    */
    RPOOL_NCHW = 1   /* the reference's layout: (N,C,H,W), (R,C,PH,PW)           */
} rpool_layout;

typedef enum rpool_roi_format {
    RPOOL_ROI_XY = 0,  /* [batch, x1, y1, x2, y2]  roi_align_2d.py:296-297        */
    RPOOL_ROI_YX = 1   /* [batch, y1, x1, y2, x2]  the heads' indices_and_rois    */
} rpool_roi_format;

typedef enum rpool_coord_mode {
    /* reference NumPy path: one sample at the bin centre, forward and backward
     * coordinates rounded exactly as roi_align_2d.py:56-78 / :154-178 do under
     * NumPy 2.x.  Requires sampling_ratio == 1. */
    RPOOL_COORD_CHAINER = 0,
    /* caffe2 semantics of the reference C++ port (caffe2_roi_align.cpp:19-113):
     * sampling_ratio x sampling_ratio samples per bin (<=0: adaptive grid),
     * samples beyond [-1, size] contribute zero; backward is its adjoint. */
    RPOOL_COORD_CAFFE2 = 1
} rpool_coord_mode;

/* Per-call options (all zero = defaults).  They live in the problem block: the
 * library keeps no process-wide mutable state besides a per-device cache of the
 * shared-memory attribute it has already raised. */
typedef enum rpool_schedule {
    RPOOL_SCHED_DEFAULT = 0,     /* RoIs binned by (image, level fine->coarse): L2 locality */
    RPOOL_SCHED_INPUT = 1,       /* input order */
    RPOOL_SCHED_LEVEL_DESC = 2,  /* (image, level coarse->fine) */
    RPOOL_SCHED_COARSE_FIRST = 3 /* (level coarse->fine, image) */
} rpool_schedule;

typedef enum rpool_path {
    RPOOL_PATH_AUTO = 0,
    RPOOL_PATH_GENERIC = 1, /* tap by tap in the reference's operation order (any layout) */
    RPOOL_PATH_TABLE = 2    /* vectorised channels-last path where the shapes allow it    */
} rpool_path;

typedef struct rpool_options {
    int32_t cta_threads;    /* CTA size of the pooling kernels: multiple of 32 in [32,256]; 0 = 128, or 256
                             * for a problem of at most 592 RoIs (less than one wave of 128-thread CTAs) */
    int32_t schedule;       /* rpool_schedule */
    int32_t force_path;     /* rpool_path */
    int32_t fuse_heads_backward; /* 0: one backward launch per pooled size; 1: one launch for both */
    int32_t prefetch_rows;  /* backward, rows variant: bulk L2 prefetch of the upstream gradient rows
                             * the warp's task `prefetch_rows` window rows ahead is the first to need;
                             * 0 = default (2), -1 = off, 1..16 */
    int32_t prefetch_rois;  /* backward, rows variant: > 0 replaces the row-ahead prefetch by the whole
                             * RoI of the CTA scheduled prefetch_rois - 1 slots later */
    int32_t zero_fill_in_tail; /* rpool_backward / rpool_zero_fill: 1 = the zero fill may start while the
                             * kernel queued before it on the stream is still draining (programmatic
                             * dependent launch).  The CALLER guarantees that that kernel neither touches
                             * the gradient buffers nor produces the upstream gradients -- e.g. it is this
                             * problem's rpool_forward, as in the step helper of the Python package.
                             * 0 = plain stream order */
    int32_t backward_variant; /* rpool_backward, atomic path: 0 or 1 = "rows" kernel (a CTA per RoI, gy through
                             * registers; the default), 2 = "staged" kernel (a persistent CTA per SM, gy copied
                             * into shared memory by the TMA unit; measured slower) where it
                             * applies -- channels-last, C a multiple of 128 up to 512, one pooled size per
                             * launch whose staged item fits, at most 128 RoIs per SM -- and rows elsewhere */
} rpool_options;

/* One pyramid level.  `data` is the feature map in rpool_forward (read) and
 * the dense feature gradient in rpool_backward (written). */
typedef struct rpool_level {
    void *data;
    int32_t n_images;
    int32_t height;
    int32_t width;
    float spatial_scale; /* 1/stride, feature_pyramid_network.py:9-11 */
} rpool_level;

typedef struct rpool_problem {
    /* pyramid */
    int32_t n_levels;   /* 1 = the single-level op roi_align_2d                  */
    int32_t channels;
    int32_t feat_layout; /* rpool_layout of every level                           */
    int32_t pool_layout; /* rpool_layout of pooled outputs / their gradients      */
    rpool_level level[RPOOL_MAX_LEVELS];

    /* RoIs: device pointer to (n_rois, 5) float32 */
    const float *rois;
    int32_t n_rois;
    int32_t roi_format; /* rpool_roi_format */

    /* Level of each RoI: device int32[n_rois] (clipped to the pyramid like
     * maskrcnn.py:141), or NULL to assign on the device from the area
     * thresholds below (level = k_min + #{t : thresholds[t] <= area}). */
    const int32_t *roi_levels;
    /* Same, as the float32 array map_rois_to_fpn_levels returns (the heads cast
     * it with astype(int32), fpn_roi_mask_head.py:58).  Used when roi_levels is
     * NULL and this is not. */
    const float *roi_levels_f32;
    float level_thresholds[RPOOL_MAX_LEVELS];
    int32_t n_thresholds;
    int32_t k_min;

    /* heads pooled in the same launch (box 7x7, mask 14x14 ...) */
    int32_t n_heads;
    int32_t out_h[RPOOL_MAX_HEADS];
    int32_t out_w[RPOOL_MAX_HEADS];
    /* forward: outputs (written).  backward: upstream gradients gy (read). */
    void *pooled[RPOOL_MAX_HEADS];

    int32_t sampling_ratio;
    int32_t coord_mode; /* rpool_coord_mode */

    /* backward only */
    int32_t accumulate;    /* 0: gradients are zero-filled first; 1: += into them */
    int32_t deterministic; /* 0: atomics; 1: run-to-run reproducible summation     */
    /* deterministic == 1 only: scratch for the per-RoI private windows, sized by
     * rpool_backward_det_bytes (device memory, 16-byte aligned) */
    void *det_workspace;
    size_t det_workspace_bytes;

    rpool_options opt;
} rpool_problem;

RPOOL_API int rpool_version(void);
RPOOL_API const char *rpool_last_error(void);

/* Number of kernels this library has launched in the calling process. */
RPOOL_API uint64_t rpool_launch_count(void);

/* Hash of the sources this library was compiled from ("unknown" when the build
 * did not define RPOOL_BUILD_ID): bindings that have the sources at hand compare
 * it with their own hash and refuse a stale binary. */
RPOOL_API const char *rpool_build_id(void);

/* Host helper: float32 area thresholds of floor(lvl0 + log2(sqrt(area)/s0 + eps))
 * for levels k_min+1..k_max, found by bisection with this libc's log2f.  The
 * Python shim derives them from NumPy instead (the reference's arithmetic) and
 * the tests assert that both agree.  Writes k_max-k_min floats. */
RPOOL_API int rpool_level_thresholds(float s0, float lvl0, float eps, int k_min, int k_max,
                           float *out_thresholds);

/* map_rois_to_fpn_levels on the device.  boxes: (n, box_stride) float32 whose
 * LAST four columns are the box in `roi_format` order (box_stride 4 or 5).
 * level = clip(k_min + #{t: thr[t] <= (y2-y1)*(x2-x1)}, k_min, k_cap).
 * Either output may be NULL. */
RPOOL_API int rpool_assign_levels(const float *boxes, int32_t n, int32_t box_stride, int32_t roi_format,
                        const float *thresholds_host, int32_t n_thresholds,
                        int32_t k_min, int32_t k_cap,
                        float *levels_f32, int32_t *levels_i32, void *stream);

/* Scratch needed by rpool_plan/forward/backward for up to n_rois RoIs (16-byte
 * aligned device memory): schedule arrays plus one record per RoI (footprint
 * tables, about 2 KB per head).  rpool_workspace_bytes is the bound for any
 * problem; the _ex form is exact for a head count and rpool_coord_mode. */
RPOOL_API size_t rpool_workspace_bytes(int32_t n_rois);
RPOOL_API size_t rpool_workspace_bytes_ex(int32_t n_rois, int32_t n_heads, int32_t coord_mode);

/* sizeof(rpool_problem) as compiled, so that FFI bindings can verify their
 * struct layout against the library's. */
RPOOL_API size_t rpool_problem_size(void);

/* Bins the RoIs by (image, level) into the launch schedule kept in `workspace`
 * (level of every RoI + a stable permutation) and builds every RoI's footprint
 * tables there (rpool_tables_kernel: all RoIs in parallel, once, instead of in
 * every pooling CTA).  Reads the geometry only: level/pooled addresses are not
 * dereferenced.  Must precede forward/backward on the same stream; a plan stays
 * valid while rois/roi_levels, the level shapes and scales, the pooled sizes,
 * sampling_ratio, coord_mode and the layouts are unchanged. */
RPOOL_API int rpool_plan(const rpool_problem *problem, void *workspace, size_t workspace_bytes,
               void *stream);

/* pooled[h][r] = RoIAlign(level[lvl(r)].data, rois[r]) for every head h; row r
 * of every output corresponds to input RoI r (fpn_roi_mask_head.py:59-63). */
RPOOL_API int rpool_forward(const rpool_problem *problem, void *workspace, size_t workspace_bytes,
                  void *stream);

/* level[l].data (+)= sum over heads and RoIs of the transposed interpolation
 * applied to pooled[h] (= gy).  No gradient w.r.t. RoIs (roi_align_2d.py:190). */
RPOOL_API int rpool_backward(const rpool_problem *problem, void *workspace, size_t workspace_bytes,
                   void *stream);

/* Zero fill of every level[l].data (N*H*W*C floats each) in one launch: what
 * rpool_backward does first when accumulate == 0.  Exposed so that a caller can
 * run it early on another stream (it depends on nothing) and then call
 * rpool_backward with accumulate = 1 -- the step helper of the Python package
 * overlaps it with rpool_plan / rpool_forward that way.  Needs no workspace. */
RPOOL_API int rpool_zero_fill(const rpool_problem *problem, void *stream);

/* Deterministic variant (problem->deterministic = 1): a segmented reduction.
 * Every RoI writes its window contribution to a private window in
 * det_workspace with plain stores, then every feature cell sums the windows that
 * cover it in schedule order and is written exactly once (no atomics, no zero
 * fill): bit-identical from run to run.  The scratch size depends on the RoIs:
 * rpool_backward_det_bytes computes it on the device and SYNCHRONISES `stream`
 * to return it.  RoIs that would need the generic kernel path
 * cannot be ordered: they are skipped and flagged (RPOOL_FLAG_DET_*). */
RPOOL_API int rpool_backward_det_bytes(const rpool_problem *problem, void *workspace,
                                       size_t workspace_bytes, void *stream, size_t *bytes_out);

/* Sticky flags the kernels raise in the workspace (cleared by rpool_plan):
 * what the reference reports with an exception (IndexError in the NumPy path,
 * roi_align_2d.py:76-86) is computed safely here and flagged instead.
 * rpool_status_flags copies them to the host and SYNCHRONISES `stream`. */
#define RPOOL_FLAG_BAD_BATCH 1     /* a RoI's batch index is outside [0, n_images): zero rows, no gradient */
#define RPOOL_FLAG_LEVEL_CLIPPED 2 /* informational: a GIVEN level was outside the pyramid and was clipped like maskrcnn.py:141 */
#define RPOOL_FLAG_DET_GENERIC 4   /* deterministic backward: a RoI needs the generic path and was skipped */
#define RPOOL_FLAG_DET_SCRATCH 8   /* deterministic backward: det_workspace does not match the plan */
RPOOL_API int rpool_status_flags(const void *workspace, int32_t n_rois, void *stream,
                                 int32_t *flags_out);

/* Read back the schedule of the last plan (for tests): device->host copies of
 * the per-RoI level and the permutation; synchronises `stream`. */
RPOOL_API int rpool_read_plan(const void *workspace, int32_t n_rois, int32_t *levels_host,
                    int32_t *order_host, void *stream);

/* Layout conversion for callers that hold the reference's NCHW arrays. */
RPOOL_API int rpool_nchw_to_nhwc(const float *src, float *dst, int32_t n, int32_t c, int32_t h,
                       int32_t w, void *stream);
RPOOL_API int rpool_nhwc_to_nchw(const float *src, float *dst, int32_t n, int32_t c, int32_t h,
                       int32_t w, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RPOOL_B200_H_ */
