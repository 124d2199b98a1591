/*
 * oracle/roialign_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's RoIAlign arithmetic, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under chainer-maskrcnn_b200/ may link, load
 * or call this file.
 *
 * Two families, both on NCHW float32 features and (R,5) RoIs laid out as
 * [batch_index, x1, y1, x2, y2] -- exactly what the reference op receives:
 *
 *  "chainer" -- the reference's NumPy path, one bilinear sample at the bin
 *      centre.  Restates
 *        chainer_maskrcnn/functions/roi_align/roi_align_2d.py:39-88  (forward_cpu)
 *        chainer_maskrcnn/functions/roi_align/roi_align_2d.py:148-190 (backward_cpu)
 *      as they evaluate under NumPy 2.x (NEP 50): every coordinate op is a
 *      separately rounded float32 op, EXCEPT where the reference drops into
 *      Python floats (double): the forward stride when the max(.,1.) clamp
 *      fires (:62-65) and the backward stride always (:161-165).
 *
 *  "caffe2" -- the semantics of the reference's C++ forward port, general in
 *      sampling_ratio.  Restates
 *        .../caffe2_operation/caffe2_roi_align.cpp:19-113 (tap pre-calculation)
 *        .../caffe2_operation/caffe2_roi_align.cpp:115-226 (ROIAlignForward)
 *      The reference has no backward for it; orc_backward_caffe2 is the exact
 *      adjoint of that forward (each tap receives w*gy/count), which is also
 *      what caffe2's own RoIAlignGradient does.
 *
 * Pinning: tests/test_oracle.py checks these bit-for-bit against (a) the
 * unmodified reference Python module imported from /root/reference (when
 * present), (b) the reference C++ compiled into oracle/_ref, and (c) the
 * golden vectors under tests/golden/ that were generated from (a) and (b).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp; contraction
 * must stay off so that every mul/add rounds on its own, as NumPy does).
 *
 * Threading: the *_mt entry points split work so that every output element
 * is still accumulated by one thread in the reference's sequential order;
 * results are bit-identical to the single-threaded calls.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    int y0, y1, x0, x1; /* row/col indices of the four taps            */
    float p, q;         /* fractional parts along y and x              */
} orc_tap1;

/* One bin-centre tap set of the chainer path; `fwd` selects the forward
 * (:56-78) or backward (:154-178) coordinate arithmetic. */
static void chainer_roi_geometry(const float *roi, float scale, int outh, int outw,
                                 int fwd, float *ymin_o, float *xmin_o,
                                 float *sh_f, float *sw_f, double *sh_d, double *sw_d,
                                 int *h_is_double, int *w_is_double)
{
    /* idx, xmin, ymin, xmax, ymax = bottom_rois[i]; each times the scale (f32) */
    float xmin = roi[1] * scale;
    float ymin = roi[2] * scale;
    float xmax = roi[3] * scale;
    float ymax = roi[4] * scale;
    float rw = xmax - xmin;
    float rh = ymax - ymin;
    /* Python's max(a, 1.) keeps `a` unless 1. > a (NaN therefore survives). */
    int w_clamped = (1.0f > rw);
    int h_clamped = (1.0f > rh);
    *ymin_o = ymin;
    *xmin_o = xmin;
    if (fwd) {
        /* stride = 1. * size / out: float32 division unless the clamp produced
         * the Python float 1.0, in which case the stride is a double. */
        *h_is_double = h_clamped;
        *w_is_double = w_clamped;
        *sh_f = rh / (float)outh;
        *sw_f = rw / (float)outw;
        *sh_d = 1.0 / (double)outh;
        *sw_d = 1.0 / (double)outw;
    } else {
        /* stride = float(size) / float(out): always a double. */
        *h_is_double = 1;
        *w_is_double = 1;
        *sh_d = (h_clamped ? 1.0 : (double)rh) / (double)outh;
        *sw_d = (w_clamped ? 1.0 : (double)rw) / (double)outw;
        *sh_f = 0.f;
        *sw_f = 0.f;
    }
}

static inline float chainer_centre(int i, float s_f, double s_d, int is_double, float origin)
{
    /* (i + 0.5) * stride + origin.  With a float32 stride both ops are
     * float32; with a double stride the product is formed in double and is
     * rounded to float32 when it meets the float32 origin. */
    if (is_double) {
        float prod = (float)(((double)i + 0.5) * s_d);
        return prod + origin;
    } else {
        float prod = ((float)i + 0.5f) * s_f;
        return prod + origin;
    }
}

static inline void chainer_axis(float c, int bound, int *i0, int *i1, float *frac)
{
    float fl = floorf(c);
    *frac = c - fl;
    int lo = (int)fl;
    if (lo < 0) lo = 0;          /* numpy.maximum(floor, 0)               */
    int hi = lo + 1;
    if (hi > bound) hi = bound;  /* numpy.minimum(x0 + 1, size - 1)       */
    *i0 = lo;
    *i1 = hi;
}

static void chainer_taps(const float *roi, float scale, int H, int W, int outh, int outw,
                         int fwd, orc_tap1 *taps)
{
    float ymin, xmin, sh_f, sw_f;
    double sh_d, sw_d;
    int hd, wd;
    chainer_roi_geometry(roi, scale, outh, outw, fwd, &ymin, &xmin, &sh_f, &sw_f,
                         &sh_d, &sw_d, &hd, &wd);
    for (int ph = 0; ph < outh; ++ph) {
        float cy = chainer_centre(ph, sh_f, sh_d, hd, ymin);
        int y0, y1;
        float p;
        chainer_axis(cy, H - 1, &y0, &y1, &p);
        for (int pw = 0; pw < outw; ++pw) {
            float cx = chainer_centre(pw, sw_f, sw_d, wd, xmin);
            orc_tap1 *t = &taps[ph * outw + pw];
            t->y0 = y0;
            t->y1 = y1;
            t->p = p;
            chainer_axis(cx, W - 1, &t->x0, &t->x1, &t->q);
        }
    }
}

/* Forward of the reference NumPy path.  Returns 0, or 1+r if RoI r indexes
 * outside the map (the reference raises IndexError there). */
static int forward_chainer_range(const float *x, int N, int C, int H, int W,
                                 const float *rois, int r_begin, int r_end,
                                 int outh, int outw, float scale, float *top)
{
    orc_tap1 *taps = (orc_tap1 *)malloc(sizeof(orc_tap1) * (size_t)outh * outw);
    const size_t plane = (size_t)H * W;
    const int bins = outh * outw;
    for (int r = r_begin; r < r_end; ++r) {
        const float *roi = rois + (size_t)r * 5;
        int b = (int)roi[0];
        chainer_taps(roi, scale, H, W, outh, outw, 1, taps);
        for (int k = 0; k < bins; ++k) {
            if (taps[k].y0 >= H || taps[k].x0 >= W || b < 0 || b >= N) {
                free(taps);
                return 1 + r;
            }
        }
        for (int c = 0; c < C; ++c) {
            const float *src = x + ((size_t)b * C + c) * plane;
            float *dst = top + ((size_t)r * C + c) * bins;
            for (int k = 0; k < bins; ++k) {
                const orc_tap1 t = taps[k];
                float omp_ = 1.0f - t.p, omq = 1.0f - t.q;
                /* A*(1-p)*(1-q) + B*p*(1-q) + C*(1-p)*q + D*p*q, left to right */
                float a = (src[t.y0 * W + t.x0] * omp_) * omq;
                float bq = (src[t.y1 * W + t.x0] * t.p) * omq;
                float cq = (src[t.y0 * W + t.x1] * omp_) * t.q;
                float d = (src[t.y1 * W + t.x1] * t.p) * t.q;
                dst[k] = ((a + bq) + cq) + d;
            }
        }
    }
    free(taps);
    return 0;
}

ORC_API int orc_forward_chainer(const float *x, int N, int C, int H, int W,
                                const float *rois, int R, int outh, int outw,
                                float scale, float *top)
{
    return forward_chainer_range(x, N, C, H, W, rois, 0, R, outh, outw, scale, top);
}

ORC_API int orc_forward_chainer_mt(const float *x, int N, int C, int H, int W,
                                   const float *rois, int R, int outh, int outw,
                                   float scale, float *top, int threads)
{
    int err = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int r = 0; r < R; ++r) {
        int e = forward_chainer_range(x, N, C, H, W, rois, r, r + 1, outh, outw, scale, top);
        if (e) {
#pragma omp critical
            err = e;
        }
    }
    return err;
}

/* Backward of the reference NumPy path for channels [c_begin, c_end). */
static int backward_chainer_channels(const float *gy, const float *rois, int R,
                                     int N, int C, int H, int W, int outh, int outw,
                                     float scale, float *bottom_delta,
                                     int c_begin, int c_end)
{
    orc_tap1 *taps = (orc_tap1 *)malloc(sizeof(orc_tap1) * (size_t)outh * outw);
    const size_t plane = (size_t)H * W;
    const int bins = outh * outw;
    for (int r = 0; r < R; ++r) {
        const float *roi = rois + (size_t)r * 5;
        int b = (int)roi[0];
        chainer_taps(roi, scale, H, W, outh, outw, 0, taps);
        for (int k = 0; k < bins; ++k) {
            if (taps[k].y0 >= H || taps[k].x0 >= W || b < 0 || b >= N) {
                free(taps);
                return 1 + r;
            }
        }
        for (int c = c_begin; c < c_end; ++c) {
            float *dst = bottom_delta + ((size_t)b * C + c) * plane;
            const float *g = gy + ((size_t)r * C + c) * bins;
            for (int k = 0; k < bins; ++k) {
                const orc_tap1 t = taps[k];
                float omp_ = 1.0f - t.p, omq = 1.0f - t.q;
                /* weight product first, then times gy; four += in this order */
                dst[t.y0 * W + t.x0] += (omp_ * omq) * g[k];
                dst[t.y1 * W + t.x0] += (t.p * omq) * g[k];
                dst[t.y0 * W + t.x1] += (omp_ * t.q) * g[k];
                dst[t.y1 * W + t.x1] += (t.p * t.q) * g[k];
            }
        }
    }
    free(taps);
    return 0;
}

/* bottom_delta must be zero-filled by the caller (the reference allocates
 * numpy.zeros, :152); this function only accumulates. */
ORC_API int orc_backward_chainer(const float *gy, const float *rois, int R,
                                 int N, int C, int H, int W, int outh, int outw,
                                 float scale, float *bottom_delta)
{
    return backward_chainer_channels(gy, rois, R, N, C, H, W, outh, outw, scale,
                                     bottom_delta, 0, C);
}

ORC_API int orc_backward_chainer_mt(const float *gy, const float *rois, int R,
                                    int N, int C, int H, int W, int outh, int outw,
                                    float scale, float *bottom_delta, int threads)
{
    int err = 0;
    int chunk = 8;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int c0 = 0; c0 < C; c0 += chunk) {
        int c1 = c0 + chunk < C ? c0 + chunk : C;
        int e = backward_chainer_channels(gy, rois, R, N, C, H, W, outh, outw, scale,
                                          bottom_delta, c0, c1);
        if (e) {
#pragma omp critical
            err = e;
        }
    }
    return err;
}

/* ------------------------------------------------------------------ */
/* caffe2 semantics                                                    */
/* ------------------------------------------------------------------ */

typedef struct {
    int pos[4];
    float w[4];
} orc_tap4;

static int caffe2_grid(int sampling_ratio, float roi_size, int pooled)
{
    if (sampling_ratio > 0) return sampling_ratio;
    return (int)ceilf(roi_size / (float)pooled);
}

/* Tap table of one RoI: bins x grid_h x grid_w entries in (ph, pw, iy, ix)
 * order.  Returns the number of entries. */
static size_t caffe2_taps(const float *roi, float scale, int H, int W, int outh, int outw,
                          int sampling_ratio, orc_tap4 **buf, size_t *cap,
                          int *grid_h_o, int *grid_w_o)
{
    float start_w = roi[1] * scale;
    float start_h = roi[2] * scale;
    float end_w = roi[3] * scale;
    float end_h = roi[4] * scale;
    float roi_w = fmaxf(end_w - start_w, 1.0f);
    float roi_h = fmaxf(end_h - start_h, 1.0f);
    float bin_h = roi_h / (float)outh;
    float bin_w = roi_w / (float)outw;
    int gh = caffe2_grid(sampling_ratio, roi_h, outh);
    int gw = caffe2_grid(sampling_ratio, roi_w, outw);
    size_t n = (size_t)outh * outw * gh * gw;
    if (n > *cap) {
        *buf = (orc_tap4 *)realloc(*buf, n * sizeof(orc_tap4));
        *cap = n;
    }
    orc_tap4 *t = *buf;
    size_t k = 0;
    for (int ph = 0; ph < outh; ++ph)
        for (int pw = 0; pw < outw; ++pw)
            for (int iy = 0; iy < gh; ++iy) {
                /* start + ph*bin + (iy + .5)*bin/grid, each op rounded to f32 */
                float yy = (start_h + (float)ph * bin_h) +
                           (((float)iy + 0.5f) * bin_h) / (float)gh;
                for (int ix = 0; ix < gw; ++ix, ++k) {
                    float xx = (start_w + (float)pw * bin_w) +
                               (((float)ix + 0.5f) * bin_w) / (float)gw;
                    float y = yy, xq = xx;
                    if (y < -1.0f || y > (float)H || xq < -1.0f || xq > (float)W) {
                        memset(&t[k], 0, sizeof(orc_tap4));
                        continue;
                    }
                    if (y <= 0.f) y = 0.f;
                    if (xq <= 0.f) xq = 0.f;
                    int yl = (int)y, xl = (int)xq, yh, xh;
                    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                    if (xl >= W - 1) { xh = xl = W - 1; xq = (float)xl; } else xh = xl + 1;
                    float ly = y - (float)yl, lx = xq - (float)xl;
                    float hy = 1.0f - ly, hx = 1.0f - lx;
                    t[k].pos[0] = yl * W + xl; t[k].w[0] = hy * hx;
                    t[k].pos[1] = yl * W + xh; t[k].w[1] = hy * lx;
                    t[k].pos[2] = yh * W + xl; t[k].w[2] = ly * hx;
                    t[k].pos[3] = yh * W + xh; t[k].w[3] = ly * lx;
                }
            }
    *grid_h_o = gh;
    *grid_w_o = gw;
    return n;
}

static int forward_caffe2_range(const float *x, int N, int C, int H, int W,
                                const float *rois, int r_begin, int r_end,
                                int outh, int outw, float scale, int sampling_ratio,
                                float *top)
{
    orc_tap4 *taps = NULL;
    size_t cap = 0;
    const size_t plane = (size_t)H * W;
    const int bins = outh * outw;
    for (int r = r_begin; r < r_end; ++r) {
        const float *roi = rois + (size_t)r * 5;
        int b = (int)roi[0];
        if (b < 0 || b >= N) { free(taps); return 1 + r; }
        int gh, gw;
        caffe2_taps(roi, scale, H, W, outh, outw, sampling_ratio, &taps, &cap, &gh, &gw);
        const int per_bin = gh * gw;
        const float count = (float)per_bin;
        for (int c = 0; c < C; ++c) {
            const float *src = x + ((size_t)b * C + c) * plane;
            float *dst = top + ((size_t)r * C + c) * bins;
            const orc_tap4 *t = taps;
            for (int k = 0; k < bins; ++k) {
                float acc = 0.f;
                for (int s = 0; s < per_bin; ++s, ++t) {
                    /* acc += w1*v1 + w2*v2 + w3*v3 + w4*v4 */
                    float v = ((t->w[0] * src[t->pos[0]] + t->w[1] * src[t->pos[1]]) +
                               t->w[2] * src[t->pos[2]]) + t->w[3] * src[t->pos[3]];
                    acc += v;
                }
                dst[k] = acc / count;
            }
        }
    }
    free(taps);
    return 0;
}

ORC_API int orc_forward_caffe2(const float *x, int N, int C, int H, int W,
                               const float *rois, int R, int outh, int outw,
                               float scale, int sampling_ratio, float *top)
{
    return forward_caffe2_range(x, N, C, H, W, rois, 0, R, outh, outw, scale,
                                sampling_ratio, top);
}

ORC_API int orc_forward_caffe2_mt(const float *x, int N, int C, int H, int W,
                                  const float *rois, int R, int outh, int outw,
                                  float scale, int sampling_ratio, float *top, int threads)
{
    int err = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int r = 0; r < R; ++r) {
        int e = forward_caffe2_range(x, N, C, H, W, rois, r, r + 1, outh, outw, scale,
                                     sampling_ratio, top);
        if (e) {
#pragma omp critical
            err = e;
        }
    }
    return err;
}

static int backward_caffe2_channels(const float *gy, const float *rois, int R,
                                    int N, int C, int H, int W, int outh, int outw,
                                    float scale, int sampling_ratio, float *bottom_delta,
                                    int c_begin, int c_end)
{
    orc_tap4 *taps = NULL;
    size_t cap = 0;
    const size_t plane = (size_t)H * W;
    const int bins = outh * outw;
    for (int r = 0; r < R; ++r) {
        const float *roi = rois + (size_t)r * 5;
        int b = (int)roi[0];
        if (b < 0 || b >= N) { free(taps); return 1 + r; }
        int gh, gw;
        caffe2_taps(roi, scale, H, W, outh, outw, sampling_ratio, &taps, &cap, &gh, &gw);
        const int per_bin = gh * gw;
        const float count = (float)per_bin;
        for (int c = c_begin; c < c_end; ++c) {
            float *dst = bottom_delta + ((size_t)b * C + c) * plane;
            const float *g = gy + ((size_t)r * C + c) * bins;
            const orc_tap4 *t = taps;
            for (int k = 0; k < bins; ++k)
                for (int s = 0; s < per_bin; ++s, ++t)
                    for (int q = 0; q < 4; ++q)
                        dst[t->pos[q]] += (g[k] * t->w[q]) / count;
        }
    }
    free(taps);
    return 0;
}

ORC_API int orc_backward_caffe2(const float *gy, const float *rois, int R,
                                int N, int C, int H, int W, int outh, int outw,
                                float scale, int sampling_ratio, float *bottom_delta)
{
    return backward_caffe2_channels(gy, rois, R, N, C, H, W, outh, outw, scale,
                                    sampling_ratio, bottom_delta, 0, C);
}

ORC_API int orc_backward_caffe2_mt(const float *gy, const float *rois, int R,
                                   int N, int C, int H, int W, int outh, int outw,
                                   float scale, int sampling_ratio, float *bottom_delta,
                                   int threads)
{
    int err = 0;
    int chunk = 8;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int c0 = 0; c0 < C; c0 += chunk) {
        int c1 = c0 + chunk < C ? c0 + chunk : C;
        int e = backward_caffe2_channels(gy, rois, R, N, C, H, W, outh, outw, scale,
                                         sampling_ratio, bottom_delta, c0, c1);
        if (e) {
#pragma omp critical
            err = e;
        }
    }
    return err;
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
