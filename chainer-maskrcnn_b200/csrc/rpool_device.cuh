// rpool_device.cuh -- device-side building blocks of the fused multi-level
// RoIAlign kernels (sm_100a).  Geometry follows, operation for operation, the
// reference's two coordinate recipes so that sample coordinates are bit-equal
// to the reference's (a 1-ulp difference in a coordinate near 300 is a 3e-5
// change of the interpolation weight, above the 1e-5 forward tolerance):
//   RPOOL_COORD_CHAINER: chainer_maskrcnn/functions/roi_align/roi_align_2d.py
//       forward :56-78, backward :154-178 (as evaluated by NumPy 2.x)
//   RPOOL_COORD_CAFFE2:  .../caffe2_operation/caffe2_roi_align.cpp:36-94,147-174
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/rpool_b200.h"

namespace rpool {

constexpr int kNT = 4;        // footprint width kept per bin and axis on the table path
constexpr int kPMax = 32;     // largest pooled extent served by the table path (forward)
constexpr int kPBwd = 16;     // ... by the table path of the backward kernel
constexpr int kSW = 8;        // window columns a chunk of bins may span (forward registers)
constexpr int kExt = 64;      // largest window extent (cells per axis) of the backward table path
#ifndef RPOOL_MIN_BLOCKS
#define RPOOL_MIN_BLOCKS 2   // resident CTAs per SM the pooling kernels are register-capped for
#endif
constexpr int kMaxThreads = 256;  // CTA size the pooling kernels are compiled for
constexpr int kMaxHeads = RPOOL_MAX_HEADS;
constexpr int kMaxLevels = RPOOL_MAX_LEVELS;

enum Path : int { kPathAuto = 0, kPathGeneric = 1, kPathTable = 2 };

struct LevelDev {
    float *data;
    int n_images, H, W;
    float scale;
};

struct KParams {
    LevelDev lvl[kMaxLevels];
    int n_levels;
    int C;
    int feat_layout, pool_layout;
    const float *rois;
    int R;
    int roi_format;
    const int *roi_level;  // resolved per-RoI level (plan output), never null at launch
    const int *order;      // launch schedule (plan output)
    int n_heads;
    int PH[kMaxHeads], PW[kMaxHeads];
    float *pooled[kMaxHeads];
    int S;
    int mode;
    int strip_cols;  // backward: entries of a warp's strip (512 bytes each)
    int reverse;     // walk the schedule backwards (set for the backward launch)
    // deterministic backward: per-RoI window rectangles {x0,y0,x1,y1}, float offsets of
    // their private windows in `det_scratch`, error flag (all in the plan workspace)
    int det;
    const int *det_rects;
    const unsigned long long *det_woff;
    float *det_scratch;
    unsigned long long det_scratch_floats;
    int *det_err;
    int prefetch;    // backward, L2 prefetch of gy: -1 off; -1-k row-ahead with a lead of k window rows
                     // (default k = 2); >= 0 whole RoI of the CTA scheduled that many slots later
    const unsigned char *recs;  // per-slot RoI records (rpool_tables_kernel), rec_stride bytes apart
    int rec_stride;
    int rec_head;    // first head part of the stored record this launch uses (0 unless heads are split)
    int force_path;
    int wait_fill;   // backward: this launch started in the zero fill's tail and must wait before it adds
    int staged_slots; // staged backward: schedule slots per CTA (ceil(R / gridDim.x))
    // tail of a launch in finer pieces: the RoIs at launch places >= tail_start are each served by
    // tail_parts CTAs (a contiguous share of the tasks each), so that the launch drains in a
    // fraction of a full CTA's duration.  tail_start = R, tail_parts = 1: one CTA per RoI throughout
    int tail_start, tail_parts;
};

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
struct AxisGeom {
    float start;      // RoI origin on the feature map
    float stride_f;   // chainer: float32 stride; caffe2: bin size
    double stride_d;  // chainer: double stride where the reference uses Python floats
    int use_dbl;
    int grid;         // samples per bin along this axis
    int size;         // H or W
    float inv_grid;
};

__device__ __forceinline__ AxisGeom make_axis(int mode, bool bwd, float lo, float hi,
                                              float scale, int P, int S, int size)
{
    AxisGeom g;
    g.start = __fmul_rn(lo, scale);
    const float end = __fmul_rn(hi, scale);
    const float len = __fsub_rn(end, g.start);
    g.size = size;
    g.stride_d = 0.0;
    g.use_dbl = 0;
    if (mode == RPOOL_COORD_CHAINER) {
        // max(len, 1.) is Python's max: keeps len unless 1. > len.
        const bool clamped = (1.0f > len);
        if (!bwd) {
            // stride = 1. * size / out  (float32) -- or a double when the clamp
            // returned the Python float 1.0                        (:62-65)
            g.use_dbl = clamped;
            g.stride_f = __fdiv_rn(len, (float)P);
            g.stride_d = 1.0 / (double)P;
        } else {
            // stride = float(size) / float(out): always a double    (:161-165)
            g.use_dbl = 1;
            g.stride_f = 0.f;
            g.stride_d = (clamped ? 1.0 : (double)len) / (double)P;
        }
        g.grid = 1;
    } else {
        const float l = (len < 1.0f) ? 1.0f : len;  // std::max(len, 1)  (:162-163)
        g.stride_f = __fdiv_rn(l, (float)P);         // bin size          (:164-165)
        g.grid = (S > 0) ? S : (int)ceilf(g.stride_f);  // roi_bin_grid    (:167-171)
    }
    g.inv_grid = 1.0f / (float)(g.grid > 0 ? g.grid : 1);
    return g;
}

// One sample along one axis: the two taps (i0 <= i1), their weights, and
// whether the sample counts at all (caffe2 drops samples beyond [-1, size]).
__device__ __forceinline__ bool axis_sample(const AxisGeom &g, int mode, int p, int s,
                                            int &i0, int &i1, float &w0, float &w1)
{
    const int bound = g.size - 1;
    if (mode == RPOOL_COORD_CHAINER) {
        float c;
        if (g.use_dbl)
            c = __fadd_rn((float)(((double)p + 0.5) * g.stride_d), g.start);
        else
            c = __fadd_rn(__fmul_rn((float)p + 0.5f, g.stride_f), g.start);
        const float fl = floorf(c);
        const float fr = __fsub_rn(c, fl);
        int lo = (int)fl;
        lo = lo < 0 ? 0 : lo;          // numpy.maximum(floor, 0)
        lo = lo > bound ? bound : lo;  // safety only: the reference raises IndexError here
        int hi = lo + 1;
        hi = hi > bound ? bound : hi;  // numpy.minimum(x0 + 1, size - 1)
        i0 = lo;
        i1 = hi;
        w1 = fr;
        w0 = __fsub_rn(1.0f, fr);
        return true;
    } else {
        // start + p*bin + (s + .5)*bin/grid, left to right, each op rounded
        float c = __fadd_rn(__fadd_rn(g.start, __fmul_rn((float)p, g.stride_f)),
                            __fdiv_rn(__fmul_rn((float)s + 0.5f, g.stride_f), (float)g.grid));
        const bool valid = !(c < -1.0f || c > (float)g.size);
        if (c <= 0.f) c = 0.f;
        int lo = (int)c, hi;
        if (lo >= bound) {
            hi = lo = bound;
            c = (float)lo;
        } else {
            hi = lo + 1;
        }
        const float l = __fsub_rn(c, (float)lo);
        i0 = lo;
        i1 = hi;
        w1 = l;
        w0 = __fsub_rn(1.0f, l);
        return valid;
    }
}

struct RoiBox {
    int b;
    float x1, y1, x2, y2;
};

__device__ __forceinline__ RoiBox load_roi(const float *rois, int r, int roi_format)
{
    const float *p = rois + (size_t)r * 5;
    RoiBox q;
    q.b = (int)__ldg(p);
    const float a = __ldg(p + 1), b = __ldg(p + 2), c = __ldg(p + 3), d = __ldg(p + 4);
    if (roi_format == RPOOL_ROI_YX) {
        q.y1 = a; q.x1 = b; q.y2 = c; q.x2 = d;
    } else {
        q.x1 = a; q.y1 = b; q.x2 = c; q.y2 = d;
    }
    return q;
}

// ---------------------------------------------------------------------------
// per-RoI footprint tables (fast path)
// ---------------------------------------------------------------------------
// For bin p of one axis the S samples' 2S taps are merged into a dense run of
// at most kNT cells starting at lo[p]; w[p] holds the summed weights (already
// divided by the grid size).  lo[] is non-decreasing in p.
struct AxisTab {
    int lo[kPMax];
    int n[kPMax];
    float4 w[kPMax];
};

// Per-head part of a RoI's record: footprint tables and the chunking of its bins.
struct HeadCtl {
    AxisTab tab[2];                      // [0 = y, 1 = x]
    int nmax[2];                         // widest footprint of any bin, per axis
    int nchunk;                          // bins cut into chunks of bounded x extent
    int pad0_;
    unsigned char cstart[kPMax + 4];     // first bin of every chunk (+ end marker)
    unsigned char cmask[kPMax];          // span offsets of a chunk that carry any weight
    unsigned char pad1_[4];
    int cx0[kPMax];                      // first column of each chunk's span
    unsigned long long ccnt[kPMax];      // bins per span offset, one byte each
};

// One RoI's record.  rpool_tables_kernel builds it once per plan (all RoIs in
// parallel) and stores it in the workspace; the pooling kernels copy their RoI's
// record into shared memory with one round of 128-bit loads instead of decoding
// the RoI and building the tables themselves (a serial, latency-bound prologue).
struct BlockCtl {
    int wmin[2], wmax[2];       // window extent over all heads: [0] rows, [1] cols
    int r, lvl, b, flags;       // RoI index, level, image; kRec* bits
    HeadCtl hd[kMaxHeads];
};
constexpr int kRecValid = 1;    // image index inside the level's tensor
constexpr int kRecShape = 2;    // layouts, channel count, pooled sizes and map width admit the table path
constexpr int kRecFits = 4;     // every footprint fits kNT cells
constexpr int kRecHeader = 32;
static_assert(sizeof(HeadCtl) % 16 == 0 && offsetof(BlockCtl, hd) == kRecHeader, "record layout");
__host__ __device__ constexpr int rec_bytes(int n_heads) { return kRecHeader + n_heads * (int)sizeof(HeadCtl); }

// Footprint of bin p along one axis: first cell, cell count (<= kNT), summed
// weights.  Returns false when the footprint does not fit kNT cells.
__device__ __forceinline__ bool axis_footprint(const AxisGeom &g, int mode, int p, int &lo_out,
                                               int &n_out, float (&w)[kNT])
{
    int lo = 0, n = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < kNT; ++k) w[k] = 0.f;
    for (int s = 0; s < g.grid; ++s) {
        int i0, i1;
        float w0, w1;
        const bool valid = axis_sample(g, mode, p, s, i0, i1, w0, w1);
        if (s == 0) lo = i0;
        if (!valid) continue;
        const int o0 = i0 - lo, o1 = i1 - lo;
        if (o1 >= kNT || o0 < 0) {
            ok = false;
            break;
        }
#pragma unroll
        for (int k = 0; k < kNT; ++k) {
            if (k == o0) w[k] += w0 * g.inv_grid;
            if (k == o1) w[k] += w1 * g.inv_grid;
        }
        n = o1 + 1 > n ? o1 + 1 : n;
    }
    if (g.grid > kNT) ok = false;
    lo_out = lo;
    n_out = n;
    return ok;
}

// Fills entry p of `t`; returns false when the footprint does not fit kNT cells.
__device__ __forceinline__ bool fill_axis_entry(AxisTab &t, const AxisGeom &g, int mode, int p,
                                                int &lo_out, int &hi_out)
{
    float w[kNT];
    int lo, n;
    const bool ok = axis_footprint(g, mode, p, lo, n, w);
    t.lo[p] = lo;
    t.n[p] = n;
    t.w[p] = make_float4(w[0], w[1], w[2], w[3]);
    lo_out = lo;
    hi_out = lo + n - 1;
    return ok;
}

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// streaming (evict-first) global accesses for data touched exactly once
#ifndef RPOOL_ST_POLICY
#define RPOOL_ST_POLICY ".cs"
#endif
__device__ __forceinline__ void stg_stream128(float *p, float4 v)
{
#ifdef RPOOL_DIAG_NOSTORE
    if (v.x != 1.2345e-30f) return;   // diagnostics only: keeps the value live, never stores
#endif
    asm volatile("st.global" RPOOL_ST_POLICY ".v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg_nc128(const float *p)
{
    float4 v;
#ifdef RPOOL_DIAG_NOLOAD
    const float a = __int_as_float((int)(reinterpret_cast<uintptr_t>(p) >> 4) & 0x3fffffff);  // diagnostics only
    return make_float4(a, a, a, a);
#endif
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// vector float reduction to global memory (sm_90+): one L2 atomic per 16 bytes
__device__ __forceinline__ void red_add_v4(float *p, float4 v)
{
#ifdef RPOOL_DIAG_NORED
    if (v.x != 1.2345e-30f) return;   // diagnostics only
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg_cg128(const float *p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg128(float *p, float4 v)
{
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Programmatic dependent launch: tells the scheduler that the NEXT kernel in the stream, when
// it was launched with the programmatic-serialisation attribute, may start as soon as every CTA
// of this grid has got here -- i.e. in this grid's tail, on the SMs that have run out of CTAs.
// (No effect otherwise.)  Used where the next kernel does not read what this one writes.
__device__ __forceinline__ void allow_dependents_early()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// The other half: a kernel that WAS launched early waits here until the kernels before it in the
// stream have completed and their writes are visible (returns at once in a normal launch).
__device__ __forceinline__ void wait_for_predecessors()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// bulk L2 prefetch through the TMA unit (bytes: multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}


// ---------------------------------------------------------------------------
// bulk copies (TMA unit) into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// global -> this CTA's shared memory; bytes: multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

}  // namespace rpool
