// rpool_kernels.cuh -- the fused multi-level RoIAlign kernels (sm_100a).
//
// One CTA per RoI, scheduled through the plan's (image, level) binning.  A CTA
//   1. decodes its RoI and level and builds per-axis footprint tables in shared
//      memory: per bin the merged bilinear taps of its S samples (a dense run of
//      at most kNT cells and their summed weights) -- the interpolation is
//      separable, out = WY * X * WX^T;
//   2. picks a path:
//        table   -- channels-last tensors, C % 4 == 0: warps take independent
//                   (row, 128-channel slab) tasks, a lane owns 4 channels and
//                   moves them with 128-bit accesses; the y-interpolated row
//                   lives in a per-warp strip of shared memory, so both passes
//                   are branch-free gathers (see the two kernels below);
//        generic -- any layout / sampling grid / pooled size, tap by tap, in
//                   the reference's own operation order (bit-equal forward).
//   Forward streams the bins out with evict-first 128-bit stores; backward is
//   the adjoint written as a gather per window cell, so the only atomics are one
//   128-bit vector reduction (red.global.add.v4.f32) per window cell and RoI.
#pragma once
#include "rpool_device.cuh"

namespace rpool {

// ---------------------------------------------------------------------------
// shared prologue: RoI decode + tables
// ---------------------------------------------------------------------------
struct RoiCtx {
    int r, lvl, b;
    bool valid;       // batch index inside the level's tensor
    bool fast_ok;     // layouts, channel count, pooled sizes and map width allow the table paths
    LevelDev L;
    RoiBox box;
};

// Geometry of one axis of head h (computed where it is needed only: keeping it
// per thread for every head would live in local memory).
__device__ __forceinline__ AxisGeom axis_of(const KParams &P, const RoiCtx &c, bool bwd, int h, int axis)
{
    return axis ? make_axis(P.mode, bwd, c.box.x1, c.box.x2, c.L.scale, P.PW[h], P.S, c.L.W)
                : make_axis(P.mode, bwd, c.box.y1, c.box.y2, c.L.scale, P.PH[h], P.S, c.L.H);
}

// Layouts, channel count, pooled sizes and map width admit the table paths.
__device__ __forceinline__ bool shapes_allow_tables(const KParams &P, const LevelDev &L)
{
    bool ok = (P.feat_layout == RPOOL_NHWC) && (P.pool_layout == RPOOL_NHWC) && (P.C % 4 == 0);
    for (int h = 0; h < P.n_heads; ++h) ok = ok && P.PH[h] <= kPMax && P.PW[h] <= kPMax;
    // the table path reads spans of kSW columns: the map must be that wide
    return ok && L.W >= kSW;
}

__device__ __forceinline__ void roi_decode(const KParams &P, int slot, RoiCtx &c)
{
    c.r = P.order[slot];
    int lvl = P.roi_level[c.r];
    lvl = lvl < 0 ? 0 : (lvl >= P.n_levels ? P.n_levels - 1 : lvl);
    c.lvl = lvl;
    c.L = P.lvl[lvl];
    const RoiBox q = load_roi(P.rois, c.r, P.roi_format);
    c.b = q.b;
    c.box = q;
    c.valid = (q.b >= 0 && q.b < c.L.n_images);
    c.fast_ok = shapes_allow_tables(P, c.L);
}

// Pieces per RoI in the tail of a launch (see split_tail in rpool_api.cu).  Measured
// (profiles/r02_experiments.log, r03j..l): two pieces per RoI in the forward launch's tail take 2.4 % off
// the forward of configs[1] and 5.3 % off a 2 000-RoI shard of configs[3]; three or four are no better;
// the backward launch, which already ends on its shortest CTAs, does not gain -- and with 1 here the
// backward kernel carries none of the bookkeeping.
constexpr int kTailPartsFwd = 2, kTailPartsBwd = 1;

// Launch place of this CTA (the RoI's position in launch order) and, in the split tail of the
// launch, which share of the RoI's tasks it takes.
__device__ __forceinline__ int launch_place(const KParams &P, int &part)
{
    const int b = (int)blockIdx.x;
    part = 0;
    if (b < P.tail_start) return b;
    const int j = b - P.tail_start;
    part = j % P.tail_parts;
    return P.tail_start + j / P.tail_parts;
}
__device__ __forceinline__ int launch_parts(const KParams &P)
{
    return (int)blockIdx.x < P.tail_start ? 1 : P.tail_parts;
}
__device__ __forceinline__ int launch_slot(const KParams &P)
{
    // the backward launch walks the schedule from its far end: coarse levels (the
    // widest windows, the longest CTAs) first, short CTAs in the tail of the launch
    int part;
    const int place = launch_place(P, part);
    return P.reverse ? P.R - 1 - place : place;
}

__device__ __forceinline__ void roi_prologue(const KParams &P, RoiCtx &c)
{
    roi_decode(P, launch_slot(P), c);
}

// Copies this CTA's RoI record (header + n_heads head parts) from the workspace
// into shared memory; returns with the CTA synchronised.
__device__ __forceinline__ void load_record(const KParams &P, BlockCtl *ctl)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(P.recs + (size_t)launch_slot(P) * P.rec_stride);
    uint4 *dst = reinterpret_cast<uint4 *>(ctl);
    // header, then this launch's n_heads head parts starting at part rec_head of the stored
    // record (a launch over one pooled size of a two-size plan reads its own part as head 0)
    constexpr int kHdr16 = kRecHeader >> 4;
    const int skip16 = P.rec_head * (int)(sizeof(HeadCtl) >> 4);
    const int n16 = rec_bytes(P.n_heads) >> 4;
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
        dst[i] = __ldg(src + (i < kHdr16 ? i : i + skip16));
    // the record of the CTA scheduled kRecAhead places later: into L2 meanwhile (part of the plan's output
    // has been pushed out to HBM by the maps streaming through L2 by the time its CTA starts).  Measured
    // on cfg 1: forward 0.1805 -> 0.1777 ms with 74, 0.1785 with 148-296, 0.1789 with 592; backward unchanged
    constexpr int kRecAhead = 74;
    if (threadIdx.x == 32 && (int)blockIdx.x + kRecAhead < P.tail_start) {
        const int ahead = (int)blockIdx.x + kRecAhead;
        const int slot = P.reverse ? P.R - 1 - ahead : ahead;
        prefetch_l2_bulk(P.recs + (size_t)slot * P.rec_stride, (unsigned)P.rec_stride);
    }
    __syncthreads();
}

// What the table path needs beyond the record: the tensors of this launch must
// allow 128-bit accesses (the record was built from the geometry alone).
__device__ __forceinline__ bool pointers_aligned(const KParams &P, const LevelDev &L)
{
    bool ok = (reinterpret_cast<uintptr_t>(L.data) & 15) == 0;
    for (int h = 0; h < P.n_heads; ++h) ok = ok && ((reinterpret_cast<uintptr_t>(P.pooled[h]) & 15) == 0);
    return ok;
}

// ---------------------------------------------------------------------------
// generic path (reference operation order; any layout)
// ---------------------------------------------------------------------------
struct Strides4 {
    long long s0, s1, s2, s3;  // element strides of (n|r, c, h, w)
};
__device__ __forceinline__ Strides4 strides_of(int layout, int C, int H, int W)
{
    Strides4 s;
    if (layout == RPOOL_NHWC) {
        s.s0 = (long long)H * W * C; s.s1 = 1; s.s2 = (long long)W * C; s.s3 = C;
    } else {
        s.s0 = (long long)C * H * W; s.s1 = (long long)H * W; s.s2 = W; s.s3 = 1;
    }
    return s;
}

// (noinline, and decoding the RoI again itself, so that neither its registers nor
// a spilled context reach the table path)
__device__ __noinline__ void generic_forward(const KParams &P)
{
    RoiCtx c;
    roi_prologue(P, c);
    const int C = P.C;
    const Strides4 fs = strides_of(P.feat_layout, C, c.L.H, c.L.W);
    const float *feat = c.L.data + (long long)(c.valid ? c.b : 0) * fs.s0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int PH = P.PH[h], PW = P.PW[h];
        const Strides4 os = strides_of(P.pool_layout, C, PH, PW);
        float *out = P.pooled[h] + (long long)c.r * os.s0;
        const AxisGeom gy = axis_of(P, c, false, h, 0), gx = axis_of(P, c, false, h, 1);
        const int total = PH * PW * C;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            int ch, bin;
            if (P.pool_layout == RPOOL_NHWC) { ch = idx % C; bin = idx / C; }
            else { bin = idx % (PH * PW); ch = idx / (PH * PW); }
            const int ph = bin / PW, pw = bin % PW;
            float res = 0.f;
            if (c.valid) {
                const float *f = feat + ch * fs.s1;
                if (P.mode == RPOOL_COORD_CHAINER) {
                    int y0, y1, x0, x1;
                    float a0, a1, b0, b1;
                    axis_sample(gy, P.mode, ph, 0, y0, y1, a0, a1);  // a0 = 1-p, a1 = p
                    axis_sample(gx, P.mode, pw, 0, x0, x1, b0, b1);  // b0 = 1-q, b1 = q
                    // A*(1-p)*(1-q) + B*p*(1-q) + C*(1-p)*q + D*p*q, left to right (:82-86)
                    const float A = __ldg(f + y0 * fs.s2 + x0 * fs.s3);
                    const float B = __ldg(f + y1 * fs.s2 + x0 * fs.s3);
                    const float Cc = __ldg(f + y0 * fs.s2 + x1 * fs.s3);
                    const float D = __ldg(f + y1 * fs.s2 + x1 * fs.s3);
                    float v = __fmul_rn(__fmul_rn(A, a0), b0);
                    v = __fadd_rn(v, __fmul_rn(__fmul_rn(B, a1), b0));
                    v = __fadd_rn(v, __fmul_rn(__fmul_rn(Cc, a0), b1));
                    v = __fadd_rn(v, __fmul_rn(__fmul_rn(D, a1), b1));
                    res = v;
                } else {
                    float acc = 0.f;
                    for (int iy = 0; iy < gy.grid; ++iy) {
                        int y0, y1;
                        float hy, ly;
                        const bool vy = axis_sample(gy, P.mode, ph, iy, y0, y1, hy, ly);
                        for (int ix = 0; ix < gx.grid; ++ix) {
                            int x0, x1;
                            float hx, lx;
                            const bool vx = axis_sample(gx, P.mode, pw, ix, x0, x1, hx, lx);
                            if (!(vy && vx)) continue;
                            // w1*v1 + w2*v2 + w3*v3 + w4*v4 (:205-208)
                            const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx);
                            const float w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
                            float v = __fmul_rn(w1, __ldg(f + y0 * fs.s2 + x0 * fs.s3));
                            v = __fadd_rn(v, __fmul_rn(w2, __ldg(f + y0 * fs.s2 + x1 * fs.s3)));
                            v = __fadd_rn(v, __fmul_rn(w3, __ldg(f + y1 * fs.s2 + x0 * fs.s3)));
                            v = __fadd_rn(v, __fmul_rn(w4, __ldg(f + y1 * fs.s2 + x1 * fs.s3)));
                            acc = __fadd_rn(acc, v);
                        }
                    }
                    res = __fdiv_rn(acc, (float)(gy.grid * gx.grid));  // output_val /= count
                }
            }
            out[ch * os.s1 + ph * os.s2 + pw * os.s3] = res;
        }
    }
}

__device__ __noinline__ void generic_backward(const KParams &P, int slot)
{
    RoiCtx c;
    roi_decode(P, slot, c);
    if (!c.valid) return;
    if (P.wait_fill) wait_for_predecessors();      // (see bwd_tasks: the gradients must be clean)
    const int C = P.C;
    const Strides4 fs = strides_of(P.feat_layout, C, c.L.H, c.L.W);
    float *grad = c.L.data + (long long)c.b * fs.s0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int PH = P.PH[h], PW = P.PW[h];
        const Strides4 os = strides_of(P.pool_layout, C, PH, PW);
        const float *gyp = P.pooled[h] + (long long)c.r * os.s0;
        const AxisGeom gy = axis_of(P, c, true, h, 0), gx = axis_of(P, c, true, h, 1);
        const int total = PH * PW * C;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            int ch, bin;
            if (P.pool_layout == RPOOL_NHWC) { ch = idx % C; bin = idx / C; }
            else { bin = idx % (PH * PW); ch = idx / (PH * PW); }
            const int ph = bin / PW, pw = bin % PW;
            const float g = __ldg(gyp + ch * os.s1 + ph * os.s2 + pw * os.s3);
            float *d = grad + ch * fs.s1;
            if (P.mode == RPOOL_COORD_CHAINER) {
                int y0, y1, x0, x1;
                float a0, a1, b0, b1;
                axis_sample(gy, P.mode, ph, 0, y0, y1, a0, a1);
                axis_sample(gx, P.mode, pw, 0, x0, x1, b0, b1);
                // (1-p)*(1-q)*gy: weight product first (:181-188)
                atomicAdd(d + y0 * fs.s2 + x0 * fs.s3, __fmul_rn(__fmul_rn(a0, b0), g));
                atomicAdd(d + y1 * fs.s2 + x0 * fs.s3, __fmul_rn(__fmul_rn(a1, b0), g));
                atomicAdd(d + y0 * fs.s2 + x1 * fs.s3, __fmul_rn(__fmul_rn(a0, b1), g));
                atomicAdd(d + y1 * fs.s2 + x1 * fs.s3, __fmul_rn(__fmul_rn(a1, b1), g));
            } else {
                const float count = (float)(gy.grid * gx.grid);
                for (int iy = 0; iy < gy.grid; ++iy) {
                    int y0, y1;
                    float hy, ly;
                    const bool vy = axis_sample(gy, P.mode, ph, iy, y0, y1, hy, ly);
                    for (int ix = 0; ix < gx.grid; ++ix) {
                        int x0, x1;
                        float hx, lx;
                        const bool vx = axis_sample(gx, P.mode, pw, ix, x0, x1, hx, lx);
                        if (!(vy && vx)) continue;
                        atomicAdd(d + y0 * fs.s2 + x0 * fs.s3, __fdiv_rn(__fmul_rn(g, __fmul_rn(hy, hx)), count));
                        atomicAdd(d + y0 * fs.s2 + x1 * fs.s3, __fdiv_rn(__fmul_rn(g, __fmul_rn(hy, lx)), count));
                        atomicAdd(d + y1 * fs.s2 + x0 * fs.s3, __fdiv_rn(__fmul_rn(g, __fmul_rn(ly, hx)), count));
                        atomicAdd(d + y1 * fs.s2 + x1 * fs.s3, __fdiv_rn(__fmul_rn(g, __fmul_rn(ly, lx)), count));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// table path, forward
// ---------------------------------------------------------------------------
// Task = (head, bin row ph, 128-channel slab), one warp each; a lane owns 4
// channels.  The row's bins are cut into chunks whose footprints fit a span of
// kSW window columns starting at column cx0 (chosen inside the image so that
// all kSW columns exist: no load is ever masked).  Per chunk:
//   column pass  V[s] = sum_j wy[j] * X[ylo + j][cx0 + s], s < kSW, kept in
//                registers; 128-bit loads through L1 (a window row is re-read by
//                the ~5 bin rows whose footprint contains it), the kSW loads of
//                one window row in flight together;
//   bin pass     out[pw] = sum_k wx[pw][k] * V[lo[pw] - cx0 + k].  Bins are
//                visited in order and their offsets lo[pw] - cx0 never decrease,
//                so the pass is a static sequence over the kSW offsets, each
//                running its (table-supplied) count of bins with compile-time
//                register operands; evict-first 128-bit stores.
// kC > 0 fixes the channel count at compile time (address arithmetic folds
// into immediates); kC == 0 reads it from the parameters.
__device__ __forceinline__ void fma4(float4 &a, float w, const float4 &v)
{
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ float4 mul4(float w, const float4 &v)
{
    return make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
}

template <int NX, int R>
__device__ __forceinline__ float4 taps(const float4 (&V)[kSW], const float4 w)
{
    // taps past the span carry zero weight (they lie beyond the window)
    float4 o = mul4(w.x, V[R]);
    if (NX > 1 && R + 1 < kSW) fma4(o, w.y, V[R + 1]);
    if (NX > 2 && R + 2 < kSW) fma4(o, w.z, V[R + 2]);
    if (NX > 3 && R + 3 < kSW) fma4(o, w.w, V[R + 3]);
    return o;
}

template <int NX, int R>
__device__ __forceinline__ void fwd_bins_at(const float4 (&V)[kSW], unsigned long long cnt,
                                            const float4 *&wp, float *&o_ptr, int C, bool active)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float4 o = taps<NX, R>(V, *wp);
        if (active) stg_stream128(o_ptr, o);
        ++wp;
        o_ptr += C;
    }
}

template <int NX>
__device__ __forceinline__ void fwd_bin_pass(const float4 (&V)[kSW], unsigned long long cnt,
                                             const float4 *wp, float *o_ptr, int C, bool active)
{
    fwd_bins_at<NX, 0>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 1>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 2>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 3>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 4>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 5>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 6>(V, cnt, wp, o_ptr, C, active);
    fwd_bins_at<NX, 7>(V, cnt, wp, o_ptr, C, active);
}

template <int kC>
__device__ __forceinline__ void fwd_load_row(float4 (&t)[kSW], const float *__restrict__ p, int C)
{
#pragma unroll
    for (int s = 0; s < kSW; ++s) t[s] = ldg_nc128(p + (kC ? s * kC : s * C));
}

template <int kC>
__device__ __forceinline__ void fwd_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl)
{
    const int C = kC ? kC : P.C;
    const int slabs = (C + 127) >> 7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int row_stride = c.L.W * C;
    const float *img = c.L.data + (size_t)c.b * c.L.H * row_stride;

    int ntask = 0;
    for (int h = 0; h < P.n_heads; ++h) ntask += P.PH[h] * slabs;
    // (split tail of the launch: this CTA's share of the tasks, whole bin rows first)
    int part;
    launch_place(P, part);
    const int parts = launch_parts(P);
    const int t_begin = ntask * part / parts, t_end = ntask * (part + 1) / parts;
    for (int t = t_begin + warp; t < t_end; t += nwarps) {
        int h = 0, tt = t;
        while (tt >= P.PH[h] * slabs) { tt -= P.PH[h] * slabs; ++h; }
        const int ph = tt / slabs;
        const int ch = (tt - ph * slabs) * 128 + lane * 4;
        const bool active = ch < C;
        const int PH = P.PH[h], PW = P.PW[h];
        const AxisTab &yt = ctl->hd[h].tab[0];
        const AxisTab &xt = ctl->hd[h].tab[1];
        float *out = P.pooled[h] + (((size_t)c.r * PH + ph) * PW) * C + ch;
        const int ny = yt.n[ph];
        const int NX = ctl->hd[h].nmax[1];
        if (ny == 0 || NX == 0) {
            if (active)
                for (int pw = 0; pw < PW; ++pw) stg_stream128(out + (size_t)pw * C, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const float4 wy = yt.w[ph];
        // lanes past the last channel of a partial slab read channel 0 and store nothing
        const float *rowp = img + (size_t)yt.lo[ph] * row_stride + (active ? ch : 0);
        const int nchunk = ctl->hd[h].nchunk;
        for (int q = 0; q < nchunk; ++q) {
            const int pa = ctl->hd[h].cstart[q];
            const unsigned long long cnt = ctl->hd[h].ccnt[q];
            const float *p = rowp + (size_t)ctl->hd[h].cx0[q] * C;
            // rows 0 and 1 of the footprint are requested together (16 loads in flight,
            // one exposed wait); a one-row footprint re-reads row 0 with weight zero
            float4 V[kSW], tmp[kSW];
            fwd_load_row<kC>(V, p, C);
            fwd_load_row<kC>(tmp, p + (ny > 1 ? row_stride : 0), C);
            const float wy1 = ny > 1 ? wy.y : 0.f;
#pragma unroll
            for (int s = 0; s < kSW; ++s) {
                V[s] = mul4(wy.x, V[s]);
                fma4(V[s], wy1, tmp[s]);
            }
            if (ny > 2) {
                fwd_load_row<kC>(tmp, p + 2 * (size_t)row_stride, C);
#pragma unroll
                for (int s = 0; s < kSW; ++s) fma4(V[s], wy.z, tmp[s]);
            }
            if (ny > 3) {
                fwd_load_row<kC>(tmp, p + 3 * (size_t)row_stride, C);
#pragma unroll
                for (int s = 0; s < kSW; ++s) fma4(V[s], wy.w, tmp[s]);
            }
            const float4 *wp = &xt.w[pa];
            float *o_ptr = out + (size_t)pa * C;
            if (NX <= 2) fwd_bin_pass<2>(V, cnt, wp, o_ptr, C, active);
            else if (NX == 3) fwd_bin_pass<3>(V, cnt, wp, o_ptr, C, active);
            else fwd_bin_pass<4>(V, cnt, wp, o_ptr, C, active);
        }
    }
}

__device__ __forceinline__ void ctx_from_record(const KParams &P, const BlockCtl *ctl, RoiCtx &c)
{
    c.r = ctl->r; c.lvl = ctl->lvl; c.b = ctl->b;
    c.L = P.lvl[c.lvl];
    c.valid = (ctl->flags & kRecValid) != 0;
    c.fast_ok = (ctl->flags & kRecShape) != 0;
}

__global__ void __launch_bounds__(kMaxThreads, RPOOL_MIN_BLOCKS)
rpool_forward_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);
    wait_for_predecessors();      // launched early, in the plan kernel's tail: the records must be there
    allow_dependents_early();     // a zero fill queued behind this launch may use the tail

    load_record(P, ctl);
    RoiCtx c;
    ctx_from_record(P, ctl, c);
    const int need = kRecValid | kRecShape | kRecFits;
    if ((ctl->flags & need) != need || P.force_path == kPathGeneric || !pointers_aligned(P, c.L)) {
        int part;
        launch_place(P, part);
        if (part == 0) generic_forward(P);
        return;
    }
    if (P.C == 256) fwd_tasks<256>(P, c, ctl);
    else fwd_tasks<0>(P, c, ctl);
}

// ---------------------------------------------------------------------------
// table path, backward
// ---------------------------------------------------------------------------
// The adjoint as a gather, so that no two warps ever add to the same place:
// task = (window row y, 128-channel slab), one warp each.  Per task and head:
//   row pass     Z[pw] = sum over bin rows ph covering y of wyT[y][ph] * gy[ph][pw]
//                (128-bit loads, kZ independent loads per bin row), written to the
//                warp's strip in shared memory;
//   column pass  G[x] = sum over bins pw covering x of wxT[x][pw] * Z[pw];
//                one 128-bit vector reduction per window cell into the dense gradient.
//                Like the forward bin pass it walks the bins in order with the span
//                offsets static, accumulating G[kSW] in registers.
// The transposed y table (dense, kExt x kPBwd) is built in shared memory from the
// forward footprint table.  The gy bin rows a warp's next task is the first to need
// are pulled into L2 with one bulk prefetch (TMA unit) per bin row while the column
// pass of the current task runs.
// A CTA of a launch that started in the zero fill's tail: wait for the fill (no-op when some warp
// already did), then let the next launch go.  Every such CTA passes here before it exits, so the
// next launch can never start adding before the fill is complete.
__device__ __forceinline__ void bwd_release(const KParams &P)
{
    if (P.wait_fill) {
        wait_for_predecessors();
        allow_dependents_early();
    }
}

struct TTab {
    int pa[kExt], pb[kExt];  // covering bin rows of window row i: [pa, pb)
};

// Weight of bin row ph on window row y (absolute): component y - lo[ph] of the forward
// footprint entry, read straight from the record (0 outside the footprint).
__device__ __forceinline__ float wy_of(const AxisTab &yt, int ph, int y)
{
    const int k = y - yt.lo[ph];
    return (k >= 0 && k < yt.n[ph]) ? reinterpret_cast<const float *>(&yt.w[ph])[k] : 0.f;
}

__device__ __forceinline__ void build_ttabs(const KParams &P, const BlockCtl *ctl, TTab *tt)
{
    // per head, for the y axis: window row -> the contiguous run of bin rows covering it
    // (footprints start at non-decreasing rows, so the run is an interval)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ext = ctl->wmax[0] - ctl->wmin[0] + 1;
    for (int h = 0; h < P.n_heads; ++h)
        for (int i = tid; i < ext; i += nt) { tt[h].pa[i] = 0x7fffffff; tt[h].pb[i] = 0; }
    __syncthreads();
    int base = 0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int p = tid - base;
        if (p >= 0 && p < P.PH[h]) {
            const AxisTab &t = ctl->hd[h].tab[0];
            TTab &T = tt[h];
            const int n = t.n[p], row0 = t.lo[p] - ctl->wmin[0];
            for (int k = 0; k < n; ++k) {
                atomicMin(&T.pa[row0 + k], p);
                atomicMax(&T.pb[row0 + k], p + 1);
            }
        }
        base += P.PH[h];
    }
    __syncthreads();
}

constexpr int kZ = 7;  // bins per register chunk of the row pass

template <int NX, int R>
__device__ __forceinline__ void bwd_bins_at(float4 (&G)[kSW], unsigned long long cnt, const float4 *&wp,
                                            uint32_t &zp)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float4 w = *wp;
        const float4 z = lds128(zp);
        fma4(G[R], w.x, z);
        if (NX > 1 && R + 1 < kSW) fma4(G[R + 1], w.y, z);
        if (NX > 2 && R + 2 < kSW) fma4(G[R + 2], w.z, z);
        if (NX > 3 && R + 3 < kSW) fma4(G[R + 3], w.w, z);
        ++wp;
        zp += 512u;
    }
}

template <int NX>
__device__ __forceinline__ void bwd_col_pass(float4 (&G)[kSW], unsigned long long cnt, const float4 *wp,
                                             uint32_t zp)
{
    bwd_bins_at<NX, 0>(G, cnt, wp, zp);
    bwd_bins_at<NX, 1>(G, cnt, wp, zp);
    bwd_bins_at<NX, 2>(G, cnt, wp, zp);
    bwd_bins_at<NX, 3>(G, cnt, wp, zp);
    bwd_bins_at<NX, 4>(G, cnt, wp, zp);
    bwd_bins_at<NX, 5>(G, cnt, wp, zp);
    bwd_bins_at<NX, 6>(G, cnt, wp, zp);
    bwd_bins_at<NX, 7>(G, cnt, wp, zp);
}

// kC as in the forward kernel; kExact: PW is a multiple of kZ and C of 128, so
// no lane and no chunk position is ever masked.
// kDet: the deterministic variant's window pass (each RoI's contribution goes to its private window
// with plain stores).  A compile-time switch and a kernel of its own, so that the bookkeeping of the
// first-touch stores costs the atomic kernel no registers.
template <int kC, bool kExact, bool kDet>
__device__ __forceinline__ void bwd_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl,
                                          const TTab *tt, uint32_t strip, float *det_win)
{
    const int C = kC ? kC : P.C;
    const int slabs = (C + 127) >> 7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int y0 = ctl->wmin[0], x0 = ctl->wmin[1];
    const int Hc = ctl->wmax[0] - y0 + 1;
    const int det_wc = ctl->wmax[1] - x0 + 1;
    float *img = c.L.data + (size_t)c.b * c.L.H * c.L.W * C;

    const int ntask = Hc * slabs;
    int t_begin = 0, t_end = ntask;
    if (kTailPartsBwd > 1) {
        int part;
        launch_place(P, part);
        const int parts = launch_parts(P);
        t_begin = ntask * part / parts;
        t_end = ntask * (part + 1) / parts;
    }
    bool waited = false;
    for (int t = t_begin + warp; t < t_end; t += nwarps) {
        const int i = t / slabs;
        const int ch = (t - i * slabs) * 128 + lane * 4;
        const bool active = kExact || ch < C;
        float *grow_img = img + (size_t)(y0 + i) * c.L.W * C + ch;   // column 0 of the window row
        bool ahead_done = false;
        unsigned long long wrote = 0;      // deterministic: window columns of this row already stored
        auto row_ahead = [&]() {
            if (P.prefetch <= -2 && lane == 0 && t - i * slabs == 0) {
                // row-ahead mode: pull the gy bin rows that window row i + lead is the first to
                // need into L2 (one bulk prefetch per bin row: PW * C floats, contiguous), so
                // that the row pass of that task finds them at L2 latency.  Short leads only:
                // measured on cfg 1, lead 2 (this warp's next task) 0.2265 ms, lead 1 0.2287,
                // lead 3 0.2378, none 0.2431, the whole RoI at CTA start 0.2559
                const int lead = -P.prefetch - 1;
                const int in = i + lead;
                if (in < Hc) {
                    for (int h = 0; h < P.n_heads; ++h) {
                        const int pb_n = tt[h].pb[in];
                        int pa_n = tt[h].pa[in];
                        const int seen = tt[h].pb[in - 1];      // (in >= 1: lead >= 1)
                        pa_n = pa_n < seen ? seen : pa_n;
                        const size_t row_floats = (size_t)P.PW[h] * C;
                        for (int ph = pa_n; ph < pb_n; ++ph)
                            prefetch_l2_bulk(P.pooled[h] + ((size_t)c.r * P.PH[h] + ph) * row_floats,
                                             (unsigned)(row_floats * 4));
                    }
                }
            }
        };
        for (int h = 0; h < P.n_heads; ++h) {
            const TTab &Ty = tt[h];
            const int PH = P.PH[h], PW = P.PW[h];
            const int pa = Ty.pa[i], pb = Ty.pb[i];
            if (pa >= pb) continue;  // a window row between this head's footprints
            // ---- row pass
            const float *gbase = P.pooled[h] + ((size_t)c.r * PH + pa) * PW * C + ch;
            const int gstep = PW * C;
            for (int pw0 = 0; pw0 < PW; pw0 += kZ) {
                float4 Z[kZ], v[kZ];
                const float *g = gbase + (size_t)pw0 * C;
                const AxisTab &ytab = ctl->hd[h].tab[0];
                const int yrow = y0 + i;
                auto load_bins = [&](float4 (&dst)[kZ], const float *src) {
#pragma unroll
                    for (int k = 0; k < kZ; ++k) {
                        if (kExact) {
                            dst[k] = ldg_nc128(src + (kC ? k * kC : k * C));
                        } else {
                            dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (active && pw0 + k < PW) dst[k] = ldg_nc128(src + (size_t)k * C);
                        }
                    }
                };
                // the first two covering bin rows are requested together (2 * kZ loads in
                // flight, one exposed wait); a single covering row is re-read with weight 0
                const bool two = pa + 1 < pb;
                load_bins(Z, g);
                load_bins(v, g + (two ? gstep : 0));
                {
                    const float w0 = wy_of(ytab, pa, yrow), w1 = two ? wy_of(ytab, pa + 1, yrow) : 0.f;
#pragma unroll
                    for (int k = 0; k < kZ; ++k) {
                        Z[k] = mul4(w0, Z[k]);
                        fma4(Z[k], w1, v[k]);
                    }
                }
                g += 2 * (size_t)gstep;
                for (int ph = pa + 2; ph < pb; ++ph, g += gstep) {
                    const float w = wy_of(ytab, ph, yrow);
                    load_bins(v, g);
#pragma unroll
                    for (int k = 0; k < kZ; ++k) fma4(Z[k], w, v[k]);
                }
#pragma unroll
                for (int k = 0; k < kZ; ++k)
                    if (kExact || pw0 + k < PW) sts128(strip + (uint32_t)(pw0 + k) * 512u, Z[k]);
            }
            __syncwarp();
            // (issued once this task's own gy loads have landed, so that it does not queue
            // ahead of them)
            if (!ahead_done) { row_ahead(); ahead_done = true; }
            // ---- column pass: bins in order, span offsets static (as in the forward bin pass)
            const AxisTab &xt = ctl->hd[h].tab[1];
            const int NX = ctl->hd[h].nmax[1];
            const int nchunk = ctl->hd[h].nchunk;
            for (int q = 0; q < nchunk; ++q) {
                const int pa_q = ctl->hd[h].cstart[q];
                float4 G[kSW];
#pragma unroll
                for (int s = 0; s < kSW; ++s) G[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 *wp = &xt.w[pa_q];
                uint32_t zp = strip + (uint32_t)pa_q * 512u;
                const unsigned long long cnt = ctl->hd[h].ccnt[q];
                if (NX <= 2) bwd_col_pass<2>(G, cnt, wp, zp);
                else if (NX == 3) bwd_col_pass<3>(G, cnt, wp, zp);
                else bwd_col_pass<4>(G, cnt, wp, zp);
                const unsigned m = ctl->hd[h].cmask[q];
                if (!kDet) {
                    // (launched early, in the zero fill's tail: everything up to here only read gy
                    // and the plan; the gradients must be clean before the first reduction)
                    if (P.wait_fill && !waited) { bwd_release(P); waited = true; }
                    float *gp = grow_img + (size_t)ctl->hd[h].cx0[q] * C;
#pragma unroll
                    for (int s = 0; s < kSW; ++s)
                        if (active && ((m >> s) & 1u)) red_add_v4(gp + (kC ? s * kC : s * C), G[s]);
                } else {
                    // deterministic: this RoI's private window (this warp is the only writer of its
                    // row and slab; heads and chunks in order).  The first contribution to a cell is a
                    // plain store, later ones (chunk overlaps, the second pooled size) read-modify-write;
                    // windows wider than 64 columns were zero-filled and always read-modify-write.
                    const int col0 = ctl->hd[h].cx0[q] - x0;
                    float *gp = det_win + ((size_t)i * det_wc + col0) * C + ch;
#pragma unroll
                    for (int s = 0; s < kSW; ++s) {
                        if (active && ((m >> s) & 1u)) {
                            float *a = gp + (kC ? s * kC : s * C);
                            float4 v = G[s];
                            if (det_wc > 64 || ((wrote >> (col0 + s)) & 1ull)) {
                                const float4 o = ldg_cg128(a);
                                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                            }
                            stg128(a, v);
                        }
                    }
                    // (a span starts left of the window when the window touches the map's right edge)
                    if (det_wc <= 64) wrote |= col0 >= 0 ? (unsigned long long)m << col0 : (unsigned long long)m >> -col0;
                }
            }
            __syncwarp();
        }
        if (kDet && det_wc <= 64 && active) {
            // deterministic: the cells of this window row that received nothing
            float *gp = det_win + (size_t)i * det_wc * C + ch;
            for (int col = 0; col < det_wc; ++col)
                if (!((wrote >> col) & 1ull)) stg128(gp + (size_t)col * C, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
}

template <bool kDet>
__device__ __forceinline__ void backward_body(const KParams &P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);
    // The launch for the next pooled size may use this one's tail.  It does not wait for anything
    // itself, so when THIS launch started early (in the zero fill's tail, wait_fill) a CTA gives
    // the signal only once it has seen the fill complete -- see bwd_release.
    if (!P.wait_fill && !kDet) allow_dependents_early();
    const int kCtlBytes = (rec_bytes(P.n_heads) + 127) & ~127;   // only this launch's head parts are loaded
    constexpr int kTTabBytes = (sizeof(TTab) + 127) & ~127;
    TTab *tt = reinterpret_cast<TTab *>(smem_raw + kCtlBytes);

    // pull the upstream gradient of the RoI scheduled `prefetch` CTAs later (0: this
    // CTA's own, overlapping the table construction below) into L2
    if (P.prefetch >= 0 && P.pool_layout == RPOOL_NHWC && (P.C & 3) == 0) {
        const int nb = blockIdx.x + P.prefetch;
        if (nb < P.R) {
            int row = threadIdx.x;
            for (int h = 0; h < P.n_heads; ++h) {
                if (row >= 0 && row < P.PH[h]) {
                    const int r2 = P.order[P.reverse ? P.R - 1 - nb : nb];
                    const size_t row_floats = (size_t)P.PW[h] * P.C;
                    prefetch_l2_bulk(P.pooled[h] + ((size_t)r2 * P.PH[h] + row) * row_floats,
                                     (unsigned)(row_floats * 4));
                }
                row -= P.PH[h];
            }
        }
    }

    load_record(P, ctl);
    if (kDet) {
        // launched in the scan kernel's tail: its offsets from here on.  The gather launch queued behind
        // this one lists its windows (rectangles: the scan's output) as soon as every CTA has got here.
        wait_for_predecessors();
        allow_dependents_early();
    }
    RoiCtx c;
    ctx_from_record(P, ctl, c);
    if (!c.valid) { bwd_release(P); return; }
    const int need = kRecShape | kRecFits;
    bool table_ok = (ctl->flags & need) == need && P.force_path != kPathGeneric && pointers_aligned(P, c.L);
    for (int h = 0; h < P.n_heads; ++h) table_ok = table_ok && P.PH[h] <= kPBwd && P.PW[h] <= kPBwd;
    const int x0 = ctl->wmin[1], x1 = ctl->wmax[1];
    const int y0 = ctl->wmin[0], y1 = ctl->wmax[0];
    if (!table_ok || y1 - y0 >= kExt) {
        int part = 0;
        if (kTailPartsBwd > 1) launch_place(P, part);
        if (kDet) { if (threadIdx.x == 0) atomicExch(P.det_err, 1); }   // no ordered generic path
        else if (part == 0) generic_backward(P, launch_slot(P));
        bwd_release(P);
        return;
    }
    if (x1 < x0 || y1 < y0) { bwd_release(P); return; }
    float *det_win = nullptr;
    if (kDet) {
        // this RoI's private window in the scratch buffer: zero it, then the tasks
        // below add into it with plain stores
        const int *rc = P.det_rects + 4 * (size_t)launch_slot(P);
        const unsigned long long off = P.det_woff[launch_slot(P)];
        const unsigned long long n = (unsigned long long)(x1 - x0 + 1) * (y1 - y0 + 1) * P.C;
        if (rc[0] != x0 || rc[1] != y0 || rc[2] != x1 || rc[3] != y1 || off + n > P.det_scratch_floats) {
            if (threadIdx.x == 0) atomicExch(P.det_err, 2);
            return;
        }
        det_win = P.det_scratch + off;
        if (x1 - x0 + 1 > 64)      // (narrower windows: first-touch stores, see bwd_tasks)
            for (unsigned long long i = threadIdx.x * 4ull; i < n; i += blockDim.x * 4ull)
                stg128(det_win + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    build_ttabs(P, ctl, tt);
    if (P.prefetch <= -2 && P.pool_layout == RPOOL_NHWC && (P.C & 3) == 0) {
        // row-ahead mode, head of the window: the bin rows of window rows [0, lead)
        const int lead = -P.prefetch - 1;
        int row = threadIdx.x;
        // (a CTA of the split tail starts at its own first window row)
        const int Hc = y1 - y0 + 1;
        int first = 0;
        if (kTailPartsBwd > 1) {
            int part;
            launch_place(P, part);
            first = (Hc * ((P.C + 127) >> 7) * part / launch_parts(P)) / ((P.C + 127) >> 7);
        }
        for (int h = 0; h < P.n_heads; ++h) {
            if (row >= 0 && row < P.PH[h]) {
                const int last = (first + lead < Hc ? first + lead : Hc) - 1;
                int hi = 0, lo = 0x7fffffff;
                for (int i = first; i <= last; ++i) {
                    hi = tt[h].pb[i] > hi ? tt[h].pb[i] : hi;
                    lo = tt[h].pa[i] < lo ? tt[h].pa[i] : lo;
                }
                if (row < hi && row >= lo) {
                    const size_t row_floats = (size_t)P.PW[h] * P.C;
                    prefetch_l2_bulk(P.pooled[h] + ((size_t)c.r * P.PH[h] + row) * row_floats,
                                     (unsigned)(row_floats * 4));
                }
            }
            row -= P.PH[h];
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t strip = smem_u32(smem_raw + kCtlBytes + P.n_heads * kTTabBytes) +
                           (uint32_t)warp * (uint32_t)P.strip_cols * 512u + (uint32_t)lane * 16u;
    bool exact = (P.C == 256);
    for (int h = 0; h < P.n_heads; ++h) exact = exact && (P.PW[h] % kZ == 0);
    if (exact) bwd_tasks<256, true, kDet>(P, c, ctl, tt, strip, det_win);
    else bwd_tasks<0, false, kDet>(P, c, ctl, tt, strip, det_win);
    bwd_release(P);
}

__global__ void __launch_bounds__(kMaxThreads, RPOOL_MIN_BLOCKS)
rpool_backward_kernel(const __grid_constant__ KParams P)
{
    backward_body<false>(P);
}

// the deterministic variant's window pass (KParams::det = 1)
__global__ void __launch_bounds__(kMaxThreads, RPOOL_MIN_BLOCKS)
rpool_backward_det_window_kernel(const __grid_constant__ KParams P)
{
    backward_body<true>(P);
}

// ---------------------------------------------------------------------------
// table path, backward, staged variant (atomic path, one pooled size per launch)
// ---------------------------------------------------------------------------
// Opt-in (opt.backward_variant = 2); measured slower than the rows kernel, kept as the reference point for
// a copy-engine design of this pass (profiles/r02_experiments.log, r02w_staged_ncu_summary.txt).
// One persistent CTA per SM.  The upstream gradient of a (RoI, 128-channel slab) item --
// PH x PW pieces of 512 bytes, or PH contiguous bin rows when C = 128 -- is copied into shared
// memory by the TMA unit (bulk copies completing on an mbarrier), two items deep, together with
// the RoI's record; every byte of gy crosses L2 -> SM once (the rows kernel re-reads each bin row
// for the 2-3 window rows it covers) and no load occupies a register while it is in flight.  The
// CTA's warps take (item, window row) tasks from a shared counter, in order; a task sums its
// covering bin rows straight from the staged item while it walks the bins (no strip), accumulates
// the window cells of a chunk in registers as the rows kernel does and reduces them into the
// gradient.  The warp that finishes the last task of an item refills its buffer with the item two
// further on.
// What the measurements say (B200): 512-byte bulk copies are bound by the copy engine's request
// rate (~8 bytes per clock and SM: 0.42 ms for configs[1] against 0.19 ms); with one 7 KB copy per
// bin row (C = 128) the staging keeps up, and the 8 warps that fit beside two staged items become
// the limit -- issue slots 40 % busy at 2 warps per scheduler, no memory stall left (0.160 ms
// against 0.125 ms); 16 warps leave no buffer loading in the background (0.177 ms).  One RoI's gy
// (196 KB at C = 256) is the whole shared memory: data in flight plus the rows live under 16
// concurrent tasks do not fit.
constexpr int kStWarps = 8;
constexpr int kStThreads = kStWarps * 32;
constexpr int kStMaxSlots = 128;                  // schedule slots one CTA may own
constexpr int kStMaxSlabs = 4;                    // C <= 512
constexpr int kStMaxItems = kStMaxSlots * kStMaxSlabs;
constexpr int kStCtlBytes = 6144;                 // control block
constexpr int kStRecBytes = 2048;                 // one record (header + one head part), padded

struct StagedCtl {
    unsigned long long full[2];        // mbarriers: item data + record have landed in buffer b
    int loaded[2];                     // item whose copies were issued into buffer b
    int done[2];                       // finished tasks of the item in buffer b
    int next_task, n_items, total_tasks, pad_;
    int pref[kStMaxItems + 1];         // first task of every item
    unsigned short item[kStMaxItems];  // owned slot index * 4 + slab
    int r[kStMaxSlots];                // RoI of every owned slot
    short hc[kStMaxSlots];             // window rows of every owned slot (0: nothing to do here)
    unsigned char gen[kStMaxSlots];    // the slot needs the generic path
};
static_assert(sizeof(StagedCtl) <= kStCtlBytes, "staged control block");
static_assert(kRecHeader + sizeof(HeadCtl) <= kStRecBytes, "staged record slot");

template <int NX, int R>
__device__ __forceinline__ void st_bins_at(float4 (&G)[kSW], unsigned long long cnt, const float4 *&wp,
                                           uint32_t &zp, uint32_t o1, uint32_t rs, int nr,
                                           float w0, float w1, float w2, float w3)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float4 w = *wp;
        const float4 a0 = lds128(zp);
        const float4 a1 = lds128(zp + o1);       // (a single covering row: row 0 again, weight 0)
        float4 z = mul4(w0, a0);
        fma4(z, w1, a1);
        if (nr > 2) fma4(z, w2, lds128(zp + 2 * rs));
        if (nr > 3) fma4(z, w3, lds128(zp + 3 * rs));
        fma4(G[R], w.x, z);
        if (NX > 1 && R + 1 < kSW) fma4(G[R + 1], w.y, z);
        if (NX > 2 && R + 2 < kSW) fma4(G[R + 2], w.z, z);
        if (NX > 3 && R + 3 < kSW) fma4(G[R + 3], w.w, z);
        ++wp;
        zp += 512u;
    }
}

template <int NX>
__device__ __forceinline__ void st_col_pass(float4 (&G)[kSW], unsigned long long cnt, const float4 *wp,
                                            uint32_t zp, uint32_t o1, uint32_t rs, int nr,
                                            float w0, float w1, float w2, float w3)
{
    st_bins_at<NX, 0>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 1>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 2>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 3>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 4>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 5>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 6>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
    st_bins_at<NX, 7>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
}

// One task: window row i of the item staged at `buf` (record `ctl`), channels slab * 128 + lane * 4.
template <int kC>
__device__ __forceinline__ void staged_task(const KParams &P, const BlockCtl *ctl, uint32_t buf, int i, int slab,
                                            int lane, bool &waited)
{
    const int C = kC ? kC : P.C;
    const int PH = P.PH[0], PW = P.PW[0];
    const HeadCtl &hd = ctl->hd[0];
    const AxisTab &yt = hd.tab[0];
    const AxisTab &xt = hd.tab[1];
    const int y = ctl->wmin[0] + i;
    // the bin rows whose footprint holds window row y: an interval (footprints start at non-decreasing rows)
    const bool cov = lane < PH && yt.lo[lane] <= y && y < yt.lo[lane] + yt.n[lane];
    const unsigned m = __ballot_sync(0xffffffffu, cov);
    if (!m) return;
    const int pa = __ffs(m) - 1, pb = 32 - __clz(m);
    const LevelDev L = P.lvl[ctl->lvl];
    float *grow = L.data + (((size_t)ctl->b * L.H + y) * L.W) * C + slab * 128 + lane * 4;
    const uint32_t rs = (uint32_t)PW * 512u;
    const int NX = hd.nmax[1];
    const int nchunk = hd.nchunk;
    for (int q = 0; q < nchunk; ++q) {
        const int pa_q = hd.cstart[q];
        const unsigned long long cnt = hd.ccnt[q];
        float4 G[kSW];
#pragma unroll
        for (int s = 0; s < kSW; ++s) G[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = pa; g < pb; g += 4) {       // (more than four covering rows: tiny RoIs)
            const int nr = pb - g < 4 ? pb - g : 4;
            const float w0 = wy_of(yt, g, y);
            const float w1 = nr > 1 ? wy_of(yt, g + 1, y) : 0.f;
            const float w2 = nr > 2 ? wy_of(yt, g + 2, y) : 0.f;
            const float w3 = nr > 3 ? wy_of(yt, g + 3, y) : 0.f;
            const uint32_t zp = buf + (uint32_t)(g * PW + pa_q) * 512u + (uint32_t)lane * 16u;
            const float4 *wp = &xt.w[pa_q];
            const uint32_t o1 = nr > 1 ? rs : 0u;
            if (NX <= 2) st_col_pass<2>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
            else if (NX == 3) st_col_pass<3>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
            else st_col_pass<4>(G, cnt, wp, zp, o1, rs, nr, w0, w1, w2, w3);
        }
        // (launched in the zero fill's tail: the gradients must be clean before the first reduction)
        if (P.wait_fill && !waited) { wait_for_predecessors(); waited = true; }
        const unsigned cm = hd.cmask[q];
        float *gp = grow + (size_t)hd.cx0[q] * C;
#pragma unroll
        for (int s = 0; s < kSW; ++s)
            if ((cm >> s) & 1u) red_add_v4(gp + (kC ? s * kC : s * C), G[s]);
    }
}

__device__ __forceinline__ int staged_slot(const KParams &P, int k)
{
    const int idx = (int)blockIdx.x + k * (int)gridDim.x;      // launch order
    return P.reverse ? P.R - 1 - idx : idx;
}

// Issues the copies of item j into buffer b (one warp, converged).
__device__ __forceinline__ void staged_issue(const KParams &P, StagedCtl *sc, uint32_t smem_base, int item_bytes,
                                             int j, int b, int lane)
{
    const int code = sc->item[j];
    const int k = code >> 2, slab = code & 3;
    const int PHW = P.PH[0] * P.PW[0];
    const uint32_t bar = smem_u32(&sc->full[b]);
    const uint32_t rec_dst = smem_base + kStCtlBytes + (uint32_t)b * kStRecBytes;
    const uint32_t buf = smem_base + kStCtlBytes + 2u * kStRecBytes + (uint32_t)b * (uint32_t)item_bytes;
    const unsigned char *rec = P.recs + (size_t)staged_slot(P, k) * P.rec_stride;
    // (the buffer's previous readers are done: ordered by the done counter; generic-proxy reads
    // before async-proxy writes)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(kRecHeader + sizeof(HeadCtl)) + (uint32_t)PHW * 512u);
    __syncwarp();
    if (lane == 0) bulk_g2s(rec_dst, rec, kRecHeader, bar);
    if (lane == 1) bulk_g2s(rec_dst + kRecHeader, rec + kRecHeader + (size_t)P.rec_head * sizeof(HeadCtl),
                            (uint32_t)sizeof(HeadCtl), bar);
    const float *src = P.pooled[0] + (size_t)sc->r[k] * PHW * P.C + slab * 128;
    if (P.C == 128) {
        // one slab wide: a bin row is contiguous
        if (lane < P.PH[0])
            bulk_g2s(buf + (uint32_t)(lane * P.PW[0]) * 512u, src + (size_t)lane * P.PW[0] * 128,
                     (uint32_t)P.PW[0] * 512u, bar);
    } else {
        for (int pcs = lane; pcs < PHW; pcs += 32)
            bulk_g2s(buf + (uint32_t)pcs * 512u, src + (size_t)pcs * P.C, 512u, bar);
    }
    // the upstream gradient of the RoI this CTA takes up two items further on: into L2 meanwhile
    if (slab == 0 && k + 2 < P.staged_slots) {
        const int slot2 = staged_slot(P, k + 2);
        if (slot2 >= 0 && slot2 < P.R && sc->hc[k + 2] > 0 && lane < P.PH[0]) {
            const size_t row_floats = (size_t)P.PW[0] * P.C;
            prefetch_l2_bulk(P.pooled[0] + ((size_t)sc->r[k + 2] * P.PH[0] + lane) * row_floats,
                             (unsigned)(row_floats * 4));
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        *reinterpret_cast<volatile int *>(&sc->loaded[b]) = j;
    }
}

__global__ void __launch_bounds__(kStThreads, 1)
rpool_backward_staged_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StagedCtl *sc = reinterpret_cast<StagedCtl *>(smem_raw);
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int PHW = P.PH[0] * P.PW[0];
    const int item_bytes = PHW * 512;
    const int slabs = P.C >> 7;
    if (!P.wait_fill) allow_dependents_early();

    // ---- this CTA's schedule slots: window rows, RoI index, which path
    int n_my = 0;
    for (int k = 0; k < P.staged_slots; ++k) {
        const int idx = (int)blockIdx.x + k * (int)gridDim.x;
        if (idx < P.R) n_my = k + 1;
    }
    for (int k = tid; k < n_my; k += kStThreads) {
        const unsigned char *rec = P.recs + (size_t)staged_slot(P, k) * P.rec_stride;
        const int4 a = __ldg(reinterpret_cast<const int4 *>(rec));        // wmin[0], wmin[1], wmax[0], wmax[1]
        const int4 f = __ldg(reinterpret_cast<const int4 *>(rec) + 1);    // r, lvl, b, flags
        const bool valid = (f.w & kRecValid) != 0;
        const int need = kRecShape | kRecFits;
        const bool table_ok = (f.w & need) == need && P.force_path != kPathGeneric &&
                              pointers_aligned(P, P.lvl[f.y]);
        int hc = 0;
        if (valid && table_ok && a.w >= a.y && a.z >= a.x) hc = a.z - a.x + 1;
        sc->hc[k] = (short)(hc > 32767 ? 32767 : hc);
        sc->gen[k] = (valid && !table_ok) ? 1 : 0;
        sc->r[k] = f.x;
    }
    if (tid == 0) {
        mbar_init(smem_u32(&sc->full[0]), 1);
        mbar_init(smem_u32(&sc->full[1]), 1);
        mbar_init_fence();
        sc->loaded[0] = sc->loaded[1] = -1;
        sc->done[0] = sc->done[1] = 0;
        sc->next_task = 0;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0, pref = 0;
        for (int k = 0; k < n_my; ++k) {
            const int hc = sc->hc[k];
            if (hc <= 0) continue;
            for (int sl = 0; sl < slabs; ++sl) {
                sc->item[n] = (unsigned short)(k * 4 + sl);
                sc->pref[n] = pref;
                pref += hc;
                ++n;
            }
        }
        sc->pref[n] = pref;
        sc->n_items = n;
        sc->total_tasks = pref;
    }
    __syncthreads();
    // ---- RoIs outside the table path (rare): tap by tap, the whole CTA
    for (int k = 0; k < n_my; ++k)
        if (sc->gen[k]) generic_backward(P, staged_slot(P, k));
    const int n_items = sc->n_items, total = sc->total_tasks;
    if (warp == 0) {
        if (n_items > 0) staged_issue(P, sc, smem_base, item_bytes, 0, 0, lane);
        if (n_items > 1) staged_issue(P, sc, smem_base, item_bytes, 1, 1, lane);
    }

    // ---- tasks
    int j = 0;
    bool waited = false;
    for (;;) {
        int seq = 0;
        if (lane == 0) seq = atomicAdd(&sc->next_task, 1);
        seq = __shfl_sync(0xffffffffu, seq, 0);
        if (seq >= total) break;
        while (seq >= sc->pref[j + 1]) ++j;
        const int i = seq - sc->pref[j];
        const int b = j & 1;
        if (lane == 0)
            while (*reinterpret_cast<volatile int *>(&sc->loaded[b]) != j) __nanosleep(40);
        __syncwarp();
        const uint32_t bar = smem_u32(&sc->full[b]);
        while (!mbar_try_wait(bar, (uint32_t)((j >> 1) & 1))) {}
        const BlockCtl *ctl = reinterpret_cast<const BlockCtl *>(smem_raw + kStCtlBytes + b * kStRecBytes);
        const uint32_t buf = smem_base + kStCtlBytes + 2u * kStRecBytes + (uint32_t)b * (uint32_t)item_bytes;
        const int slab = sc->item[j] & 3;
        if (P.C == 256) staged_task<256>(P, ctl, buf, i, slab, lane, waited);
        else staged_task<0>(P, ctl, buf, i, slab, lane, waited);
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            const int d = atomicAdd(&sc->done[b], 1);
            last = (d == sc->pref[j + 1] - sc->pref[j] - 1) ? 1 : 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            // every task of item j is finished: its buffer takes item j + 2
            if (lane == 0) { sc->done[b] = 0; __threadfence_block(); }
            __syncwarp();
            if (j + 2 < n_items) staged_issue(P, sc, smem_base, item_bytes, j + 2, b, lane);
        }
    }
    bwd_release(P);
}

// ---------------------------------------------------------------------------
// plan: level assignment + stable (image, level) binning + every RoI's record
// ---------------------------------------------------------------------------
// One warp per input RoI i, four to a CTA (all RoIs in parallel, one wave for the sizes
// of BASELINE.json; nothing in a RoI's plan needs more than warp-wide cooperation, so
// there is no CTA barrier on the path).  The warp finds its RoI's slot in the launch schedule -- the stable
// rank of (key_i, i) among all RoIs, key = (image, level) in the order the schedule
// asks for -- and writes the RoI's record (footprint tables + chunking, for the
// forward geometry and, in RPOOL_COORD_CHAINER mode, a second one for the backward
// geometry) at that slot.  There is no single-CTA sort on the critical path:
//   R <= kPlanSingle one launch: every CTA derives the keys of all RoIs once into shared
//                    memory (R / 128 RoIs per thread, 20 bytes each, L1/L2 hits) and its
//                    warps rank their RoIs from there;
//   R  > kPlanSingle rpool_keys_kernel first writes the keys and one histogram per
//                    block of kKeyBlock RoIs; a plan CTA then sums the histogram
//                    entries that precede (key_i, block_i) and ranks its RoI inside
//                    its own block.
struct PlanParams {
    const float *rois;
    int R;
    int roi_format;
    const int *given_levels;  // may be null
    const float *given_levels_f32;  // may be null
    float thr[kMaxLevels];
    int n_thr;
    int k_min;
    int n_levels;
    int n_images;   // max over levels
    int order_mode; // rpool_schedule
    int K;          // number of keys
    int by_image;   // keys include the image index (n_images * n_levels <= kPlanMaxKeys)
    int n_blocks;   // key blocks (0: single launch)
    int *levels;    // out: level of every RoI
    int *order;     // out: schedule slot -> RoI
    int *keys;      // out (two-launch mode): key of every RoI
    int *bh;        // out (two-launch mode): [n_blocks][K] histograms
    int *gstart;    // out: first schedule slot of every (image, level) key, [K..] = R
    int *rflags;    // out: RPOOL_FLAG_* of every RoI
    int *det_err;   // cleared here; raised by the deterministic backward
};

__device__ __forceinline__ int level_from_area(float y1, float x1, float y2, float x2,
                                               const float *thr, int n_thr, int k_min)
{
    // area = (y2 - y1) * (x2 - x1) in float32; the level is the count of
    // thresholds not above it (bit-exact image of the reference's
    // floor(4 + log2(sqrt(area)/224 + 1e-6)), multilevel_region_proposal_network.py:24-30)
    const float area = __fmul_rn(__fsub_rn(y2, y1), __fsub_rn(x2, x1));
    int k = k_min;
    for (int t = 0; t < n_thr; ++t) k += (area >= thr[t]) ? 1 : 0;
    return k;
}

constexpr int kPlanWarps = 4;               // RoIs per plan CTA: one warp each
constexpr int kPlanThreads = kPlanWarps * 32;
constexpr int kPlanMaxKeys = 256;
constexpr int kKeyBlock = 1024;             // RoIs per block of rpool_keys_kernel
constexpr int kPlanSingle = 1024;            // up to this many RoIs every plan warp derives all keys itself

// Level (clipped to the pyramid, maskrcnn.py:141), schedule key and flags of RoI i.
__device__ __forceinline__ int plan_key(const PlanParams &p, int i, int &lvl_out, int &flags_out)
{
    const RoiBox q = load_roi(p.rois, i, p.roi_format);
    const int L = p.n_levels;
    int lvl;
    if (p.given_levels) lvl = __ldg(p.given_levels + i);
    else if (p.given_levels_f32) lvl = (int)__ldg(p.given_levels_f32 + i);  // astype(int32)
    else lvl = level_from_area(q.y1, q.x1, q.y2, q.x2, p.thr, p.n_thr, p.k_min);
    int flags = 0;
    if ((p.given_levels || p.given_levels_f32) && (lvl < 0 || lvl >= L)) flags |= RPOOL_FLAG_LEVEL_CLIPPED;
    lvl = lvl < 0 ? 0 : (lvl >= L ? L - 1 : lvl);
    if (q.b < 0 || q.b >= p.n_images) flags |= RPOOL_FLAG_BAD_BATCH;
    const int b = q.b < 0 ? 0 : (q.b >= p.n_images ? p.n_images - 1 : q.b);
    const int lk = (p.order_mode >= RPOOL_SCHED_LEVEL_DESC) ? (L - 1 - lvl) : lvl;
    lvl_out = lvl;
    flags_out = flags;
    // COARSE_FIRST: coarse levels (the widest windows, the longest CTAs) of every image first
    return !p.by_image ? lk : (p.order_mode == RPOOL_SCHED_COARSE_FIRST ? lk * p.n_images + b : b * L + lk);
}

__global__ void __launch_bounds__(kKeyBlock)
rpool_keys_kernel(const __grid_constant__ PlanParams p)
{
    allow_dependents_early();     // the plan kernel builds its tables meanwhile and waits before ranking
    __shared__ int hist[kPlanMaxKeys];
    const int tid = threadIdx.x;
    for (int k = tid; k < p.K; k += kKeyBlock) hist[k] = 0;
    __syncthreads();
    const int i = blockIdx.x * kKeyBlock + tid;
    int key = -1 - (tid & 31);
    if (i < p.R) {
        int lvl, fl;
        key = plan_key(p, i, lvl, fl);
        p.keys[i] = key;
    }
    // one shared-memory atomic per distinct key and warp
    const unsigned same = __match_any_sync(0xffffffffu, key);
    if (i < p.R && (same & ((1u << (tid & 31)) - 1u)) == 0) atomicAdd(&hist[key], __popc(same));
    __syncthreads();
    for (int k = tid; k < p.K; k += kKeyBlock) p.bh[blockIdx.x * p.K + k] = hist[k];
}

// Footprint tables of one RoI by one warp: entries (pooled size, axis, bin) are dealt to
// the lanes, the window extent and the widest footprints are warp reductions.
__device__ __forceinline__ void warp_build_tables(const KParams &P, const RoiCtx &c, bool bwd, BlockCtl *ctl,
                                                  int lane)
{
    int total = 0;
    for (int h = 0; h < P.n_heads; ++h) total += P.PH[h] + P.PW[h];
    int wmin[2] = {0x7fffffff, 0x7fffffff}, wmax[2] = {-1, -1};
    int nmax[kMaxHeads][2];
    for (int h = 0; h < kMaxHeads; ++h) nmax[h][0] = nmax[h][1] = 0;
    bool fits = true;
    for (int e0 = lane; e0 < total; e0 += 32) {
        int h = 0, e = e0;
        while (e >= P.PH[h] + P.PW[h]) { e -= P.PH[h] + P.PW[h]; ++h; }
        const int ny = P.PH[h];
        const int axis = e < ny ? 0 : 1;
        const int pbin = axis ? e - ny : e;
        int lo, hi;
        fits = fill_axis_entry(ctl->hd[h].tab[axis], axis_of(P, c, bwd, h, axis), P.mode, pbin, lo, hi) && fits;
        if (hi >= lo) {
            wmin[axis] = lo < wmin[axis] ? lo : wmin[axis];
            wmax[axis] = hi > wmax[axis] ? hi : wmax[axis];
#pragma unroll
            for (int hh = 0; hh < kMaxHeads; ++hh)
                if (hh == h) nmax[hh][axis] = hi - lo + 1 > nmax[hh][axis] ? hi - lo + 1 : nmax[hh][axis];
        }
    }
    fits = __all_sync(0xffffffffu, fits);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        wmin[a] = __reduce_min_sync(0xffffffffu, wmin[a]);
        wmax[a] = __reduce_max_sync(0xffffffffu, wmax[a]);
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) nmax[h][a] = __reduce_max_sync(0xffffffffu, nmax[h][a]);
    }
    if (lane == 0) {
        ctl->wmin[0] = wmin[0]; ctl->wmin[1] = wmin[1];
        ctl->wmax[0] = wmax[0]; ctl->wmax[1] = wmax[1];
        for (int h = 0; h < kMaxHeads; ++h) { ctl->hd[h].nmax[0] = nmax[h][0]; ctl->hd[h].nmax[1] = nmax[h][1]; }
        if (!fits) ctl->flags &= ~kRecFits;
    }
    __syncwarp();
}

// Cuts every pooled size's bins into chunks whose x footprints fit a span of kSW columns
// (as build_chunks), one lane per bin: the end of a chunk is the first bin that no longer
// fits, found with one ballot; its per-offset bin counts and column mask are reductions.
__device__ __forceinline__ void warp_build_chunks(const KParams &P, const RoiCtx &c, BlockCtl *ctl, int max_bins,
                                                  int lane)
{
    for (int h = 0; h < P.n_heads; ++h) {
        constexpr int kSpan = kSW;
        unsigned char *cstart = ctl->hd[h].cstart;
        int *cx0 = ctl->hd[h].cx0;
        unsigned long long *ccnt = ctl->hd[h].ccnt;
        const AxisTab &xt = ctl->hd[h].tab[1];
        const int PW = P.PW[h];
        const int W = c.L.W;
        int NX = ctl->hd[h].nmax[1];
        NX = NX < 1 ? 1 : NX;
        const int lo_l = lane < PW ? xt.lo[lane] : 0;
        const int n_l = lane < PW ? xt.n[lane] : 0;
        int n = 0, pa = 0;
        while (pa < PW) {
            const int lo_a = __shfl_sync(0xffffffffu, lo_l, pa);
            int x0 = lo_a < W - kSpan ? lo_a : W - kSpan;   // span [x0, x0 + kSpan) inside the image
            x0 = x0 < 0 ? 0 : x0;
            const int last = lo_l + NX - 1;
            const int lim = last < W - 1 ? last : W - 1;   // taps beyond the image carry no weight
            const bool out = lane > pa && lane < PW && (lim - x0 >= kSpan || lane - pa >= max_bins);
            const unsigned mo = __ballot_sync(0xffffffffu, out);
            const int pe = mo ? __ffs(mo) - 1 : PW;          // first bin of the next chunk
            const bool in = lane >= pa && lane < pe;
            // per-offset counts (one byte each): sum of 1 << 8 * (lo - x0) over the chunk's bins
            const int sh8 = in ? 8 * (lo_l - x0) : 0;
            unsigned c_lo = (in && sh8 < 32) ? (1u << sh8) : 0u;
            unsigned c_hi = (in && sh8 >= 32) ? (1u << (sh8 - 32)) : 0u;
            unsigned msk = in ? (((1u << n_l) - 1u) << (lo_l - x0)) : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                c_lo += __shfl_xor_sync(0xffffffffu, c_lo, d);
                c_hi += __shfl_xor_sync(0xffffffffu, c_hi, d);
            }
            msk = __reduce_or_sync(0xffffffffu, msk);
            if (lane == 0) {
                cstart[n] = (unsigned char)pa;
                cx0[n] = x0;
                ccnt[n] = ((unsigned long long)c_hi << 32) | c_lo;
                ctl->hd[h].cmask[n] = (unsigned char)(msk & 0xffu);
            }
            ++n;
            pa = pe;
        }
        if (lane == 0) {
            cstart[n] = (unsigned char)PW;
            ctl->hd[h].nchunk = n;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kPlanThreads)
rpool_plan_kernel(const __grid_constant__ KParams P, const __grid_constant__ PlanParams p,
                  unsigned char *recs_fwd, unsigned char *recs_bwd)
{
    __shared__ __align__(16) BlockCtl ctl_s[kPlanWarps];
    __shared__ int s_hist[kPlanMaxKeys + 1];
    __shared__ unsigned char s_keys[kPlanSingle];  // single-launch mode: every RoI's key (K <= 256)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kPlanWarps + warp;
    const bool first = (blockIdx.x == 0);         // this CTA also writes gstart

    // ---- single-launch mode: the CTA derives all keys once, its warps rank from shared memory
    if (p.n_blocks == 0 && p.order_mode != RPOOL_SCHED_INPUT) {
        for (int j = tid; j < p.R; j += kPlanThreads) {
            int l2, f2;
            s_keys[j] = (unsigned char)plan_key(p, j, l2, f2);
        }
        __syncthreads();
    }

    // ---- the RoI's record for the forward geometry, built before anything of rpool_keys_kernel is
    // read: in two-launch mode this kernel starts in that kernel's shadow
    const bool mine = i < p.R;
    int lvl = 0, fl = 0, key = 0;
    BlockCtl *ctl = &ctl_s[warp];
    RoiCtx c;
    const int n16 = rec_bytes(P.n_heads) >> 4;
    const int n_sets = recs_bwd != recs_fwd ? 2 : 1;
    auto build = [&](int bwd) {
        __syncwarp();    // the previous set has left shared memory
        if (lane == 0) {
            ctl->r = c.r; ctl->lvl = c.lvl; ctl->b = c.b;
            ctl->flags = (c.valid ? kRecValid : 0) | (c.fast_ok ? kRecShape : 0) | kRecFits;
        }
        __syncwarp();
        if (c.valid && c.fast_ok) {      // (uniform) otherwise the consumer takes the generic path
            warp_build_tables(P, c, bwd != 0, ctl, lane);
            if (ctl->flags & kRecFits) warp_build_chunks(P, c, ctl, kPMax, lane);
        }
        __syncwarp();
    };
    if (mine) {
        key = plan_key(p, i, lvl, fl);
        c.r = i;
        c.lvl = lvl;
        c.L = P.lvl[lvl];
        c.box = load_roi(P.rois, i, P.roi_format);
        c.b = c.box.b;
        c.valid = (c.b >= 0 && c.b < c.L.n_images);
        c.fast_ok = shapes_allow_tables(P, c.L);
        build(0);
    }
    wait_for_predecessors();      // rpool_keys_kernel's keys and histograms from here on

    // ---- gstart: first slot of every key (CTA 0, all its warps)
    if (first && p.gstart) {
        for (int k = tid; k <= kPlanMaxKeys; k += kPlanThreads) s_hist[k] = 0;
        __syncthreads();
        if (p.order_mode != RPOOL_SCHED_INPUT) {
            if (p.n_blocks == 0) {
                for (int j = tid; j < p.R; j += kPlanThreads) atomicAdd(&s_hist[s_keys[j]], 1);
            } else {
                const int n = p.n_blocks * p.K;
                for (int e = tid; e < n; e += kPlanThreads) atomicAdd(&s_hist[e % p.K], __ldg(p.bh + e));
            }
        }
        __syncthreads();
        if (tid == 0) {
            const bool groups = p.by_image && p.order_mode == RPOOL_SCHED_DEFAULT;
            int run = 0;
            for (int k = 0; k <= kPlanMaxKeys; ++k) {
                const int t = s_hist[k];
                p.gstart[k] = groups ? (k < p.K ? run : p.R) : -1;
                run += t;
            }
            *p.det_err = 0;
        }
    }
    if (!mine) return;

    // ---- slot of RoI i: stable rank of (key_i, i)
    int slot = i;
    if (p.order_mode != RPOOL_SCHED_INPUT) {
        int acc = 0;
        if (p.n_blocks == 0) {
            for (int j = lane; j < p.R; j += 32) {
                const int kj = s_keys[j];
                acc += (kj < key || (kj == key && j < i)) ? 1 : 0;
            }
        } else {
            const int blk = i / kKeyBlock;
            const int n = p.n_blocks * p.K;
            for (int e = lane; e < n; e += 32) {
                const int b = e / p.K, k = e - b * p.K;
                acc += (k < key || (k == key && b < blk)) ? __ldg(p.bh + e) : 0;
            }
            for (int j = blk * kKeyBlock + lane; j < i; j += 32)
                acc += (__ldg(p.keys + j) == key) ? 1 : 0;
        }
        slot = __reduce_add_sync(0xffffffffu, acc);
    }
    if (lane == 0) {
        p.levels[i] = lvl;
        p.order[slot] = i;
        p.rflags[i] = fl;
    }
    for (int bwd = 0; bwd < n_sets; ++bwd) {
        if (bwd) build(bwd);
        const uint4 *src = reinterpret_cast<const uint4 *>(ctl);
        uint4 *dst = reinterpret_cast<uint4 *>((bwd ? recs_bwd : recs_fwd) + (size_t)slot * P.rec_stride);
        for (int k = lane; k < n16; k += 32) dst[k] = src[k];
    }
}

// OR of the per-RoI plan flags and the deterministic pass's flag (rpool_status_flags)
__global__ void rpool_flags_kernel(const int *__restrict__ rflags, int n, const int *det_err, int *out)
{
    __shared__ int acc;
    if (threadIdx.x == 0) acc = 0;
    __syncthreads();
    int f = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) f |= rflags[i];
    if (f) atomicOr(&acc, f);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int e = *det_err;
        *out = acc | (e == 1 ? RPOOL_FLAG_DET_GENERIC : 0) | (e == 2 ? RPOOL_FLAG_DET_SCRATCH : 0);
    }
}

// standalone level mapper (map_rois_to_fpn_levels on the device)
struct LevelParams {
    const float *boxes;
    int n, stride, roi_format;
    float thr[kMaxLevels];
    int n_thr, k_min, k_cap;
    float *out_f;
    int *out_i;
};
__global__ void rpool_levels_kernel(const __grid_constant__ LevelParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float *b = p.boxes + (size_t)i * p.stride + (p.stride - 4);
    float y1, x1, y2, x2;
    if (p.roi_format == RPOOL_ROI_YX) { y1 = b[0]; x1 = b[1]; y2 = b[2]; x2 = b[3]; }
    else { x1 = b[0]; y1 = b[1]; x2 = b[2]; y2 = b[3]; }
    int k = level_from_area(y1, x1, y2, x2, p.thr, p.n_thr, p.k_min);
    k = k > p.k_cap ? p.k_cap : k;
    if (p.out_f) p.out_f[i] = (float)k;
    if (p.out_i) p.out_i[i] = k;
}

// ---------------------------------------------------------------------------
// deterministic backward: segmented reduction instead of atomics
// ---------------------------------------------------------------------------
// 1. rpool_det_scan_kernel    window rectangle of every RoI (from its record's header) and an
//                             exclusive scan of the window sizes in schedule order -> offset
//                             of each RoI's private window in the scratch;
// 2. rpool_backward_kernel    (det = 1) writes each RoI's window contribution to its private
//                             window: one writer per cell, the first contribution a plain store;
// 3. rpool_det_gather_kernel  every feature cell sums, in schedule order, the windows that
//                             cover it (four windows' loads in flight at a time) and is written
//                             exactly once (this also replaces the zero fill).
// The summation order of every cell is fixed by the schedule, so results are
// bit-identical from run to run.
struct ScanParams {
    const unsigned char *recs;   // per-slot records (backward geometry): the window extent is in the header
    int rec_stride;
    int R, C;
    int shapes_ok;               // pooled sizes within the backward table path, table path not disabled
    int *rects;                  // out, per slot: x0, y0, x1, y1
    unsigned long long *woff;    // out, per slot: float offset of its private window
    unsigned long long *total;   // out: floats of scratch needed
    int *err;
};

constexpr int kScanThreads = 256;   // (a small CTA: it fits beside the pooling CTAs still draining when it
constexpr int kScanPer = 4;         //  starts in the previous launch's tail); slots per thread and round

__global__ void __launch_bounds__(kScanThreads)
rpool_det_scan_kernel(const __grid_constant__ ScanParams p)
{
    __shared__ unsigned long long wsum[kScanThreads / 32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    allow_dependents_early();     // the backward launch loads its records meanwhile and waits before it reads the offsets
    if (tid == 0) { carry = 0; *p.err = 0; }
    __syncthreads();
    for (int base = 0; base < p.R; base += kScanThreads * kScanPer) {
        unsigned long long v[kScanPer], mine = 0ull;
#pragma unroll
        for (int u = 0; u < kScanPer; ++u) {
            const int slot = base + tid * kScanPer + u;
            v[u] = 0ull;
            if (slot < p.R) {
                const int4 a = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride));
                const int4 f = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride) + 1);
                // a = wmin[0] (y0), wmin[1] (x0), wmax[0] (y1), wmax[1] (x1);  f = r, lvl, b, flags
                const bool valid = (f.w & kRecValid) != 0;
                const bool ok = p.shapes_ok && (f.w & (kRecShape | kRecFits)) == (kRecShape | kRecFits) &&
                                a.z - a.x < kExt;
                if (valid && !ok) atomicExch(p.err, 1);   // this RoI needs the generic path: not orderable
                int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
                if (valid && ok && a.w >= a.y && a.z >= a.x) {
                    x0 = a.y; y0 = a.x; x1 = a.w; y1 = a.z;
                    v[u] = (unsigned long long)(x1 - x0 + 1) * (y1 - y0 + 1) * p.C;
                }
                *reinterpret_cast<int4 *>(p.rects + 4 * (size_t)slot) = make_int4(x0, y0, x1, y1);
            }
            mine += v[u];
        }
        unsigned long long x = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        unsigned long long before = carry + x - mine;
        for (int w = 0; w < warp; ++w) before += wsum[w];
#pragma unroll
        for (int u = 0; u < kScanPer; ++u) {
            const int slot = base + tid * kScanPer + u;
            if (slot < p.R) p.woff[slot] = before;
            before += v[u];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry = before;
        __syncthreads();
    }
    if (tid == 0) *p.total = carry;
    // (launched in the previous kernel's tail: this grid must not complete before that kernel has,
    // because the launch behind it waits for THIS grid only)
    wait_for_predecessors();
}

struct GatherParams {
    LevelDev lvl[kMaxLevels];
    long long strip_base[kMaxLevels + 1];  // first CTA of every level, in launch order (coarse first)
    int strips[kMaxLevels];                // CTAs per map row
    int cells[kMaxLevels];                 // cells per CTA (a multiple of kGatherWarps)
    int n_levels, C, accumulate;
    const int *gstart;                     // first slot of every (image, level) group
    const int *rects;                      // per slot: x0, y0, x1, y1 of the RoI's window
    const unsigned long long *woff;        // per slot: float offset of its private window
    const float *scratch;
    unsigned long long scratch_floats;     // windows that end beyond it were never written (flagged by the
                                           // window pass): they are not listed, so nothing outside is read
};

constexpr int kGatherWarps = 8;       // cells in flight per CTA
constexpr int kGatherRowCtas = 8;     // at most this many CTAs per map row
constexpr int kGatherList = 1024;     // windows a CTA lists per pass
constexpr int kGatherSlabs = 4;       // 128-channel slabs held in registers: C <= 512
constexpr int kGatherU = 4;           // windows whose loads are in flight together (per warp); 6 or 8 cost
                                      // more in resident warps than they gain (profiles/r02_experiments.log)

// One CTA per strip of a map row (a few CTAs per row: wide strips on the fine levels, where most
// cells are covered by nothing and the CTA's fixed cost would otherwise dominate).  The CTA lists,
// in schedule order, the windows of its (image, level) group that meet the strip (one coalesced
// 16-byte load per slot, ordered compaction); its warps then take the strip's cells one by one.
template <int kSlabs>
__global__ void __launch_bounds__(kGatherWarps * 32)
rpool_det_gather_kernel(const __grid_constant__ GatherParams p)
{
    __shared__ int s_x0[kGatherList], s_x1[kGatherList], s_y0[kGatherList], s_wc[kGatherList];
    __shared__ unsigned long long s_off[kGatherList];
    __shared__ int s_wcount[kGatherWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // coarse levels come first in the launch: their strips meet the most windows (the longest CTAs)
    int li = 0;
    while (li + 1 < p.n_levels && (long long)blockIdx.x >= p.strip_base[li + 1]) ++li;
    const int l = p.n_levels - 1 - li;
    const LevelDev L = p.lvl[l];
    long long idx = (long long)blockIdx.x - p.strip_base[li];
    const int sx = (int)(idx % p.strips[l]);
    idx /= p.strips[l];
    const int y = (int)(idx % L.H);
    const int b = (int)(idx / L.H);
    const int xs = sx * p.cells[l];
    int xe = xs + p.cells[l] - 1;
    xe = xe < L.W - 1 ? xe : L.W - 1;
    const int key = b * p.n_levels + l;
    const int g0 = p.gstart[key], g1 = p.gstart[key + 1];
    const int C = p.C;

    int next = g0;
    bool more = true;
    for (int pass = 0; more; ++pass) {
        // ---- list the next windows (at most kGatherList) that meet this strip, in schedule order
        int n = 0;
        while (next < g1 && n + kGatherWarps * 32 <= kGatherList) {
            const int slot = next + tid;
            bool hit = false;
            int4 rc = make_int4(0, 0, -1, -1);
            if (slot < g1) {
                rc = __ldg(reinterpret_cast<const int4 *>(p.rects) + slot);
                hit = rc.y <= y && y <= rc.w && rc.x <= xe && rc.z >= xs;
            }
            unsigned long long woff = 0ull;
            if (hit) {
                woff = __ldg(p.woff + slot);
                hit = woff + (unsigned long long)(rc.z - rc.x + 1) * (rc.w - rc.y + 1) * C <= p.scratch_floats;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[warp] = __popc(m);
            __syncthreads();
            int pos = n + __popc(m & ((1u << lane) - 1u));
            int total = 0;
            for (int w = 0; w < kGatherWarps; ++w) {
                if (w < warp) pos += s_wcount[w];
                total += s_wcount[w];
            }
            if (hit) {
                s_x0[pos] = rc.x; s_x1[pos] = rc.z; s_y0[pos] = rc.y; s_wc[pos] = rc.z - rc.x + 1;
                s_off[pos] = woff;
            }
            n += total;
            next += kGatherWarps * 32;
            __syncthreads();
        }
        more = next < g1;
        // (launched in the backward launch's tail: the private windows -- and the gradient map, when
        // it is accumulated into -- from here on)
        if (pass == 0) wait_for_predecessors();
        // ---- the strip's cells
        for (int x = xs + warp; x <= xe; x += kGatherWarps) {
            float *dst = L.data + (((size_t)b * L.H + y) * L.W + x) * C + lane * 4;
            float4 acc[kSlabs];
#pragma unroll
            for (int k = 0; k < kSlabs; ++k) {
                acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((pass > 0 || p.accumulate) && k * 128 + lane * 4 < C)
                    acc[k] = *reinterpret_cast<const float4 *>(dst + k * 128);
            }
            // 32 list entries at a time: a lane tests one and works out where this cell lies in that
            // window; the hits are then taken four at a time -- all their loads in flight together --
            // and added in list (= schedule) order
            for (int e0 = 0; e0 < n; e0 += 32) {
                const int e = e0 + lane;
                const bool cov = e < n && s_x0[e] <= x && x <= s_x1[e];
                unsigned long long mine = 0ull;
                if (cov) mine = s_off[e] + ((unsigned long long)(y - s_y0[e]) * s_wc[e] + (x - s_x0[e])) * C;
                unsigned hits = __ballot_sync(0xffffffffu, cov);
                while (hits) {
                    float4 v[kGatherU][kSlabs];
#pragma unroll
                    for (int u = 0; u < kGatherU; ++u) {
                        const int src_lane = hits ? __ffs(hits) - 1 : 0;
                        const bool on = hits != 0;
                        hits &= hits - 1;
                        const unsigned long long off = __shfl_sync(0xffffffffu, mine, src_lane);
                        const float *src = p.scratch + off + lane * 4;
#pragma unroll
                        for (int k = 0; k < kSlabs; ++k) {
                            v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (on && k * 128 + lane * 4 < C) v[u][k] = ldg_nc128(src + k * 128);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kGatherU; ++u)
#pragma unroll
                        for (int k = 0; k < kSlabs; ++k) {
                            acc[k].x += v[u][k].x; acc[k].y += v[u][k].y; acc[k].z += v[u][k].z; acc[k].w += v[u][k].w;
                        }
                }
            }
#pragma unroll
            for (int k = 0; k < kSlabs; ++k)
                if (k * 128 + lane * 4 < C) stg128(dst + k * 128, acc[k]);
        }
        // the list is consumed before the next pass overwrites it.  (No barrier after the last pass: a warp
        // that has finished its cells leaves at once -- with the barrier the gather ran 159 us instead of 120.)
        if (more) __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// zero fill of all gradient levels in one launch
// ---------------------------------------------------------------------------
struct ZeroParams {
    float *ptr[kMaxLevels];
    unsigned long long n4[kMaxLevels];    // float4 count
    unsigned long long tail[kMaxLevels];  // remaining floats
    int n;
};
__global__ void rpool_zero_kernel(const __grid_constant__ ZeroParams p)
{
    allow_dependents_early();     // the backward launch queued behind the fill loads its records and
                                  // gy rows meanwhile and waits before its first reduction
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int l = 0; l < p.n; ++l) {
        float4 *d = reinterpret_cast<float4 *>(p.ptr[l]);
        for (unsigned long long i = t0; i < p.n4[l]; i += stride) d[i] = z;
        if (t0 < p.tail[l]) p.ptr[l][p.n4[l] * 4 + t0] = 0.f;
    }
}

// ---------------------------------------------------------------------------
// NCHW <-> NHWC (per image: C x HW  <->  HW x C), 32x32 tiles through smem
// ---------------------------------------------------------------------------
__global__ void rpool_transpose_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                       int rows, int cols)
{
    // src: (batch, rows, cols) -> dst: (batch, cols, rows)
    __shared__ float tile[32][33];
    const size_t img = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, cc = c0 + threadIdx.x;
        if (r < rows && cc < cols) tile[j][threadIdx.x] = src[img + (size_t)r * cols + cc];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int cc = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && cc < cols) dst[img + (size_t)cc * rows + r] = tile[threadIdx.x][j];
    }
}

}  // namespace rpool
