// rpool_kernels.cuh -- the fused multi-level RoIAlign kernels (sm_100a).
//
// One CTA per RoI, scheduled through the plan's (image, level) binning.  A CTA
//   1. decodes its RoI and level, builds per-axis footprint tables in shared
//      memory (merged bilinear taps per bin, separable in y and x);
//   2. picks a path:
//        staged  -- the RoI's feature window (all channels) is copied into
//                   shared memory with cp.async.bulk (one bulk copy per window
//                   row, completion on an mbarrier), in bands of bin rows when
//                   the whole window does not fit;
//        direct  -- same arithmetic straight from global memory (window wider
//                   than the staging buffer);
//        generic -- any layout / sampling grid / pooled size, tap by tap, in
//                   the reference's own operation order;
//   3. forward: every warp takes (head, bin row, 128-channel slab) tasks; a
//      lane owns 4 channels (128-bit loads), interpolates along y once per
//      window column, slides a kNT-column register window along x and streams
//      the bins out with evict-first 128-bit stores;
//      backward: every warp owns a 32-channel slab of the CTA's private
//      gradient window in shared memory (no shared-memory atomics), walks all
//      bins of all heads, and the CTA flushes the window once with 128-bit
//      vector reductions (red.global.add.v4.f32) into the dense gradient.
#pragma once
#include "rpool_device.cuh"

namespace rpool {

// ---------------------------------------------------------------------------
// shared prologue: RoI decode + tables
// ---------------------------------------------------------------------------
struct RoiCtx {
    int r, lvl, b;
    bool valid;       // batch index inside the level's tensor
    bool fast_ok;     // layouts/alignment allow the table-driven paths at all
    LevelDev L;
    AxisGeom gy[kMaxHeads], gx[kMaxHeads];
};

__device__ __forceinline__ void roi_prologue(const KParams &P, bool bwd, RoiCtx &c)
{
    c.r = P.order[blockIdx.x];
    int lvl = P.roi_level[c.r];
    lvl = lvl < 0 ? 0 : (lvl >= P.n_levels ? P.n_levels - 1 : lvl);
    c.lvl = lvl;
    c.L = P.lvl[lvl];
    const RoiBox q = load_roi(P.rois, c.r, P.roi_format);
    c.b = q.b;
    c.valid = (q.b >= 0 && q.b < c.L.n_images);
    for (int h = 0; h < P.n_heads; ++h) {
        c.gy[h] = make_axis(P.mode, bwd, q.y1, q.y2, c.L.scale, P.PH[h], P.S, c.L.H);
        c.gx[h] = make_axis(P.mode, bwd, q.x1, q.x2, c.L.scale, P.PW[h], P.S, c.L.W);
    }
    bool ok = (P.feat_layout == RPOOL_NHWC) && (P.pool_layout == RPOOL_NHWC) && (P.C % 4 == 0);
    ok = ok && ((reinterpret_cast<uintptr_t>(c.L.data) & 15) == 0);
    for (int h = 0; h < P.n_heads; ++h) {
        ok = ok && P.PH[h] <= kPMax && P.PW[h] <= kPMax;
        ok = ok && ((reinterpret_cast<uintptr_t>(P.pooled[h]) & 15) == 0);
    }
    c.fast_ok = ok && (P.force_path != kPathGeneric);
}

// Builds the tables; returns with ctl fully populated and the CTA synchronised.
__device__ __forceinline__ void build_tables(const KParams &P, const RoiCtx &c, BlockCtl *ctl)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        ctl->wmin[0] = ctl->wmin[1] = 0x7fffffff;
        ctl->wmax[0] = ctl->wmax[1] = -1;
        ctl->eligible = 1;
    }
    __syncthreads();
    // entry e -> (head, axis, bin)
    int base = 0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int ny = P.PH[h], nx = P.PW[h];
        const int e = tid - base;
        if (e >= 0 && e < ny + nx) {
            const int axis = e < ny ? 0 : 1;
            const int p = axis ? e - ny : e;
            int lo, hi;
            const bool ok = fill_axis_entry(ctl->tab[h][axis], axis ? c.gx[h] : c.gy[h], P.mode,
                                            p, lo, hi);
            if (!ok) ctl->eligible = 0;
            if (hi >= lo) {
                atomicMin(&ctl->wmin[axis], lo);
                atomicMax(&ctl->wmax[axis], hi);
            }
        }
        base += ny + nx;
    }
    __syncthreads();
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

__device__ void generic_forward(const KParams &P, const RoiCtx &c)
{
    const int C = P.C;
    const Strides4 fs = strides_of(P.feat_layout, C, c.L.H, c.L.W);
    const float *feat = c.L.data + (long long)(c.valid ? c.b : 0) * fs.s0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int PH = P.PH[h], PW = P.PW[h];
        const Strides4 os = strides_of(P.pool_layout, C, PH, PW);
        float *out = P.pooled[h] + (long long)c.r * os.s0;
        const AxisGeom gy = c.gy[h], gx = c.gx[h];
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

__device__ void generic_backward(const KParams &P, const RoiCtx &c)
{
    if (!c.valid) return;
    const int C = P.C;
    const Strides4 fs = strides_of(P.feat_layout, C, c.L.H, c.L.W);
    float *grad = c.L.data + (long long)c.b * fs.s0;
    for (int h = 0; h < P.n_heads; ++h) {
        const int PH = P.PH[h], PW = P.PW[h];
        const Strides4 os = strides_of(P.pool_layout, C, PH, PW);
        const float *gyp = P.pooled[h] + (long long)c.r * os.s0;
        const AxisGeom gy = c.gy[h], gx = c.gx[h];
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
// window bookkeeping shared by forward and backward
// ---------------------------------------------------------------------------
struct Window {
    int y0, y1, x0, x1;  // inclusive extents of the rows/cols held
    int row_stride;      // floats between consecutive window rows
};

// Largest p1 such that bin rows [p0, p1) of table t stay within `cap` window rows.
__device__ __forceinline__ int band_end(const AxisTab &t, int P, int p0, int cap, int &ya, int &yb)
{
    int a = 0x7fffffff, b = -1, p = p0;
    for (; p < P; ++p) {
        const int n = t.n[p];
        if (n == 0) continue;
        const int lo = t.lo[p], hi = lo + n - 1;
        const int na = lo < a ? lo : a, nb = hi > b ? hi : b;
        if (nb - na + 1 > cap && b >= 0) break;
        a = na;
        b = nb;
    }
    ya = a;
    yb = b;
    return p;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <bool kSmem>
__device__ __forceinline__ float4 win_ld(uint32_t win_s, const float *win_g, int off)
{
    if (kSmem) return lds128(win_s + (uint32_t)off * 4u);
    return ldg_nc128(win_g + off);
}

// One (bin row, 4 channels per lane) task.  `base` is the float offset of
// (row ylo, col win.x0, channel c) from the window origin.
template <bool kSmem>
__device__ __forceinline__ void fwd_row(uint32_t win_s, const float *win_g, const Window &win,
                                        int C, int c, int ylo, int ny, float4 wy,
                                        const AxisTab &xt, int PW, float *__restrict__ out)
{
    const int base = (ylo - win.y0) * win.row_stride + c;
    auto col = [&](int x) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= win.x0 && x <= win.x1) {
            const int off = base + (x - win.x0) * C;
            float4 t = win_ld<kSmem>(win_s, win_g, off);
            v.x = wy.x * t.x; v.y = wy.x * t.y; v.z = wy.x * t.z; v.w = wy.x * t.w;
            if (ny > 1) {
                t = win_ld<kSmem>(win_s, win_g, off + win.row_stride);
                v.x = fmaf(wy.y, t.x, v.x); v.y = fmaf(wy.y, t.y, v.y);
                v.z = fmaf(wy.y, t.z, v.z); v.w = fmaf(wy.y, t.w, v.w);
            }
            if (ny > 2) {
                t = win_ld<kSmem>(win_s, win_g, off + 2 * win.row_stride);
                v.x = fmaf(wy.z, t.x, v.x); v.y = fmaf(wy.z, t.y, v.y);
                v.z = fmaf(wy.z, t.z, v.z); v.w = fmaf(wy.z, t.w, v.w);
            }
            if (ny > 3) {
                t = win_ld<kSmem>(win_s, win_g, off + 3 * win.row_stride);
                v.x = fmaf(wy.w, t.x, v.x); v.y = fmaf(wy.w, t.y, v.y);
                v.z = fmaf(wy.w, t.z, v.z); v.w = fmaf(wy.w, t.w, v.w);
            }
        }
        return v;
    };
    int xb = xt.lo[0];
    float4 V0 = col(xb), V1 = col(xb + 1), V2 = col(xb + 2), V3 = col(xb + 3);
    for (int pw = 0; pw < PW; ++pw) {
        const int xl = xt.lo[pw];
        if (xl - xb >= kNT) {
            xb = xl;
            V0 = col(xb); V1 = col(xb + 1); V2 = col(xb + 2); V3 = col(xb + 3);
        } else {
            while (xb < xl) {
                V0 = V1; V1 = V2; V2 = V3;
                V3 = col(xb + kNT);
                ++xb;
            }
        }
        const float4 w = xt.w[pw];
        float4 o;
        o.x = fmaf(w.w, V3.x, fmaf(w.z, V2.x, fmaf(w.y, V1.x, w.x * V0.x)));
        o.y = fmaf(w.w, V3.y, fmaf(w.z, V2.y, fmaf(w.y, V1.y, w.x * V0.y)));
        o.z = fmaf(w.w, V3.z, fmaf(w.z, V2.z, fmaf(w.y, V1.z, w.x * V0.z)));
        o.w = fmaf(w.w, V3.w, fmaf(w.z, V2.w, fmaf(w.y, V1.w, w.x * V0.w)));
        stg_stream128(out + (size_t)pw * C, o);
    }
}

// All (bin row in [p0,p1), slab) tasks of head h against the current window.
template <bool kSmem>
__device__ __forceinline__ void fwd_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl,
                                          uint32_t win_s, const float *win_g, const Window &win,
                                          int h, int p0, int p1)
{
    const int C = P.C;
    const int slabs = (C + 127) >> 7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int PH = P.PH[h], PW = P.PW[h];
    const AxisTab &yt = ctl->tab[h][0];
    const AxisTab &xt = ctl->tab[h][1];
    const int ntask = (p1 - p0) * slabs;
    for (int t = warp; t < ntask; t += nwarps) {
        const int ph = p0 + t / slabs;
        const int ch = (t % slabs) * 128 + lane * 4;
        if (ch >= C) continue;
        float *out = P.pooled[h] + (((size_t)c.r * PH + ph) * PW) * C + ch;
        const int ny = yt.n[ph];
        if (ny == 0) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int pw = 0; pw < PW; ++pw) stg_stream128(out + (size_t)pw * C, z);
            continue;
        }
        fwd_row<kSmem>(win_s, win_g, win, C, ch, yt.lo[ph], ny, yt.w[ph], xt, PW, out);
    }
}

// Copies window rows [ya, yb] x cols [x0, x1] (all channels) of image b into
// shared memory: one bulk async copy per row, completion counted on the mbarrier.
__device__ __forceinline__ void stage_window(const RoiCtx &c, int C, BlockCtl *ctl, uint32_t win_s,
                                             int ya, int yb, int x0, int x1, unsigned &phase)
{
    const int nrows = yb - ya + 1;
    const unsigned row_bytes = (unsigned)(x1 - x0 + 1) * C * 4u;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) mbar_expect_tx(&ctl->mbar, row_bytes * nrows);
        __syncwarp();
        for (int i = threadIdx.x; i < nrows; i += 32) {
            const float *src = c.L.data + (((size_t)c.b * c.L.H + (ya + i)) * c.L.W + x0) * C;
            bulk_g2s(win_s + (uint32_t)i * row_bytes, src, row_bytes, &ctl->mbar);
        }
    }
    mbar_wait(&ctl->mbar, phase);
    phase ^= 1u;
}

__global__ void __launch_bounds__(512)
rpool_forward_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);
    constexpr int kCtlBytes = (sizeof(BlockCtl) + 127) & ~127;
    const uint32_t win_s = smem_u32(smem_raw + kCtlBytes);

    RoiCtx c;
    roi_prologue(P, false, c);
    if (!c.fast_ok || !c.valid) {
        generic_forward(P, c);
        return;
    }
    if (threadIdx.x == 0) {
        mbar_init(&ctl->mbar, 1);
        fence_mbar_init();
    }
    build_tables(P, c, ctl);
    if (!ctl->eligible) {
        generic_forward(P, c);
        return;
    }
    const int C = P.C;
    const int x0 = ctl->wmin[1], x1 = ctl->wmax[1];
    const int y0 = ctl->wmin[0], y1 = ctl->wmax[0];
    if (x1 < x0 || y1 < y0) {
        // no valid sample anywhere: all outputs are zero
        for (int h = 0; h < P.n_heads; ++h) {
            const int total = P.PH[h] * P.PW[h] * C;
            float *out = P.pooled[h] + (size_t)c.r * total;
            for (int i = threadIdx.x; i < total; i += blockDim.x) out[i] = 0.f;
        }
        return;
    }
    const int Wc = x1 - x0 + 1;
    const int row_floats = Wc * C;
    const int cap_rows = P.win_floats / row_floats;
    int path = P.force_path;
    if (path == kPathAuto) path = (cap_rows >= kNT) ? kPathStaged : kPathDirect;
    if (path == kPathStaged && cap_rows < kNT) path = kPathDirect;

    if (path == kPathDirect) {
        Window w{0, c.L.H - 1, 0, c.L.W - 1, c.L.W * C};
        const float *img = c.L.data + (size_t)c.b * c.L.H * c.L.W * C;
        for (int h = 0; h < P.n_heads; ++h)
            fwd_tasks<false>(P, c, ctl, 0, img, w, h, 0, P.PH[h]);
        return;
    }

    unsigned phase = 0;
    if (y1 - y0 + 1 <= cap_rows) {
        stage_window(c, C, ctl, win_s, y0, y1, x0, x1, phase);
        Window w{y0, y1, x0, x1, row_floats};
        for (int h = 0; h < P.n_heads; ++h)
            fwd_tasks<true>(P, c, ctl, win_s, nullptr, w, h, 0, P.PH[h]);
        return;
    }
    for (int h = 0; h < P.n_heads; ++h) {
        int p0 = 0;
        while (p0 < P.PH[h]) {
            int ya, yb;
            const int p1 = band_end(ctl->tab[h][0], P.PH[h], p0, cap_rows, ya, yb);
            if (yb >= ya) {
                stage_window(c, C, ctl, win_s, ya, yb, x0, x1, phase);
                Window w{ya, yb, x0, x1, row_floats};
                fwd_tasks<true>(P, c, ctl, win_s, nullptr, w, h, p0, p1);
                __syncthreads();  // all reads of this band done before it is overwritten
            } else {
                Window w{0, -1, x0, x1, row_floats};
                fwd_tasks<true>(P, c, ctl, win_s, nullptr, w, h, p0, p1);
            }
            p0 = p1;
        }
    }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
// Adds a*wy[j] to window cells (ylo+j, x), j < ny, for this lane's channel.
template <bool kSmem>
__device__ __forceinline__ void bwd_retire(uint32_t win_s, float *win_g, const Window &win, int C,
                                           int base, int ny, float4 wy, float a, int x)
{
    if (x < win.x0 || x > win.x1) return;
    const int off = base + (x - win.x0) * C;
    const float wv[4] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
    for (int j = 0; j < kNT; ++j) {
        if (j < ny) {
            const int o = off + j * win.row_stride;
            if (kSmem) {
                const uint32_t a32 = win_s + (uint32_t)o * 4u;
                sts32(a32, fmaf(wv[j], a, lds32(a32)));
            } else {
                red_add_f32(win_g + o, wv[j] * a);
            }
        }
    }
}

// One bin row for one channel per lane: gy values are pushed through the x
// footprints into a sliding set of kNT column accumulators; a column that
// slides out is multiplied by the row's y weights and added to the window.
template <bool kSmem>
__device__ __forceinline__ void bwd_row(uint32_t win_s, float *win_g, const Window &win, int C,
                                        int c, int ylo, int ny, float4 wy, const AxisTab &xt,
                                        int PW, const float *__restrict__ gy)
{
    const int base = (ylo - win.y0) * win.row_stride + c;
    float A0 = 0.f, A1 = 0.f, A2 = 0.f, A3 = 0.f;
    int xb = xt.lo[0];
    for (int pw0 = 0; pw0 < PW; pw0 += 8) {
        float g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            g[k] = (pw0 + k < PW) ? ldg_stream32(gy + (size_t)(pw0 + k) * C) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pw = pw0 + k;
            if (pw < PW) {
                const int xl = xt.lo[pw];
                if (xl - xb >= kNT) {
                    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A0, xb);
                    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A1, xb + 1);
                    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A2, xb + 2);
                    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A3, xb + 3);
                    A0 = A1 = A2 = A3 = 0.f;
                    xb = xl;
                } else {
                    while (xb < xl) {
                        bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A0, xb);
                        A0 = A1; A1 = A2; A2 = A3; A3 = 0.f;
                        ++xb;
                    }
                }
                const float4 w = xt.w[pw];
                A0 = fmaf(w.x, g[k], A0);
                A1 = fmaf(w.y, g[k], A1);
                A2 = fmaf(w.z, g[k], A2);
                A3 = fmaf(w.w, g[k], A3);
            }
        }
    }
    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A0, xb);
    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A1, xb + 1);
    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A2, xb + 2);
    bwd_retire<kSmem>(win_s, win_g, win, C, base, ny, wy, A3, xb + 3);
}

// Every warp owns 32-channel slabs; within a slab it walks rows [p0,p1) of head h.
template <bool kSmem>
__device__ __forceinline__ void bwd_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl,
                                          uint32_t win_s, float *win_g, const Window &win,
                                          int h, int p0, int p1)
{
    const int C = P.C;
    const int slabs = (C + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int PH = P.PH[h], PW = P.PW[h];
    const AxisTab &yt = ctl->tab[h][0];
    const AxisTab &xt = ctl->tab[h][1];
    for (int s = warp; s < slabs; s += nwarps) {
        const int ch = s * 32 + lane;
        if (ch >= C) continue;
        for (int ph = p0; ph < p1; ++ph) {
            const int ny = yt.n[ph];
            if (ny == 0) continue;
            const float *gy = P.pooled[h] + (((size_t)c.r * PH + ph) * PW) * C + ch;
            bwd_row<kSmem>(win_s, win_g, win, C, ch, yt.lo[ph], ny, yt.w[ph], xt, PW, gy);
        }
    }
}

__device__ __forceinline__ void zero_window(uint32_t win_s, int nfloats)
{
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x * 4; i < nfloats; i += blockDim.x * 4) sts128(win_s + (uint32_t)i * 4u, z);
}

// Adds the private window into the dense gradient with 128-bit reductions.
__device__ __forceinline__ void flush_window(const RoiCtx &c, int C, uint32_t win_s, int ya, int yb,
                                             int x0, int x1)
{
    const int row4 = (x1 - x0 + 1) * C / 4;
    const int total4 = (yb - ya + 1) * row4;
    for (int i = threadIdx.x; i < total4; i += blockDim.x) {
        const float4 v = lds128(win_s + (uint32_t)i * 16u);
        if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
        const int row = i / row4, rem = i - row * row4;
        float *dst = c.L.data + (((size_t)c.b * c.L.H + (ya + row)) * c.L.W + x0) * C + (size_t)rem * 4;
        red_add_v4(dst, v);
    }
}

__global__ void __launch_bounds__(512)
rpool_backward_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);
    constexpr int kCtlBytes = (sizeof(BlockCtl) + 127) & ~127;
    const uint32_t win_s = smem_u32(smem_raw + kCtlBytes);

    RoiCtx c;
    roi_prologue(P, true, c);
    if (!c.valid) return;
    if (!c.fast_ok) {
        generic_backward(P, c);
        return;
    }
    build_tables(P, c, ctl);
    if (!ctl->eligible) {
        generic_backward(P, c);
        return;
    }
    const int C = P.C;
    const int x0 = ctl->wmin[1], x1 = ctl->wmax[1];
    const int y0 = ctl->wmin[0], y1 = ctl->wmax[0];
    if (x1 < x0 || y1 < y0) return;
    const int Wc = x1 - x0 + 1;
    const int row_floats = Wc * C;
    const int cap_rows = P.win_floats / row_floats;
    int path = P.force_path;
    if (path == kPathAuto) path = (cap_rows >= kNT) ? kPathStaged : kPathDirect;
    if (path == kPathStaged && cap_rows < kNT) path = kPathDirect;

    if (path == kPathDirect) {
        Window w{0, c.L.H - 1, 0, c.L.W - 1, c.L.W * C};
        float *img = c.L.data + (size_t)c.b * c.L.H * c.L.W * C;
        for (int h = 0; h < P.n_heads; ++h)
            bwd_tasks<false>(P, c, ctl, 0, img, w, h, 0, P.PH[h]);
        return;
    }

    if (y1 - y0 + 1 <= cap_rows) {
        zero_window(win_s, (y1 - y0 + 1) * row_floats);
        __syncthreads();
        Window w{y0, y1, x0, x1, row_floats};
        for (int h = 0; h < P.n_heads; ++h)
            bwd_tasks<true>(P, c, ctl, win_s, nullptr, w, h, 0, P.PH[h]);
        __syncthreads();
        flush_window(c, C, win_s, y0, y1, x0, x1);
        return;
    }
    for (int h = 0; h < P.n_heads; ++h) {
        int p0 = 0;
        while (p0 < P.PH[h]) {
            int ya, yb;
            const int p1 = band_end(ctl->tab[h][0], P.PH[h], p0, cap_rows, ya, yb);
            if (yb >= ya) {
                zero_window(win_s, (yb - ya + 1) * row_floats);
                __syncthreads();
                Window w{ya, yb, x0, x1, row_floats};
                bwd_tasks<true>(P, c, ctl, win_s, nullptr, w, h, p0, p1);
                __syncthreads();
                flush_window(c, C, win_s, ya, yb, x0, x1);
                __syncthreads();
            }
            p0 = p1;
        }
    }
}

// ---------------------------------------------------------------------------
// plan: level assignment + stable (image, level) binning
// ---------------------------------------------------------------------------
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
    int order_mode; // 0 identity, 1 (image, level asc), 2 (image, level desc)
    int *levels;    // out
    int *order;     // out
    int *keys;      // scratch
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

constexpr int kPlanThreads = 1024;
constexpr int kPlanMaxKeys = 256;

__global__ void __launch_bounds__(kPlanThreads)
rpool_plan_kernel(const __grid_constant__ PlanParams p)
{
    __shared__ int hist[32 * kPlanMaxKeys];
    __shared__ int base[kPlanMaxKeys];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.n_levels;
    int K = p.n_images * L;
    const bool by_image = (K <= kPlanMaxKeys);
    if (!by_image) K = L;

    for (int i = tid; i < p.R; i += kPlanThreads) {
        const RoiBox q = load_roi(p.rois, i, p.roi_format);
        int lvl;
        if (p.given_levels) lvl = p.given_levels[i];
        else if (p.given_levels_f32) lvl = (int)p.given_levels_f32[i];  // astype(int32)
        else lvl = level_from_area(q.y1, q.x1, q.y2, q.x2, p.thr, p.n_thr, p.k_min);
        lvl = lvl < 0 ? 0 : (lvl >= L ? L - 1 : lvl);  // maskrcnn.py:141
        p.levels[i] = lvl;
        int b = q.b < 0 ? 0 : (q.b >= p.n_images ? p.n_images - 1 : q.b);
        const int lk = (p.order_mode == 2) ? (L - 1 - lvl) : lvl;
        p.keys[i] = by_image ? b * L + lk : lk;
    }
    if (p.order_mode == 0) {
        for (int i = tid; i < p.R; i += kPlanThreads) p.order[i] = i;
        return;
    }
    for (int i = tid; i < 32 * K; i += kPlanThreads) hist[i] = 0;
    __syncthreads();
    // warp w owns the contiguous chunk [w*chunk, (w+1)*chunk)
    const int chunk = ((p.R + 31) / 32 + 31) / 32 * 32;
    const int beg = warp * chunk;
    const int end = beg + chunk < p.R ? beg + chunk : p.R;
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&hist[warp * K + p.keys[i]], 1);
    __syncthreads();
    if (tid < K) {
        int run = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = hist[w * K + tid];
            hist[w * K + tid] = run;
            run += t;
        }
        base[tid] = run;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < K; ++k) {
            const int t = base[k];
            base[k] = run;
            run += t;
        }
    }
    __syncthreads();
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        const bool active = i < end;
        const int key = active ? p.keys[i] : -1 - lane;
        const unsigned mask = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        if (active) p.order[base[key] + hist[warp * K + key] + rank] = i;
        __syncwarp();
        if (active && rank == 0) hist[warp * K + key] += __popc(mask);
        __syncwarp();
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
