// rpool_slide.cuh -- "slide" forward kernel (sm_100a): window rows are loaded ONCE.
//
// The "rows" forward kernel gives every bin row its own warp task, so a window row is
// loaded again by each of the ~3.7 bin rows whose footprint contains it -- from L1, but
// the L1/LSU data path (128 B/clk per SM, shared by loads, stores and shared memory) is
// what bounds that kernel: 4 GB through L1 for 0.8 GB of output (profiles/r02c_rows_*).
// Here a warp task is one (pooled size, chunk of bins, 128-channel slab) and walks ALL bin
// rows of the RoI in order.  Footprints move monotonically down the window, so the two
// most recent window rows live in registers (A, B: kSL columns x 4 channels per lane
// each) and a bin row takes its footprint rows from them; only a row that has not been
// seen yet is loaded, and the row after it is requested into L1 at the same time.  Spans
// are kSL = 6 columns wide so that A, B and the y-interpolated row V (3 x 24 registers)
// fit the 128-register budget of 16 resident warps per SM.
//   column pass  V[s] = sum_j wy[j] * row(ylo + j)[cx0 + s],  s < kSL
//   bin pass     out[pw] = sum_k wx[pw][k] * V[lo[pw] - cx0 + k]: static sequence over the
//                kSL offsets, each running a table-supplied count of bins (as in "rows")
#pragma once
#include "rpool_kernels.cuh"

namespace rpool {

__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

template <int NX, int R>
__device__ __forceinline__ float4 taps6(const float4 (&V)[kSL], const float4 w)
{
    float4 o = mul4(w.x, V[R]);
    if (NX > 1 && R + 1 < kSL) fma4(o, w.y, V[R + 1]);
    if (NX > 2 && R + 2 < kSL) fma4(o, w.z, V[R + 2]);
    if (NX > 3 && R + 3 < kSL) fma4(o, w.w, V[R + 3]);
    return o;
}

template <int NX, int R>
__device__ __forceinline__ void fwd6_bins_at(const float4 (&V)[kSL], unsigned long long cnt,
                                             const float4 *&wp, float *&o_ptr, int C, bool active)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float4 o = taps6<NX, R>(V, *wp);
        if (active) stg_stream128(o_ptr, o);
        ++wp;
        o_ptr += C;
    }
}

template <int NX>
__device__ __forceinline__ void fwd6_bin_pass(const float4 (&V)[kSL], unsigned long long cnt,
                                              const float4 *wp, float *o_ptr, int C, bool active)
{
    fwd6_bins_at<NX, 0>(V, cnt, wp, o_ptr, C, active);
    fwd6_bins_at<NX, 1>(V, cnt, wp, o_ptr, C, active);
    fwd6_bins_at<NX, 2>(V, cnt, wp, o_ptr, C, active);
    fwd6_bins_at<NX, 3>(V, cnt, wp, o_ptr, C, active);
    fwd6_bins_at<NX, 4>(V, cnt, wp, o_ptr, C, active);
    fwd6_bins_at<NX, 5>(V, cnt, wp, o_ptr, C, active);
}

template <int kC>
__device__ __forceinline__ void fwd_slide_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl)
{
    const int C = kC ? kC : P.C;
    const int slabs = (C + 127) >> 7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int row_stride = c.L.W * C;
    const float *img = c.L.data + (size_t)c.b * c.L.H * row_stride;

    int ntask = 0;
    for (int h = 0; h < P.n_heads; ++h) ntask += ctl->hd[h].nchunk6 * slabs;
    for (int t = warp; t < ntask; t += nwarps) {
        int h = 0, tt = t;
        while (tt >= ctl->hd[h].nchunk6 * slabs) { tt -= ctl->hd[h].nchunk6 * slabs; ++h; }
        // consecutive warps take the slabs of the same chunk: they stream the same pixels
        const int q = tt / slabs;
        const int ch = (tt - q * slabs) * 128 + lane * 4;
        const bool active = ch < C;
        const int PH = P.PH[h], PW = P.PW[h];
        const AxisTab &yt = ctl->hd[h].tab[0];
        const AxisTab &xt = ctl->hd[h].tab[1];
        const int NX = ctl->hd[h].nmax[1];
        const int pa = ctl->hd[h].cstart6[q];
        const unsigned long long cnt = ctl->hd[h].ccnt6[q];
        // lanes past the last channel of a partial slab read channel 0 and store nothing
        const float *p0 = img + (size_t)ctl->hd[h].cx06[q] * C + (active ? ch : 0);
        // L1 prefetch of a row: kSL cells x 512 B = 4 lines of 128 B per cell, one per lane
        const float *pf0 = img + (size_t)(ctl->hd[h].cx06[q] + (lane >> 2)) * C + (ch - lane * 4) + (lane & 3) * 32;
        const bool pf_on = (lane >> 2) < kSL && (ch - lane * 4) + (lane & 3) * 32 < C;
        const int last_row = ctl->wmax[0];
        float *out = P.pooled[h] + ((size_t)c.r * PH * PW + pa) * C + ch;
        float4 A[kSL], B[kSL];
#pragma unroll
        for (int s = 0; s < kSL; ++s) A[s] = B[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        int rowA = -1, rowB = -1;
        for (int ph = 0; ph < PH; ++ph, out += (size_t)PW * C) {
            const int n = yt.n[ph];
            float4 V[kSL];
#pragma unroll
            for (int s = 0; s < kSL; ++s) V[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n > 0 && NX > 0) {
                const int lo = yt.lo[ph];
                const float4 wy = yt.w[ph];
#pragma unroll
                for (int j = 0; j < kNT; ++j) {
                    if (j < n) {
                        const float w = j == 0 ? wy.x : (j == 1 ? wy.y : (j == 2 ? wy.z : wy.w));
                        const int y = lo + j;
                        if (y == rowB) {
#pragma unroll
                            for (int s = 0; s < kSL; ++s) fma4(V[s], w, B[s]);
                        } else if (y == rowA) {
#pragma unroll
                            for (int s = 0; s < kSL; ++s) fma4(V[s], w, A[s]);
                        } else {
                            // a row not seen yet: B slides to A, the new row arrives in B, and the
                            // row below it is requested into L1 for when the footprints get there
                            const float *p = p0 + (size_t)y * row_stride;
#pragma unroll
                            for (int s = 0; s < kSL; ++s) {
                                A[s] = B[s];
                                B[s] = ldg_nc128(p + (kC ? s * kC : s * C));
                            }
                            rowA = rowB;
                            rowB = y;
                            if (P.prefetch != -1 && pf_on && y < last_row) prefetch_l1(pf0 + (size_t)(y + 1) * row_stride);
#pragma unroll
                            for (int s = 0; s < kSL; ++s) fma4(V[s], w, B[s]);
                        }
                    }
                }
            }
            const float4 *wp = &xt.w[pa];
            if (NX <= 2) fwd6_bin_pass<2>(V, cnt, wp, out, C, active);
            else if (NX == 3) fwd6_bin_pass<3>(V, cnt, wp, out, C, active);
            else fwd6_bin_pass<4>(V, cnt, wp, out, C, active);
        }
    }
}

__global__ void __launch_bounds__(kMaxThreads, RPOOL_MIN_BLOCKS)
rpool_forward_slide_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);

    load_record(P, ctl);
    RoiCtx c;
    ctx_from_record(P, ctl, c);
    const int need = kRecValid | kRecShape | kRecFits;
    if ((ctl->flags & need) != need || P.force_path == kPathGeneric || !pointers_aligned(P, c.L)) {
        generic_forward(P);
        return;
    }
    if (P.C == 256) fwd_slide_tasks<256>(P, c, ctl);
    else fwd_slide_tasks<0>(P, c, ctl);
}

}  // namespace rpool
