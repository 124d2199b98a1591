// rpool_stream.cuh -- the "stream" variants of the two pooling kernels (sm_100a).
//
// Same per-RoI records, same separable arithmetic as the "rows" kernels in
// rpool_kernels.cuh, but the work is cut the other way: a warp task is one
// (pooled size, chunk of bins, 64-channel slab) and walks ALL bin rows of the RoI, so
// every byte that comes from global memory is requested exactly once per task:
//   forward   the window rows of the chunk stream through a 4-row ring in shared memory
//             (cp.async, rows requested up to 3 ahead of the bin row being produced);
//             a bin row reads its 2-4 footprint rows from the ring -- the "rows" kernel
//             re-read every window row from L1 for each of the ~3.7 bin rows covering it;
//   backward  the upstream-gradient bin rows stream through a double buffer (cp.async,
//             one row ahead); each is column-transformed once (bins -> the chunk's 8
//             window columns) and added to up to 4 sliding row accumulators in
//             registers, which leave with one vector reduction per cell when the
//             footprints have moved past them -- the "rows" kernel re-read every bin
//             row from L2 for each of the 2-3 window rows it covers.
// A lane owns 2 channels (64-bit shared-memory accesses, 256 B per warp access): half
// the registers of the 4-channel layout, twice the tasks per RoI.
#pragma once
#include "rpool_kernels.cuh"

namespace rpool {

constexpr int kRing = 4;                 // window rows a forward task keeps in shared memory (>= kNT)
constexpr int kSlabCh = 64;              // channels per task
constexpr int kCellBytes = kSlabCh * 4;  // one cell of a slab
constexpr int kRowBytes = kSW * kCellBytes;
constexpr int kFwdStreamWarpBytes = kRing * kRowBytes;            // 8 KB
constexpr int kBwdStreamWarpBytes = 2 * kPBwd * kCellBytes;       // 8 KB: two gy bin rows of <= 16 bins

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, unsigned src_bytes)
{
    // src_bytes < 16: the rest is zero-filled (0: nothing is read)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int pending)
{
    // wait until at most `pending` of this thread's groups are still in flight
    if (pending >= 3) cp_async_wait<3>();
    else if (pending == 2) cp_async_wait<2>();
    else if (pending == 1) cp_async_wait<1>();
    else cp_async_wait<0>();
}

__device__ __forceinline__ float2 lds64(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void stg_stream64(float *p, float2 v)
{
    asm volatile("st.global" RPOOL_ST_POLICY ".v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *p, float2 v)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void fma2(float2 &a, float w, const float2 &v)
{
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y);
}
__device__ __forceinline__ float2 mul2(float w, const float2 &v) { return make_float2(w * v.x, w * v.y); }

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int NX, int R>
__device__ __forceinline__ float2 taps2(const float2 (&V)[kSW], const float4 w)
{
    float2 o = mul2(w.x, V[R]);
    if (NX > 1 && R + 1 < kSW) fma2(o, w.y, V[R + 1]);
    if (NX > 2 && R + 2 < kSW) fma2(o, w.z, V[R + 2]);
    if (NX > 3 && R + 3 < kSW) fma2(o, w.w, V[R + 3]);
    return o;
}

template <int NX, int R>
__device__ __forceinline__ void fwd2_bins_at(const float2 (&V)[kSW], unsigned long long cnt,
                                             const float4 *&wp, float *&o_ptr, int C, bool active)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float2 o = taps2<NX, R>(V, *wp);
        if (active) stg_stream64(o_ptr, o);
        ++wp;
        o_ptr += C;
    }
}

template <int NX>
__device__ __forceinline__ void fwd2_bin_pass(const float2 (&V)[kSW], unsigned long long cnt,
                                              const float4 *wp, float *o_ptr, int C, bool active)
{
    fwd2_bins_at<NX, 0>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 1>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 2>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 3>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 4>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 5>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 6>(V, cnt, wp, o_ptr, C, active);
    fwd2_bins_at<NX, 7>(V, cnt, wp, o_ptr, C, active);
}

template <int kC>
__device__ __forceinline__ void fwd_stream_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl,
                                                 uint32_t ring)
{
    const int C = kC ? kC : P.C;
    const int slabs = (C + kSlabCh - 1) / kSlabCh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int row_stride = c.L.W * C;
    const float *img = c.L.data + (size_t)c.b * c.L.H * row_stride;
    // copy side: a lane moves 16 bytes (4 channels) of one of two adjacent columns
    const int cp_col = lane >> 4, cp_ch = (lane & 15) * 4;
    const uint32_t cp_dst = ring + (uint32_t)(cp_col * kCellBytes + (lane & 15) * 16);
    const uint32_t rd = ring + (uint32_t)lane * 8u;

    int ntask = 0;
    for (int h = 0; h < P.n_heads; ++h) ntask += ctl->hd[h].nchunk * slabs;
    for (int t = warp; t < ntask; t += nwarps) {
        int h = 0, tt = t;
        while (tt >= ctl->hd[h].nchunk * slabs) { tt -= ctl->hd[h].nchunk * slabs; ++h; }
        const int q = tt / slabs;
        const int ch0 = (tt - q * slabs) * kSlabCh;
        const int ch = ch0 + lane * 2;
        const bool active = ch < C;
        const unsigned cp_bytes = (ch0 + cp_ch < C) ? 16u : 0u;
        const int PH = P.PH[h], PW = P.PW[h];
        const AxisTab &yt = ctl->hd[h].tab[0];
        const AxisTab &xt = ctl->hd[h].tab[1];
        const int NX = ctl->hd[h].nmax[1];
        const int pa = ctl->hd[h].cstart[q];
        const unsigned long long cnt = ctl->hd[h].ccnt[q];
        // rows this head touches: first footprint row of the first non-empty bin row, last of the last
        int first = 0x7fffffff, last = -1;
        for (int ph = 0; ph < PH; ++ph) {
            const int n = yt.n[ph];
            if (n > 0) {
                if (first == 0x7fffffff) first = yt.lo[ph];
                last = yt.lo[ph] + n - 1;
            }
        }
        const float *src0 = img + (size_t)(ctl->hd[h].cx0[q] + cp_col) * C + (cp_bytes ? ch0 + cp_ch : 0);
        float *out = P.pooled[h] + ((size_t)c.r * PH * PW + pa) * C + ch;
        int yl = first - 1;                    // last row requested
        __syncwarp();                          // the previous task's reads of the ring are over
        for (int ph = 0; ph < PH; ++ph, out += (size_t)PW * C) {
            const int n = yt.n[ph];
            float2 V[kSW];
#pragma unroll
            for (int s = 0; s < kSW; ++s) V[s] = make_float2(0.f, 0.f);
            if (n > 0 && NX > 0) {
                const int lo = yt.lo[ph];
                const int hi = lo + n - 1;
                // rows below lo are dead: request rows up to lo + kRing - 1 (their slots are free)
                __syncwarp();
                int target = lo + kRing - 1;
                target = target < last ? target : last;
                while (yl < target) {
                    ++yl;
                    const float *src = src0 + (size_t)yl * row_stride;
                    const uint32_t dst = cp_dst + (uint32_t)((yl - first) & (kRing - 1)) * kRowBytes;
#pragma unroll
                    for (int k = 0; k < kSW / 2; ++k)
                        cp_async16(dst + k * 2 * kCellBytes, src + (kC ? k * 2 * kC : k * 2 * C), cp_bytes);
                    cp_async_commit();
                }
                cp_async_wait_pending(yl - hi);
                __syncwarp();
                const float4 wy = yt.w[ph];
#pragma unroll
                for (int j = 0; j < kNT; ++j) {
                    if (j < n) {
                        const float w = j == 0 ? wy.x : (j == 1 ? wy.y : (j == 2 ? wy.z : wy.w));
                        const uint32_t a = rd + (uint32_t)((lo + j - first) & (kRing - 1)) * kRowBytes;
#pragma unroll
                        for (int s = 0; s < kSW; ++s) fma2(V[s], w, lds64(a + s * kCellBytes));
                    }
                }
            }
            const float4 *wp = &xt.w[pa];
            if (NX <= 2) fwd2_bin_pass<2>(V, cnt, wp, out, C, active);
            else if (NX == 3) fwd2_bin_pass<3>(V, cnt, wp, out, C, active);
            else fwd2_bin_pass<4>(V, cnt, wp, out, C, active);
        }
        cp_async_wait<0>();
    }
}

#ifndef RPOOL_STREAM_MIN_BLOCKS
#define RPOOL_STREAM_MIN_BLOCKS 3
#endif

__global__ void __launch_bounds__(kMaxThreads, RPOOL_STREAM_MIN_BLOCKS)
rpool_forward_stream_kernel(const __grid_constant__ KParams P)
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
    const int kCtlBytes = (rec_bytes(P.n_heads) + 127) & ~127;
    const uint32_t ring = smem_u32(smem_raw + kCtlBytes) + (uint32_t)(threadIdx.x >> 5) * kFwdStreamWarpBytes;
    if (P.C == 256) fwd_stream_tasks<256>(P, c, ctl, ring);
    else fwd_stream_tasks<0>(P, c, ctl, ring);
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
template <int NX, int R>
__device__ __forceinline__ void bwd2_bins_at(float2 (&T)[kSW], unsigned long long cnt, const float4 *&wp,
                                             uint32_t &zp)
{
    const int n = (int)((cnt >> (8 * R)) & 0xffull);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float4 w = *wp;
        const float2 z = lds64(zp);
        fma2(T[R], w.x, z);
        if (NX > 1 && R + 1 < kSW) fma2(T[R + 1], w.y, z);
        if (NX > 2 && R + 2 < kSW) fma2(T[R + 2], w.z, z);
        if (NX > 3 && R + 3 < kSW) fma2(T[R + 3], w.w, z);
        ++wp;
        zp += kCellBytes;
    }
}

template <int NX>
__device__ __forceinline__ void bwd2_col_pass(float2 (&T)[kSW], unsigned long long cnt, const float4 *wp,
                                              uint32_t zp)
{
    bwd2_bins_at<NX, 0>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 1>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 2>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 3>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 4>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 5>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 6>(T, cnt, wp, zp);
    bwd2_bins_at<NX, 7>(T, cnt, wp, zp);
}

template <int kC>
__device__ __forceinline__ void bwd_stream_tasks(const KParams &P, const RoiCtx &c, const BlockCtl *ctl,
                                                 uint32_t strip)
{
    const int C = kC ? kC : P.C;
    const int slabs = (C + kSlabCh - 1) / kSlabCh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int row_stride = c.L.W * C;
    float *img = c.L.data + (size_t)c.b * c.L.H * row_stride;
    const int cp_bin = lane >> 4, cp_ch = (lane & 15) * 4;
    const uint32_t cp_dst = strip + (uint32_t)(cp_bin * kCellBytes + (lane & 15) * 16);
    const uint32_t rd = strip + (uint32_t)lane * 8u;

    int ntask = 0;
    for (int h = 0; h < P.n_heads; ++h) ntask += ctl->hd[h].nchunk * slabs;
    for (int t = warp; t < ntask; t += nwarps) {
        int h = 0, tt = t;
        while (tt >= ctl->hd[h].nchunk * slabs) { tt -= ctl->hd[h].nchunk * slabs; ++h; }
        const int q = tt / slabs;
        const int ch0 = (tt - q * slabs) * kSlabCh;
        const int ch = ch0 + lane * 2;
        const bool active = ch < C;
        const unsigned cp_bytes = (ch0 + cp_ch < C) ? 16u : 0u;
        const int PH = P.PH[h], PW = P.PW[h];
        const AxisTab &yt = ctl->hd[h].tab[0];
        const AxisTab &xt = ctl->hd[h].tab[1];
        const int NX = ctl->hd[h].nmax[1];
        if (NX == 0) continue;
        const int pa = ctl->hd[h].cstart[q], pb = ctl->hd[h].cstart[q + 1];
        const int nb = pb - pa;
        const unsigned long long cnt = ctl->hd[h].ccnt[q];
        const unsigned m = ctl->hd[h].cmask[q];
        float *gcol = img + (size_t)ctl->hd[h].cx0[q] * C + ch;              // row 0 of the chunk's span
        const float *gy0 = P.pooled[h] + ((size_t)c.r * PH * PW + pa + cp_bin) * C + (cp_bytes ? ch0 + cp_ch : 0);
        const size_t gy_step = (size_t)PW * C;
        // pairs of bins per copy instruction; an odd last bin is copied by the low half-warp only
        const int npair = (nb + 1) >> 1;
        auto request = [&](int ph) {
            const float *src = gy0 + (size_t)ph * gy_step;
            const uint32_t dst = cp_dst + (uint32_t)(ph & 1) * (kPBwd * kCellBytes);
            for (int k = 0; k < npair; ++k) {
                const unsigned nbytes = (2 * k + cp_bin < nb) ? cp_bytes : 0u;
                cp_async16(dst + k * 2 * kCellBytes, nbytes ? src + (size_t)k * 2 * C : P.pooled[h], nbytes);
            }
            cp_async_commit();
        };
        // sliding accumulators: G[k] belongs to window row yb + k
        float2 G[kNT][kSW];
#pragma unroll
        for (int k = 0; k < kNT; ++k)
#pragma unroll
            for (int s = 0; s < kSW; ++s) G[k][s] = make_float2(0.f, 0.f);
        int yb = 0x7fffffff;        // no row yet
        unsigned touched = 0;       // bit k: G[k] has received something
        auto retire = [&]() {
            // G[0] leaves: one vector reduction per touched cell of window row yb
            if (touched & 1u) {
                float *gp = gcol + (size_t)yb * row_stride;
#pragma unroll
                for (int s = 0; s < kSW; ++s)
                    if (active && ((m >> s) & 1u)) red_add_v2(gp + (kC ? s * kC : s * C), G[0][s]);
            }
#pragma unroll
            for (int k = 0; k + 1 < kNT; ++k)
#pragma unroll
                for (int s = 0; s < kSW; ++s) G[k][s] = G[k + 1][s];
#pragma unroll
            for (int s = 0; s < kSW; ++s) G[kNT - 1][s] = make_float2(0.f, 0.f);
            touched >>= 1;
            ++yb;
        };
        __syncwarp();               // the previous task's reads of the strip are over
        request(0);
        for (int ph = 0; ph < PH; ++ph) {
            if (ph + 1 < PH) request(ph + 1);
            const int n = yt.n[ph];
            if (n > 0) {
                const int lo = yt.lo[ph];
                if (yb == 0x7fffffff) yb = lo;
                while (yb < lo) {
                    if (touched == 0) { yb = lo; break; }     // nothing pending: jump over the gap
                    retire();
                }
            }
            if (ph + 1 < PH) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            if (n > 0) {
                // column pass: this bin row -> the chunk's kSW window columns
                float2 T[kSW];
#pragma unroll
                for (int s = 0; s < kSW; ++s) T[s] = make_float2(0.f, 0.f);
                const float4 *wp = &xt.w[pa];
                const uint32_t zp = rd + (uint32_t)(ph & 1) * (kPBwd * kCellBytes);
                if (NX <= 2) bwd2_col_pass<2>(T, cnt, wp, zp);
                else if (NX == 3) bwd2_col_pass<3>(T, cnt, wp, zp);
                else bwd2_col_pass<4>(T, cnt, wp, zp);
                // row pass: add to the accumulators of the footprint rows
                const float4 wy = yt.w[ph];
#pragma unroll
                for (int j = 0; j < kNT; ++j) {
                    if (j < n) {
                        const float w = j == 0 ? wy.x : (j == 1 ? wy.y : (j == 2 ? wy.z : wy.w));
#pragma unroll
                        for (int s = 0; s < kSW; ++s) fma2(G[j][s], w, T[s]);
                    }
                }
                touched |= (1u << n) - 1u;
            }
            __syncwarp();           // every lane has read buffer ph & 1 before it is requested again
        }
        while (touched) retire();
    }
}

#ifndef RPOOL_STREAM_MIN_BLOCKS_BWD
#define RPOOL_STREAM_MIN_BLOCKS_BWD 2      // 64 registers of sliding accumulators per lane
#endif

__global__ void __launch_bounds__(kMaxThreads, RPOOL_STREAM_MIN_BLOCKS_BWD)
rpool_backward_stream_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockCtl *ctl = reinterpret_cast<BlockCtl *>(smem_raw);
    const int kCtlBytes = (rec_bytes(P.n_heads) + 127) & ~127;

    load_record(P, ctl);
    RoiCtx c;
    ctx_from_record(P, ctl, c);
    if (!c.valid) return;
    const int need = kRecShape | kRecFits;
    bool table_ok = (ctl->flags & need) == need && P.force_path != kPathGeneric && pointers_aligned(P, c.L);
    for (int h = 0; h < P.n_heads; ++h) table_ok = table_ok && P.PH[h] <= kPBwd && P.PW[h] <= kPBwd;
    if (!table_ok) {
        generic_backward(P);
        return;
    }
    if (ctl->wmax[1] < ctl->wmin[1] || ctl->wmax[0] < ctl->wmin[0]) return;
    const uint32_t strip = smem_u32(smem_raw + kCtlBytes) + (uint32_t)(threadIdx.x >> 5) * kBwdStreamWarpBytes;
    if (P.C == 256) bwd_stream_tasks<256>(P, c, ctl, strip);
    else bwd_stream_tasks<0>(P, c, ctl, strip);
}

}  // namespace rpool
