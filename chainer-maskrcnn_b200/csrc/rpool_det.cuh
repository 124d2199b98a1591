// rpool_det.cuh -- deterministic backward: every gradient cell has ONE owner (sm_100a).
//
// The atomic kernels let the RoIs scatter; here the feature map gathers.  A CTA owns a
// strip of one map row (kDetTiles tiles of kSW columns, all channels); its warps take the
// 64-channel slabs.  The CTA walks the RoIs of its (image, level) group in schedule order
// (the group is a contiguous run of slots, rpool_plan's gstart), keeps those whose window
// meets the strip, and for each of them and each tile adds
//     G[x] += sum over bins pw near x of wx[pw][x] * (sum over bin rows ph covering y of wy[ph][y] * gy[ph][pw])
// into 8 register accumulators per lane, straight from gy and the RoI's record (the same
// footprint tables the other kernels use).  Every cell is then written exactly once:
// no scratch, no zero fill, no atomics, and the order of every sum is fixed by the
// schedule, so the result is bit-identical from run to run and for any CTA size.
// The reference's own CuPy backward is the same formulation, one thread per element and
// a loop over ALL RoIs (roi_align_2d.py:201-276); here it is tiled, table-driven and
// restricted to the RoIs that can touch the strip.
#pragma once
#include "rpool_stream.cuh"

namespace rpool {

constexpr int kDetThreads = 128;
constexpr int kDetBatch = 128;      // slots examined per compaction round

struct DetParams {
    LevelDev lvl[kMaxLevels];
    long long cta_base[kMaxLevels + 1];   // first CTA of every level, coarse levels first
    int strips[kMaxLevels];               // CTAs per map row
    int tiles[kMaxLevels];                // tiles of kSW columns per CTA
    int n_levels, C, accumulate;
    int n_heads;
    int PH[kMaxHeads], PW[kMaxHeads];
    const float *gy[kMaxHeads];
    const int *gstart;                    // first slot of every (image, level) group
    const unsigned char *recs;            // per-slot records (backward geometry)
    int rec_stride;
    int R;
    int *det_err;
};

template <int K>
__device__ __forceinline__ void det_bins_at(float2 (&G)[kSW], int n, const float *&gyp, int row_step,
                                            int pa, int pb, const float *s_wy, const float4 *&wx, int C,
                                            bool active)
{
    // bins whose footprint starts K columns from the tile's first column (K may be negative:
    // the taps left of the tile belong to the neighbouring tile's owner)
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        float2 z = make_float2(0.f, 0.f);
        const float *g = gyp + (size_t)pa * row_step;
#pragma unroll 4
        for (int ph = pa; ph < pb; ++ph, g += row_step) {
            const float w = s_wy[ph];
            const float2 v = active ? __ldg(reinterpret_cast<const float2 *>(g)) : make_float2(0.f, 0.f);
            z.x = fmaf(w, v.x, z.x);
            z.y = fmaf(w, v.y, z.y);
        }
        const float4 w = __ldg(wx);
        if (K + 0 >= 0 && K + 0 < kSW) fma2(G[K + 0 < 0 ? 0 : (K + 0 >= kSW ? 0 : K + 0)], w.x, z);
        if (K + 1 >= 0 && K + 1 < kSW) fma2(G[K + 1 < 0 ? 0 : (K + 1 >= kSW ? 0 : K + 1)], w.y, z);
        if (K + 2 >= 0 && K + 2 < kSW) fma2(G[K + 2 < 0 ? 0 : (K + 2 >= kSW ? 0 : K + 2)], w.z, z);
        if (K + 3 >= 0 && K + 3 < kSW) fma2(G[K + 3 < 0 ? 0 : (K + 3 >= kSW ? 0 : K + 3)], w.w, z);
        ++wx;
        gyp += C;
    }
}

__global__ void __launch_bounds__(kDetThreads)
rpool_backward_det_kernel(const __grid_constant__ DetParams p)
{
    __shared__ int s_slot[kDetBatch];
    __shared__ int s_wcount[kDetThreads / 32], s_n;
    __shared__ float s_wy[kDetThreads / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kDetThreads / 32;

    // which strip: coarse levels come first in the launch (their strips do the most work)
    int li = 0;
    while (li + 1 < p.n_levels && (long long)blockIdx.x >= p.cta_base[li + 1]) ++li;
    const int l = p.n_levels - 1 - li;
    const LevelDev L = p.lvl[l];
    long long idx = (long long)blockIdx.x - p.cta_base[li];
    const int sx = (int)(idx % p.strips[l]);
    idx /= p.strips[l];
    const int y = (int)(idx % L.H);
    const int b = (int)(idx / L.H);
    const int ntile = p.tiles[l];
    const int xs = sx * ntile * kSW;                         // first column of the strip
    int xe = xs + ntile * kSW - 1;                           // last
    xe = xe < L.W - 1 ? xe : L.W - 1;
    const int key = b * p.n_levels + l;
    const int g0 = p.R > 0 ? p.gstart[key] : 0, g1 = p.R > 0 ? p.gstart[key + 1] : 0;
    const int C = p.C;
    const int slabs = (C + kSlabCh - 1) / kSlabCh;
    float *grow = L.data + (((size_t)b * L.H + y) * L.W) * C;

    for (int tile = 0; tile < ntile; ++tile) {
        const int x0 = xs + tile * kSW;
        if (x0 >= L.W) break;
        const int rounds = (slabs + nwarps - 1) / nwarps;
        for (int rd = 0; rd < rounds; ++rd) {
            const int slab = rd * nwarps + warp;
            const int ch = slab * kSlabCh + lane * 2;
            const bool active = slab < slabs && ch < C;
            float2 G[kSW];
#pragma unroll
            for (int s = 0; s < kSW; ++s) G[s] = make_float2(0.f, 0.f);
            for (int batch = g0; batch < g1; batch += kDetBatch) {
                // ordered compaction of the slots whose window meets this tile's row and columns
                const int slot = batch + tid;
                bool hit = false;
                if (slot < g1) {
                    const int4 a = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride));
                    const int4 f = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride) + 1);
                    // a = wmin[0], wmin[1], wmax[0], wmax[1];  f = r, lvl, b, flags
                    const bool valid = (f.w & kRecValid) != 0;
                    const bool table = (f.w & (kRecShape | kRecFits)) == (kRecShape | kRecFits);
                    if (valid && !table) atomicExch(p.det_err, 1);      // needs the generic path: not orderable
                    hit = valid && table && a.x <= y && y <= a.z && a.y <= x0 + kSW - 1 && a.w >= x0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_wcount[warp] = __popc(m);
                __syncthreads();
                int pos = __popc(m & ((1u << lane) - 1u));
                for (int w = 0; w < warp; ++w) pos += s_wcount[w];
                if (hit) s_slot[pos] = slot;
                if (tid == 0) {
                    int n = 0;
                    for (int w = 0; w < nwarps; ++w) n += s_wcount[w];
                    s_n = n;
                }
                __syncthreads();
                const int n_hit = s_n;
                // (RoI, pooled size) pairs in order; the table words of the next pair are
                // requested while the current one is being added (they come from L2)
                struct Pre { int r, lo, n, xlo, xn; float4 wy; };
                auto fetch = [&](int i) {
                    Pre q;
                    q.r = 0; q.lo = 0; q.n = 0; q.xlo = 0; q.xn = 0; q.wy = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < n_hit * p.n_heads) {
                        const int e = i / p.n_heads, h = i - e * p.n_heads;
                        const BlockCtl *rec = reinterpret_cast<const BlockCtl *>(p.recs + (size_t)s_slot[e] * p.rec_stride);
                        const HeadCtl *hd = &rec->hd[h];
                        q.r = __ldg(&rec->r);
                        if (lane < p.PH[h]) {
                            q.lo = __ldg(&hd->tab[0].lo[lane]);
                            q.n = __ldg(&hd->tab[0].n[lane]);
                            q.wy = __ldg(&hd->tab[0].w[lane]);
                        }
                        if (lane < p.PW[h]) {
                            q.xlo = __ldg(&hd->tab[1].lo[lane]);
                            q.xn = __ldg(&hd->tab[1].n[lane]);
                        }
                    }
                    return q;
                };
                Pre nxt = fetch(0);
                for (int i = 0; i < n_hit * p.n_heads; ++i) {
                    const Pre cur = nxt;
                    nxt = fetch(i + 1);
                    const int e = i / p.n_heads, h = i - e * p.n_heads;
                    const BlockCtl *rec = reinterpret_cast<const BlockCtl *>(p.recs + (size_t)s_slot[e] * p.rec_stride);
                    const HeadCtl *hd = &rec->hd[h];
                    const int PH = p.PH[h], PW = p.PW[h];
                    // bin rows covering map row y, and their weights on it
                    const bool cov = lane < PH && cur.lo <= y && y < cur.lo + cur.n;
                    const unsigned my = __ballot_sync(0xffffffffu, cov);
                    if (my == 0) continue;
                    const int k = y - cur.lo;
                    __syncwarp();
                    s_wy[warp][lane] = !cov ? 0.f : (k == 0 ? cur.wy.x : (k == 1 ? cur.wy.y : (k == 2 ? cur.wy.z : cur.wy.w)));
                    __syncwarp();
                    const int pa = __ffs(my) - 1, pb = 32 - __clz(my);
                    // bins whose footprint meets the tile's columns, by first-column offset
                    const bool hx = lane < PW && cur.xn > 0 && cur.xlo <= x0 + kSW - 1 && cur.xlo + cur.xn - 1 >= x0;
                    const unsigned mx = __ballot_sync(0xffffffffu, hx);
                    if (mx == 0) continue;
                    const int off = cur.xlo - x0;                  // in [-3, 7] for the bins that meet the tile
                    const int qa = __ffs(mx) - 1;
                    const float *gyp = p.gy[h] + (((size_t)cur.r * PH) * PW + qa) * C + (active ? ch : 0);
                    const float4 *wx = &hd->tab[1].w[qa];
                    const int row_step = PW * C;
#define RPOOL_DET_AT(K) det_bins_at<K>(G, __popc(__ballot_sync(0xffffffffu, hx && off == (K))), gyp, row_step, \
                                       pa, pb, s_wy[warp], wx, C, active)
                    RPOOL_DET_AT(-3); RPOOL_DET_AT(-2); RPOOL_DET_AT(-1); RPOOL_DET_AT(0);
                    RPOOL_DET_AT(1); RPOOL_DET_AT(2); RPOOL_DET_AT(3); RPOOL_DET_AT(4);
                    RPOOL_DET_AT(5); RPOOL_DET_AT(6); RPOOL_DET_AT(7);
#undef RPOOL_DET_AT
                }
                __syncthreads();        // the list is consumed before the next round overwrites it
            }
            if (active) {
                float *dst = grow + (size_t)x0 * C + ch;
#pragma unroll
                for (int s = 0; s < kSW; ++s) {
                    if (x0 + s < L.W) {
                        float2 v = G[s];
                        if (p.accumulate) {
                            const float2 o = *reinterpret_cast<const float2 *>(dst + (size_t)s * C);
                            v.x += o.x; v.y += o.y;
                        }
                        *reinterpret_cast<float2 *>(dst + (size_t)s * C) = v;
                    }
                }
            }
        }
    }
}

}  // namespace rpool
