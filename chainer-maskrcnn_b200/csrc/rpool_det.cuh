// rpool_det.cuh -- deterministic backward: every gradient cell has ONE owner (sm_100a).
//
// The atomic kernels let the RoIs scatter; here the feature map gathers.  A CTA owns a
// strip of one map row: `tiles` tiles of kSW columns, all channels.  For every tile it
// walks the RoIs of its (image, level) group in schedule order (the group is a contiguous
// run of slots, rpool_plan's gstart), keeps those whose window meets the tile (ordered
// compaction), and adds, for each of them, pooled size by pooled size,
//     G[x] += sum over bins pw near x of wx[pw][x] * Z[pw],
//     Z[pw] = sum over bin rows ph covering y of wy[ph][y] * gy[ph][pw]
// into 8 register accumulators per lane (a lane owns 2 channels of a 64-channel slab),
// straight from gy and the RoI's record (the footprint tables every kernel uses):
//   row pass     Z for 8 bins at a time in registers, all covering bin rows' loads in
//                flight together, then into the warp's strip in shared memory;
//   column pass  the bins that meet the tile are walked in order; the offset of a bin's
//                footprint from the tile's first column never decreases, so the pass is a
//                static sequence over the 11 possible offsets (-3..7), each running a
//                ballot-counted number of bins with compile-time accumulator indices.
// Every cell is then written exactly once: no scratch, no zero fill, no atomics.
// Work split inside the CTA (fixed by the data, never by timing):
//   few RoIs on the tile   warp w takes slab w and adds every RoI in schedule order;
//   many (>= kDetSplit)    the slabs are taken one after the other and the CTA's warps
//                          share the RoI list round-robin; their partial sums are added
//                          in warp order.  (The coarsest level has ~200 RoIs per tile:
//                          one warp walking them all would be the critical path.)
// The order of every sum is a function of the schedule alone, so results are
// bit-identical from run to run.  The reference's own CuPy backward is the same
// formulation, one thread per element looping over ALL RoIs (roi_align_2d.py:201-276);
// here it is tiled, table-driven and restricted to the RoIs that can touch the tile.
#pragma once
#include "rpool_stream.cuh"

namespace rpool {

constexpr int kDetThreads = 128;
constexpr int kDetWarps = kDetThreads / 32;
constexpr int kDetBatch = 128;      // slots examined per compaction round
constexpr int kDetSplit = 8;        // RoIs on a tile from which the warps share the list
constexpr int kDetZ = 8;            // bins per register chunk of the row pass

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

struct DetShared {
    int slot[kDetBatch];
    int wcount[kDetWarps], n;
    float wy[kDetWarps][32];
    float4 wx[kDetWarps][32];
    float2 part[kDetWarps][kSW][32];                       // partial sums of the split mode
    unsigned char strip[kDetWarps][kPMax * kCellBytes];    // Z of up to 32 bins per warp
};

template <int K>
__device__ __forceinline__ void det_bins_at(float2 (&G)[kSW], int n, uint32_t &zp, const float4 *&wx)
{
    // bins whose footprint starts K columns from the tile's first column (K < 0: the taps
    // left of the tile belong to the neighbouring tile's owner)
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const float2 z = lds64(zp);
        const float4 w = *wx;
        if (K + 0 >= 0 && K + 0 < kSW) fma2(G[(K + 0) & (kSW - 1)], w.x, z);
        if (K + 1 >= 0 && K + 1 < kSW) fma2(G[(K + 1) & (kSW - 1)], w.y, z);
        if (K + 2 >= 0 && K + 2 < kSW) fma2(G[(K + 2) & (kSW - 1)], w.z, z);
        if (K + 3 >= 0 && K + 3 < kSW) fma2(G[(K + 3) & (kSW - 1)], w.w, z);
        ++wx;
        zp += kCellBytes;
    }
}

// The table words of one (RoI, pooled size): lane ph / pw holds entry ph / pw.
struct DetPre {
    int r, lo, n, xlo, xn;
    float4 wy, wx;
};

__device__ __forceinline__ DetPre det_fetch(const DetParams &p, const int *s_slot, int i, int n_pairs, int lane)
{
    DetPre q;
    q.r = 0; q.lo = 0; q.n = 0; q.xlo = 0; q.xn = 0;
    q.wy = q.wx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_pairs) {
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
            q.wx = __ldg(&hd->tab[1].w[lane]);
        }
    }
    return q;
}

// Adds one (RoI, pooled size) to the tile's accumulators.
__device__ __forceinline__ void det_add(const DetParams &p, DetShared &sh, float2 (&G)[kSW], const DetPre &cur,
                                        int h, int y, int x0, int ch, bool active, int warp, int lane)
{
    const int PH = p.PH[h], PW = p.PW[h], C = p.C;
    // bin rows covering map row y, and their weights on it
    const bool cov = lane < PH && cur.lo <= y && y < cur.lo + cur.n;
    const unsigned my = __ballot_sync(0xffffffffu, cov);
    // bins whose footprint meets the tile's columns
    const bool hx = lane < PW && cur.xn > 0 && cur.xlo <= x0 + kSW - 1 && cur.xlo + cur.xn - 1 >= x0;
    const unsigned mx = __ballot_sync(0xffffffffu, hx);
    if (my == 0 || mx == 0) return;
    const int k = y - cur.lo;
    __syncwarp();
    sh.wy[warp][lane] = !cov ? 0.f : (k == 0 ? cur.wy.x : (k == 1 ? cur.wy.y : (k == 2 ? cur.wy.z : cur.wy.w)));
    sh.wx[warp][lane] = cur.wx;
    __syncwarp();
    const int pa = __ffs(my) - 1, pb = 32 - __clz(my);
    const int qa = __ffs(mx) - 1, nb = __popc(mx);          // the bins that meet the tile are contiguous
    const int off = cur.xlo - x0;                            // in [-3, 7] for those bins
    const int row_step = PW * C;
    const float *g0 = p.gy[h] + (((size_t)cur.r * PH + pa) * PW + qa) * C + (active ? ch : 0);
    const uint32_t strip = smem_u32(&sh.strip[warp][0]) + (uint32_t)lane * 8u;
    // ---- row pass
    for (int j0 = 0; j0 < nb; j0 += kDetZ) {
        float2 Z[kDetZ];
#pragma unroll
        for (int j = 0; j < kDetZ; ++j) Z[j] = make_float2(0.f, 0.f);
        const float *g = g0 + (size_t)j0 * C;
#pragma unroll 2
        for (int ph = pa; ph < pb; ++ph, g += row_step) {
            const float w = sh.wy[warp][ph];
#pragma unroll
            for (int j = 0; j < kDetZ; ++j) {
                if (active && j0 + j < nb) {
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(g + (size_t)j * C));
                    Z[j].x = fmaf(w, v.x, Z[j].x);
                    Z[j].y = fmaf(w, v.y, Z[j].y);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kDetZ; ++j)
            if (j0 + j < nb) {
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};"
                             :: "r"(strip + (uint32_t)(j0 + j) * kCellBytes), "f"(Z[j].x), "f"(Z[j].y) : "memory");
            }
    }
    __syncwarp();
    // ---- column pass
    uint32_t zp = strip;
    const float4 *wx = &sh.wx[warp][qa];
#define RPOOL_DET_AT(K) det_bins_at<K>(G, __popc(__ballot_sync(0xffffffffu, hx && off == (K))), zp, wx)
    RPOOL_DET_AT(-3); RPOOL_DET_AT(-2); RPOOL_DET_AT(-1); RPOOL_DET_AT(0);
    RPOOL_DET_AT(1); RPOOL_DET_AT(2); RPOOL_DET_AT(3); RPOOL_DET_AT(4);
    RPOOL_DET_AT(5); RPOOL_DET_AT(6); RPOOL_DET_AT(7);
#undef RPOOL_DET_AT
}

__device__ __forceinline__ void det_store(const DetParams &p, float *dst, const float2 (&G)[kSW], int x0, int W,
                                          int C)
{
#pragma unroll
    for (int s = 0; s < kSW; ++s) {
        if (x0 + s < W) {
            float2 v = G[s];
            if (p.accumulate) {
                const float2 o = *reinterpret_cast<const float2 *>(dst + (size_t)s * C);
                v.x += o.x; v.y += o.y;
            }
            *reinterpret_cast<float2 *>(dst + (size_t)s * C) = v;
        }
    }
}

__global__ void __launch_bounds__(kDetThreads)
rpool_backward_det_kernel(const __grid_constant__ DetParams p)
{
    extern __shared__ __align__(16) unsigned char det_smem[];
    DetShared &sh = *reinterpret_cast<DetShared *>(det_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // which strip: coarse levels come first in the launch (their tiles do the most work)
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
    const int key = b * p.n_levels + l;
    const int g0 = p.R > 0 ? p.gstart[key] : 0, g1 = p.R > 0 ? p.gstart[key + 1] : 0;
    const int C = p.C;
    const int slabs = (C + kSlabCh - 1) / kSlabCh;
    float *grow = L.data + (((size_t)b * L.H + y) * L.W) * C;

    for (int tile = 0; tile < ntile; ++tile) {
        const int x0 = (sx * ntile + tile) * kSW;
        if (x0 >= L.W) break;
        // Pass 1 counts the RoIs on the tile (that fixes the split mode); pass 2 adds them.
        int total = 0;
        for (int batch = g0; batch < g1; batch += kDetBatch) {
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
            total += __syncthreads_count(hit);
        }
        const bool split = total >= kDetSplit;
        // split: slabs one after the other, warps share the RoIs; else warp <-> slab
        const int rounds = split ? slabs : (slabs + kDetWarps - 1) / kDetWarps;
        for (int rd = 0; rd < rounds; ++rd) {
            const int slab = split ? rd : rd * kDetWarps + warp;
            const int ch = slab * kSlabCh + lane * 2;
            const bool active = slab < slabs && ch < C;
            float2 G[kSW];
#pragma unroll
            for (int s = 0; s < kSW; ++s) G[s] = make_float2(0.f, 0.f);
            int seen = 0;          // RoIs of the tile before this batch (round-robin phase of the split mode)
            for (int batch = g0; batch < g1; batch += kDetBatch) {
                // ordered compaction of the slots whose window meets this tile's row and columns
                const int slot = batch + tid;
                bool hit = false;
                if (slot < g1) {
                    const int4 a = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride));
                    const int4 f = __ldg(reinterpret_cast<const int4 *>(p.recs + (size_t)slot * p.rec_stride) + 1);
                    const bool ok = (f.w & (kRecValid | kRecShape | kRecFits)) == (kRecValid | kRecShape | kRecFits);
                    hit = ok && a.x <= y && y <= a.z && a.y <= x0 + kSW - 1 && a.w >= x0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) sh.wcount[warp] = __popc(m);
                __syncthreads();
                int pos = __popc(m & ((1u << lane) - 1u));
                for (int w = 0; w < warp; ++w) pos += sh.wcount[w];
                if (hit) sh.slot[pos] = slot;
                if (tid == 0) {
                    int n = 0;
                    for (int w = 0; w < kDetWarps; ++w) n += sh.wcount[w];
                    sh.n = n;
                }
                __syncthreads();
                const int n_hit = sh.n;
                // (RoI, pooled size) pairs in order; the table words of the next pair are
                // requested while the current one is being added (they come from L2)
                const int n_pairs = n_hit * p.n_heads;
                int i = 0, step = 1;
                if (split) {
                    // this warp's RoIs: global index on the tile == warp (mod kDetWarps)
                    const int e0 = (warp - seen % kDetWarps + kDetWarps) % kDetWarps;
                    i = e0 * p.n_heads;
                    step = (kDetWarps - 1) * p.n_heads + 1;      // after the last head of a RoI: skip the others' RoIs
                }
                DetPre nxt = det_fetch(p, sh.slot, i, n_pairs, lane);
                while (i < n_pairs) {
                    const DetPre cur = nxt;
                    const int h = i % p.n_heads;
                    const int inext = i + ((h == p.n_heads - 1) ? step : 1);
                    nxt = det_fetch(p, sh.slot, inext, n_pairs, lane);
                    det_add(p, sh, G, cur, h, y, x0, ch, active, warp, lane);
                    i = inext;
                }
                seen += n_hit;
                __syncthreads();        // the list is consumed before the next round overwrites it
            }
            float *dst = grow + (size_t)x0 * C + ch;
            if (!split) {
                if (active) det_store(p, dst, G, x0, L.W, C);
            } else {
                // partial sums, added in warp order: warp w finishes columns w, w + kDetWarps, ...
#pragma unroll
                for (int s = 0; s < kSW; ++s) sh.part[warp][s][lane] = G[s];
                __syncthreads();
#pragma unroll
                for (int s = 0; s < kSW; ++s) {
                    if ((s % kDetWarps) == warp) {
                        float2 a = sh.part[0][s][lane];
#pragma unroll
                        for (int w = 1; w < kDetWarps; ++w) {
                            const float2 q = sh.part[w][s][lane];
                            a.x += q.x; a.y += q.y;
                        }
                        if (active && x0 + s < L.W) {
                            float *d = dst + (size_t)s * C;
                            if (p.accumulate) {
                                const float2 o = *reinterpret_cast<const float2 *>(d);
                                a.x += o.x; a.y += o.y;
                            }
                            *reinterpret_cast<float2 *>(d) = a;
                        }
                    }
                }
                __syncthreads();        // partials are read before the next slab overwrites them
            }
        }
    }
}

}  // namespace rpool
