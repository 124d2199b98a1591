// rpool_api.cu -- extern "C" entry points of librpool_b200.so (see
// include/rpool_b200.h for the contract and the reference lines each replaces).
// Host side only validates, fills kernel parameter blocks and launches; it
// never allocates device memory and never synchronises (except rpool_read_plan).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "rpool_kernels.cuh"

using namespace rpool;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

constexpr int kMaxSmem = 227 * 1024;

#ifndef RPOOL_BUILD_ID
#define RPOOL_BUILD_ID "unknown"
#endif

// rpool_options with the defaults filled in (0 = default in the ABI)
struct Options {
    int threads, order, force_path, split_heads, prefetch, fill_in_tail, bwd_variant;
};
// CTA size of the pooling kernels when opt.cta_threads is 0: 4 warps per RoI, 4 such CTAs resident
// per SM (register-limited).  A problem that does not even fill one wave of those (148 SMs x 4) is
// bound by the latency of its longest CTA, not by throughput: 8 warps per RoI halve that chain
// (configs[0], 512 RoIs: 62 -> 55 us per step; configs[1] and [3] lose 7-9 % with 8 warps).
constexpr int kDefaultThreads = 128;
constexpr int kSmallProblemThreads = 256;
constexpr int kOneWaveRois = 148 * 4;
constexpr int kDefaultPrefetchRows = 2;

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(RPOOL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr, what)                              \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, what); \
    } while (0)

// Launch that may start in the tail of the previous kernel of the stream (programmatic dependent
// launch): once every CTA of that kernel has executed griddepcontrol.launch_dependents
// (allow_dependents_early) or exited.  What the early kernel needs from its predecessors it must
// wait for itself (wait_for_predecessors); a kernel launched normally after it still waits for
// everything before it.
template <typename... Params, typename... Args>
cudaError_t launch_in_tail(void (*kern)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                           const Args &...args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, Params(args)...);
}

// workspace layout (r = R rounded up to 32, nb = key blocks of kKeyBlock RoIs):
//   int32  levels[r] order[r] keys[r] rflags[r] gstart[288] bh[nb * 256] rects[4r]
//   uint64 woff[r] sizes[r] det_total, then int32 det_err (+ padding to 16 bytes)
//   RoI records (rpool_plan_kernel): one set of r records of rec_bytes(n_heads)
//   for the forward geometry, and for RPOOL_COORD_CHAINER -- whose backward
//   coordinates round differently (roi_align_2d.py:164-165) -- a second set
constexpr size_t kGstartInts = 288;  // kPlanMaxKeys + 1, padded
struct Workspace {
    int *levels, *order, *keys, *rflags, *gstart, *bh, *rects;
    unsigned long long *woff, *sizes, *det_total;
    int *det_err;
    unsigned char *recs_fwd, *recs_bwd;
    int rec_stride;
};
size_t ws_round(int R) { return ((size_t)(R > 0 ? R : 1) + 31) & ~(size_t)31; }
size_t ws_key_blocks(int R) { return ((size_t)(R > 0 ? R : 1) + kKeyBlock - 1) / kKeyBlock; }
size_t ws_fixed_bytes(int R)
{
    const size_t r = ws_round(R);
    const size_t n = (4 * r + kGstartInts + ws_key_blocks(R) * kPlanMaxKeys + 4 * r) * sizeof(int) +
                     (2 * r + 1) * sizeof(unsigned long long) + 16;
    return (n + 15) & ~(size_t)15;
}
int rec_sets(int coord_mode) { return coord_mode == RPOOL_COORD_CHAINER ? 2 : 1; }
size_t ws_bytes_ex(int R, int n_heads, int coord_mode)
{
    return ws_fixed_bytes(R) + (size_t)rec_sets(coord_mode) * (size_t)(R > 0 ? R : 0) * (size_t)rec_bytes(n_heads);
}
size_t ws_bytes(int R) { return ws_bytes_ex(R, RPOOL_MAX_HEADS, RPOOL_COORD_CHAINER); }
Workspace ws_split(void *ws, int R, int n_heads, int coord_mode)
{
    const size_t r = ws_round(R);
    Workspace w;
    w.levels = static_cast<int *>(ws);
    w.order = w.levels + r;
    w.keys = w.order + r;
    w.rflags = w.keys + r;
    w.gstart = w.rflags + r;
    w.bh = w.gstart + kGstartInts;
    w.rects = w.bh + ws_key_blocks(R) * kPlanMaxKeys;
    w.woff = reinterpret_cast<unsigned long long *>(w.rects + 4 * r);
    w.sizes = w.woff + r;
    w.det_total = w.sizes + r;
    w.det_err = reinterpret_cast<int *>(w.det_total + 1);
    w.rec_stride = rec_bytes(n_heads);
    w.recs_fwd = static_cast<unsigned char *>(ws) + ws_fixed_bytes(R);
    w.recs_bwd = rec_sets(coord_mode) == 2 ? w.recs_fwd + (size_t)(R > 0 ? R : 0) * (size_t)w.rec_stride : w.recs_fwd;
    return w;
}
Workspace ws_split(void *ws, const rpool_problem *p)
{
    return ws_split(ws, p->n_rois, p->n_heads, p->coord_mode);
}

// Validates rpool_problem.opt and fills in the defaults.
int read_options(const rpool_problem *p, Options &o)
{
    const rpool_options &q = p->opt;
    if (q.cta_threads != 0 && (q.cta_threads < 32 || q.cta_threads > kMaxThreads || q.cta_threads % 32))
        return fail(RPOOL_ERR_INVALID, "opt.cta_threads=%d must be 0 or a multiple of 32 in [32,%d]",
                    q.cta_threads, kMaxThreads);
    if (q.schedule < 0 || q.schedule > RPOOL_SCHED_COARSE_FIRST)
        return fail(RPOOL_ERR_INVALID, "opt.schedule=%d outside [0,3]", q.schedule);
    if (q.force_path < 0 || q.force_path > RPOOL_PATH_TABLE)
        return fail(RPOOL_ERR_INVALID, "opt.force_path=%d outside [0,2]", q.force_path);
    if (q.fuse_heads_backward < 0 || q.fuse_heads_backward > 1)
        return fail(RPOOL_ERR_INVALID, "opt.fuse_heads_backward=%d outside [0,1]", q.fuse_heads_backward);
    if (q.prefetch_rows < -1 || q.prefetch_rows > 16)
        return fail(RPOOL_ERR_INVALID, "opt.prefetch_rows=%d outside [-1,16]", q.prefetch_rows);
    if (q.prefetch_rois < 0 || q.prefetch_rois > 65536)
        return fail(RPOOL_ERR_INVALID, "opt.prefetch_rois=%d outside [0,65536]", q.prefetch_rois);
    if (q.zero_fill_in_tail < 0 || q.zero_fill_in_tail > 1)
        return fail(RPOOL_ERR_INVALID, "opt.zero_fill_in_tail=%d outside [0,1]", q.zero_fill_in_tail);
    if (q.backward_variant < 0 || q.backward_variant > 2)
        return fail(RPOOL_ERR_INVALID, "opt.backward_variant=%d outside [0,2]", q.backward_variant);
    o.bwd_variant = q.backward_variant;
    o.threads = q.cta_threads ? q.cta_threads
                              : (p->n_rois <= kOneWaveRois ? kSmallProblemThreads : kDefaultThreads);
    o.order = q.schedule;
    o.force_path = q.force_path;
    o.split_heads = q.fuse_heads_backward ? 0 : 1;
    o.fill_in_tail = q.zero_fill_in_tail;
    // kernel encoding: -1 off; -1-k row-ahead by k window rows; n >= 0 the whole RoI n slots later
    if (q.prefetch_rois > 0) o.prefetch = q.prefetch_rois - 1;
    else if (q.prefetch_rows < 0) o.prefetch = -1;
    else o.prefetch = -1 - (q.prefetch_rows ? q.prefetch_rows : kDefaultPrefetchRows);

    return RPOOL_OK;
}

int validate(const rpool_problem *p, void *ws, size_t ws_size, bool need_pooled)
{
    if (!p) return fail(RPOOL_ERR_INVALID, "problem is NULL");
    if (p->n_levels < 1 || p->n_levels > RPOOL_MAX_LEVELS)
        return fail(RPOOL_ERR_INVALID, "n_levels=%d outside [1,%d]", p->n_levels, RPOOL_MAX_LEVELS);
    if (p->channels < 1) return fail(RPOOL_ERR_INVALID, "channels=%d", p->channels);
    if (p->feat_layout != RPOOL_NHWC && p->feat_layout != RPOOL_NCHW)
        return fail(RPOOL_ERR_INVALID, "feat_layout=%d", p->feat_layout);
    if (p->pool_layout != RPOOL_NHWC && p->pool_layout != RPOOL_NCHW)
        return fail(RPOOL_ERR_INVALID, "pool_layout=%d", p->pool_layout);
    if (p->roi_format != RPOOL_ROI_XY && p->roi_format != RPOOL_ROI_YX)
        return fail(RPOOL_ERR_INVALID, "roi_format=%d", p->roi_format);
    if (p->n_rois < 0) return fail(RPOOL_ERR_INVALID, "n_rois=%d", p->n_rois);
    if (p->n_rois > 0 && !p->rois) return fail(RPOOL_ERR_INVALID, "rois is NULL");
    for (int l = 0; l < p->n_levels; ++l) {
        const rpool_level &L = p->level[l];
        if (!L.data || L.n_images < 1 || L.height < 1 || L.width < 1)
            return fail(RPOOL_ERR_INVALID, "level %d: data=%p n=%d h=%d w=%d", l, L.data,
                        L.n_images, L.height, L.width);
        if ((double)L.n_images * L.height * L.width * p->channels >= 2147483648.0 * 4)
            return fail(RPOOL_ERR_UNSUPPORTED, "level %d has more than 2^33 elements", l);
        if ((double)L.height * L.width * p->channels >= 2147483648.0)
            return fail(RPOOL_ERR_UNSUPPORTED, "level %d: one image exceeds 2^31 elements", l);
    }
    if (p->n_heads < 1 || p->n_heads > RPOOL_MAX_HEADS)
        return fail(RPOOL_ERR_INVALID, "n_heads=%d outside [1,%d]", p->n_heads, RPOOL_MAX_HEADS);
    for (int h = 0; h < p->n_heads; ++h) {
        if (p->out_h[h] < 1 || p->out_w[h] < 1)
            return fail(RPOOL_ERR_INVALID, "head %d: out %dx%d", h, p->out_h[h], p->out_w[h]);
        if (need_pooled && p->n_rois > 0 && !p->pooled[h])
            return fail(RPOOL_ERR_INVALID, "head %d: pooled pointer is NULL", h);
        if ((double)p->out_h[h] * p->out_w[h] * p->channels >= 2147483648.0)
            return fail(RPOOL_ERR_UNSUPPORTED, "head %d: one pooled RoI exceeds 2^31 elements", h);
    }
    if (p->coord_mode != RPOOL_COORD_CHAINER && p->coord_mode != RPOOL_COORD_CAFFE2)
        return fail(RPOOL_ERR_INVALID, "coord_mode=%d", p->coord_mode);
    if (p->coord_mode == RPOOL_COORD_CHAINER && p->sampling_ratio != 1)
        return fail(RPOOL_ERR_INVALID,
                    "RPOOL_COORD_CHAINER samples once per bin: sampling_ratio must be 1, got %d",
                    p->sampling_ratio);
    if (p->sampling_ratio > 64)
        return fail(RPOOL_ERR_UNSUPPORTED, "sampling_ratio=%d > 64", p->sampling_ratio);
    if (!p->roi_levels && !p->roi_levels_f32 && p->n_levels > 1) {
        if (p->n_thresholds < 0 || p->n_thresholds > RPOOL_MAX_LEVELS)
            return fail(RPOOL_ERR_INVALID, "n_thresholds=%d", p->n_thresholds);
    }
    if (!ws) return fail(RPOOL_ERR_WORKSPACE, "workspace is NULL");
    if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(RPOOL_ERR_WORKSPACE, "workspace must be 16-byte aligned");
    if (ws_size < ws_bytes_ex(p->n_rois, p->n_heads, p->coord_mode))
        return fail(RPOOL_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", ws_size,
                    ws_bytes_ex(p->n_rois, p->n_heads, p->coord_mode));
    return RPOOL_OK;
}

// Fills the kernel parameter block; returns the dynamic shared memory the
// launch needs (control block + per-warp strips [+ transposed tables]).
int fill_params(const rpool_problem *p, const Workspace &w, const Options &o, bool bwd, int threads, KParams &k)
{
    memset(&k, 0, sizeof(k));
    for (int l = 0; l < p->n_levels; ++l) {
        k.lvl[l].data = static_cast<float *>(p->level[l].data);
        k.lvl[l].n_images = p->level[l].n_images;
        k.lvl[l].H = p->level[l].height;
        k.lvl[l].W = p->level[l].width;
        k.lvl[l].scale = p->level[l].spatial_scale;
    }
    k.n_levels = p->n_levels;
    k.C = p->channels;
    k.feat_layout = p->feat_layout;
    k.pool_layout = p->pool_layout;
    k.rois = p->rois;
    k.R = p->n_rois;
    k.roi_format = p->roi_format;
    k.roi_level = w.levels;
    k.order = w.order;
    k.n_heads = p->n_heads;
    int sum_pw = 0;
    for (int h = 0; h < p->n_heads; ++h) {
        k.PH[h] = p->out_h[h];
        k.PW[h] = p->out_w[h];
        k.pooled[h] = static_cast<float *>(p->pooled[h]);
        sum_pw += p->out_w[h] <= kPBwd ? p->out_w[h] : 0;
    }
    k.S = p->sampling_ratio;
    k.mode = p->coord_mode;
    k.force_path = o.force_path;
    const int ctl = (rec_bytes(p->n_heads) + 127) & ~127;
    const int warps = threads / 32;
    k.prefetch = o.prefetch;
    k.reverse = (bwd && o.order == RPOOL_SCHED_DEFAULT) ? 1 : 0;
    k.det = 0;
    k.det_rects = w.rects;
    k.det_woff = w.woff;
    k.det_err = w.det_err;
    k.recs = bwd ? w.recs_bwd : w.recs_fwd;
    k.rec_stride = w.rec_stride;
    k.tail_start = p->n_rois;
    k.tail_parts = 1;
    if (!bwd) return ctl;
    const int ttab = (int)((sizeof(TTab) + 127) & ~(size_t)127);
    k.strip_cols = sum_pw > 0 ? sum_pw : 1;
    return ctl + p->n_heads * ttab + warps * k.strip_cols * 512;
}

// SMs of the current device (cached per device).
int sm_count()
{
    static std::atomic<int> cache[64];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    const bool cached = dev >= 0 && dev < 64;
    if (cached && cache[dev].load(std::memory_order_relaxed) > 0) return cache[dev].load(std::memory_order_relaxed);
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cached) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

// The staged backward kernel serves this (one-pooled-size) problem: returns its grid size, 0 otherwise.
int staged_grid(const rpool_problem *q, const Options &o, int &smem)
{
    if (o.bwd_variant != 2 || q->n_heads != 1 || q->deterministic || o.force_path == kPathGeneric) return 0;
    if (q->feat_layout != RPOOL_NHWC || q->pool_layout != RPOOL_NHWC) return 0;
    if (q->channels % 128 || q->channels > 128 * kStMaxSlabs) return 0;
    if (q->out_h[0] > kPBwd || q->out_w[0] > kPBwd) return 0;
    if (reinterpret_cast<uintptr_t>(q->pooled[0]) & 15) return 0;
    smem = kStCtlBytes + 2 * kStRecBytes + 2 * q->out_h[0] * q->out_w[0] * 512;
    if (smem > kMaxSmem) return 0;
    const int sms = sm_count();
    if (sms <= 0) return 0;
    const int grid = q->n_rois < sms ? q->n_rois : sms;
    if ((q->n_rois + grid - 1) / grid > kStMaxSlots) return 0;
    return grid;
}

// Raises a pooling kernel's dynamic shared memory limit.  The limit is per device
// and only ever needs to grow, so the largest value set so far is remembered per
// (kernel, device) and the driver call is skipped when it already covers `smem`.
constexpr int kSmemCacheDevices = 64;
std::atomic<int> g_smem_set[4][kSmemCacheDevices];

template <typename Kern>
int set_smem(Kern kern, int which, int smem)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
    const bool cached = dev >= 0 && dev < kSmemCacheDevices;
    if (cached && g_smem_set[which][dev].load(std::memory_order_relaxed) >= smem) return RPOOL_OK;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
             "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (cached) {
        int seen = g_smem_set[which][dev].load(std::memory_order_relaxed);
        while (seen < smem && !g_smem_set[which][dev].compare_exchange_weak(seen, smem)) {}
    }
    return RPOOL_OK;
}

// Split tail of a pooling launch (KParams::tail_start / tail_parts): the last resident-CTAs' worth of
// work is handed out in `parts` pieces per RoI, so that the launch drains in 1 / parts of a full CTA's
// duration.  Only for launches of more than two waves; returns the grid size.
template <typename Kern>
int split_tail(Kern kern, int threads, int smem, int parts, KParams &k)
{
    k.tail_start = k.R;
    k.tail_parts = 1;
    if (parts < 2 || k.det || k.prefetch >= 0) return k.R;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return k.R;
    }
    const int resident = per_sm * sm_count();
    if (resident <= 0 || k.R < 2 * resident) return k.R;
    const int split = resident / parts;            // RoIs served in pieces
    k.tail_start = k.R - split;
    k.tail_parts = parts;
    return k.tail_start + split * parts;
}

}  // namespace

extern "C" {

int rpool_version(void) { return RPOOL_VERSION; }

const char *rpool_last_error(void) { return g_err; }

uint64_t rpool_launch_count(void) { return g_launches.load(); }

const char *rpool_build_id(void) { return RPOOL_BUILD_ID; }

int rpool_level_thresholds(float s0, float lvl0, float eps, int k_min, int k_max, float *out)
{
    if (!out || k_max < k_min) return fail(RPOOL_ERR_INVALID, "bad arguments");
    auto level_of = [&](float area) -> float {
        // float32 pipeline of multilevel_region_proposal_network.py:24-30
        volatile float s = sqrtf(area);
        volatile float q = s / s0;
        volatile float a = q + eps;
        volatile float lg = log2f(a);
        volatile float t = lvl0 + lg;
        return floorf(t);
    };
    for (int k = k_min + 1; k <= k_max; ++k) {
        uint32_t lo = 0u;           // level(lo) < k
        uint32_t hi = 0x7f000000u;  // level(hi) >= k  (1.7e38)
        float fhi;
        memcpy(&fhi, &hi, 4);
        if (!(level_of(fhi) >= (float)k))
            return fail(RPOOL_ERR_INVALID, "level %d is unreachable", k);
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float fm;
            memcpy(&fm, &mid, 4);
            if (level_of(fm) >= (float)k) hi = mid; else lo = mid;
        }
        memcpy(&out[k - k_min - 1], &hi, 4);
    }
    return RPOOL_OK;
}

int rpool_assign_levels(const float *boxes, int32_t n, int32_t box_stride, int32_t roi_format,
                        const float *thr, int32_t n_thr, int32_t k_min, int32_t k_cap,
                        float *levels_f32, int32_t *levels_i32, void *stream)
{
    if (n < 0 || (n > 0 && !boxes)) return fail(RPOOL_ERR_INVALID, "boxes");
    if (box_stride != 4 && box_stride != 5) return fail(RPOOL_ERR_INVALID, "box_stride=%d", box_stride);
    if (n_thr < 0 || n_thr > RPOOL_MAX_LEVELS || (n_thr > 0 && !thr))
        return fail(RPOOL_ERR_INVALID, "n_thresholds=%d", n_thr);
    if (roi_format != RPOOL_ROI_XY && roi_format != RPOOL_ROI_YX)
        return fail(RPOOL_ERR_INVALID, "roi_format=%d", roi_format);
    if (n == 0) return RPOOL_OK;
    LevelParams p;
    memset(&p, 0, sizeof(p));
    p.boxes = boxes; p.n = n; p.stride = box_stride; p.roi_format = roi_format;
    for (int t = 0; t < n_thr; ++t) p.thr[t] = thr[t];
    p.n_thr = n_thr; p.k_min = k_min; p.k_cap = k_cap;
    p.out_f = levels_f32; p.out_i = levels_i32;
    rpool_levels_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    CUDA_TRY(cudaGetLastError(), "rpool_levels_kernel launch");
    g_launches++;
    return RPOOL_OK;
}

size_t rpool_workspace_bytes(int32_t n_rois) { return ws_bytes(n_rois); }

size_t rpool_workspace_bytes_ex(int32_t n_rois, int32_t n_heads, int32_t coord_mode)
{
    if (n_heads < 1 || n_heads > RPOOL_MAX_HEADS) n_heads = RPOOL_MAX_HEADS;
    return ws_bytes_ex(n_rois, n_heads, coord_mode);
}

size_t rpool_problem_size(void) { return sizeof(rpool_problem); }

int rpool_plan(const rpool_problem *p, void *ws, size_t ws_size, void *stream)
{
    int rc = validate(p, ws, ws_size, false);
    if (rc) return rc;
    Options o;
    rc = read_options(p, o);
    if (rc) return rc;
    if (p->n_rois == 0) return RPOOL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Workspace w = ws_split(ws, p);
    PlanParams k;
    memset(&k, 0, sizeof(k));
    k.rois = p->rois; k.R = p->n_rois; k.roi_format = p->roi_format;
    k.given_levels = p->roi_levels;
    k.given_levels_f32 = p->roi_levels ? nullptr : p->roi_levels_f32;
    k.n_thr = (p->roi_levels || p->roi_levels_f32 || p->n_levels == 1) ? 0 : p->n_thresholds;
    for (int t = 0; t < k.n_thr; ++t) k.thr[t] = p->level_thresholds[t];
    k.k_min = p->k_min;
    k.n_levels = p->n_levels;
    int nimg = 1;
    for (int l = 0; l < p->n_levels; ++l) nimg = p->level[l].n_images > nimg ? p->level[l].n_images : nimg;
    k.n_images = nimg;
    k.order_mode = o.order;
    k.by_image = ((long long)nimg * p->n_levels <= kPlanMaxKeys) ? 1 : 0;
    k.K = k.by_image ? nimg * p->n_levels : p->n_levels;
    k.levels = w.levels; k.order = w.order; k.keys = w.keys; k.bh = w.bh; k.gstart = w.gstart;
    k.rflags = w.rflags;
    k.det_err = w.det_err;
    // R <= kPlanSingle: one launch (every plan warp ranks its RoI against all the others itself);
    // larger sets get their keys and per-block histograms from rpool_keys_kernel first
    k.n_blocks = 0;
    if (p->n_rois > kPlanSingle && o.order != RPOOL_SCHED_INPUT) {
        k.n_blocks = (int)ws_key_blocks(p->n_rois);
        rpool_keys_kernel<<<k.n_blocks, kKeyBlock, 0, st>>>(k);
        CUDA_TRY(cudaGetLastError(), "rpool_keys_kernel launch");
        g_launches++;
    }
    KParams kp;
    fill_params(p, w, o, false, kPlanThreads, kp);
    const dim3 plan_grid((p->n_rois + kPlanWarps - 1) / kPlanWarps);
    if (k.n_blocks > 0) {
        // behind rpool_keys_kernel: starts in its shadow, builds the tables, waits before ranking
        CUDA_TRY(launch_in_tail(rpool_plan_kernel, plan_grid, dim3(kPlanThreads), 0, st, kp, k, w.recs_fwd, w.recs_bwd),
                 "rpool_plan_kernel launch");
    } else {
        rpool_plan_kernel<<<plan_grid, kPlanThreads, 0, st>>>(kp, k, w.recs_fwd, w.recs_bwd);
        CUDA_TRY(cudaGetLastError(), "rpool_plan_kernel launch");
    }
    g_launches++;
    return RPOOL_OK;
}

int rpool_forward(const rpool_problem *p, void *ws, size_t ws_size, void *stream)
{
    int rc = validate(p, ws, ws_size, true);
    if (rc) return rc;
    Options o;
    rc = read_options(p, o);
    if (rc) return rc;
    if (p->n_rois == 0) return RPOOL_OK;
    KParams k;
    const int threads = o.threads;
    const int smem = fill_params(p, ws_split(ws, p), o, false, threads, k);
    if (smem > kMaxSmem)
        return fail(RPOOL_ERR_UNSUPPORTED, "forward needs %d bytes of shared memory; the limit is %d",
                    smem, kMaxSmem);
    rc = set_smem(rpool_forward_kernel, 0, smem);
    if (rc) return rc;
    // (always safe: the kernel's first instruction waits for everything queued before it; what the
    // early start buys is the launch ramp, hidden in the plan kernel's tail)
    const int grid = split_tail(rpool_forward_kernel, threads, smem, kTailPartsFwd, k);
    CUDA_TRY(launch_in_tail(rpool_forward_kernel, dim3(grid), dim3(threads), (size_t)smem,
                            static_cast<cudaStream_t>(stream), k), "rpool_forward_kernel launch");
    g_launches++;
    return RPOOL_OK;
}

// Launches the two prepass kernels of the deterministic backward (window
// rectangles, offsets of the private windows).
static int det_prepass(const rpool_problem *p, void *ws, const Options &o, cudaStream_t st, KParams &k,
                       int &threads)
{
    if (o.order != RPOOL_SCHED_DEFAULT)
        return fail(RPOOL_ERR_UNSUPPORTED, "deterministic backward needs the (image, level) schedule "
                    "(opt.schedule = RPOOL_SCHED_DEFAULT)");
    if (p->feat_layout != RPOOL_NHWC || p->pool_layout != RPOOL_NHWC || (p->channels & 3) || p->channels > 128 * kGatherSlabs)
        return fail(RPOOL_ERR_UNSUPPORTED, "deterministic backward needs channels-last tensors with "
                    "C %% 4 == 0 and C <= %d", 128 * kGatherSlabs);
    int nimg = 1;
    for (int l = 0; l < p->n_levels; ++l) nimg = p->level[l].n_images > nimg ? p->level[l].n_images : nimg;
    if (nimg * p->n_levels > kPlanMaxKeys)
        return fail(RPOOL_ERR_UNSUPPORTED, "deterministic backward: images x levels = %d exceeds %d",
                    nimg * p->n_levels, kPlanMaxKeys);
    const Workspace w = ws_split(ws, p);
    threads = o.threads;
    fill_params(p, w, o, true, threads, k);
    k.reverse = 0;
    if (p->n_rois > 0) {
        ScanParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.recs = w.recs_bwd;
        sp.rec_stride = w.rec_stride;
        sp.R = p->n_rois;
        sp.C = p->channels;
        sp.shapes_ok = o.force_path != kPathGeneric;
        for (int h = 0; h < p->n_heads; ++h)
            sp.shapes_ok = sp.shapes_ok && p->out_h[h] <= kPBwd && p->out_w[h] <= kPBwd;
        sp.rects = w.rects; sp.woff = w.woff; sp.total = w.det_total; sp.err = w.det_err;
        // (resets the error flag itself.)  Under opt.zero_fill_in_tail's contract -- the kernel queued
        // before this call neither touches the gradients nor produces the upstream gradients -- the scan
        // starts in that kernel's tail: it reads the plan's records only.
        if (o.fill_in_tail) {
            CUDA_TRY(launch_in_tail(rpool_det_scan_kernel, dim3(1), dim3(kScanThreads), 0, st, sp),
                     "rpool_det_scan_kernel launch");
        } else {
            rpool_det_scan_kernel<<<1, kScanThreads, 0, st>>>(sp);
            CUDA_TRY(cudaGetLastError(), "rpool_det_scan_kernel launch");
        }
        g_launches++;
    } else {
        CUDA_TRY(cudaMemsetAsync(w.det_err, 0, sizeof(int), st), "cudaMemsetAsync(det_err)");
        CUDA_TRY(cudaMemsetAsync(w.det_total, 0, sizeof(unsigned long long), st), "cudaMemsetAsync(det_total)");
    }
    return RPOOL_OK;
}

int rpool_backward_det_bytes(const rpool_problem *p, void *ws, size_t ws_size, void *stream,
                             size_t *bytes_out)
{
    int rc = validate(p, ws, ws_size, true);
    if (rc) return rc;
    Options o;
    rc = read_options(p, o);
    if (rc) return rc;
    if (!bytes_out) return fail(RPOOL_ERR_INVALID, "bytes_out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KParams k;
    int threads;
    rc = det_prepass(p, ws, o, st, k, threads);
    if (rc) return rc;
    const Workspace w = ws_split(ws, p);
    unsigned long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, w.det_total, sizeof(total), cudaMemcpyDeviceToHost, st), "copy det_total");
    CUDA_TRY(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    *bytes_out = (size_t)total * sizeof(float) + 16;
    return RPOOL_OK;
}

int rpool_status_flags(const void *ws, int32_t n_rois, void *stream, int32_t *flags_out)
{
    if (!ws || !flags_out || n_rois < 0) return fail(RPOOL_ERR_INVALID, "bad arguments");
    const Workspace w = ws_split(const_cast<void *>(ws), n_rois, 1, RPOOL_COORD_CAFFE2);   // fixed part only
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int flags = 0;
    if (n_rois > 0) {
        // the per-RoI flags of the plan and the deterministic pass's flag are folded on the
        // device into det_err's neighbour word, then one 8-byte copy
        rpool_flags_kernel<<<1, 256, 0, st>>>(w.rflags, n_rois, w.det_err, w.det_err + 1);
        CUDA_TRY(cudaGetLastError(), "rpool_flags_kernel launch");
        g_launches++;
        CUDA_TRY(cudaMemcpyAsync(&flags, w.det_err + 1, sizeof(int), cudaMemcpyDeviceToHost, st), "copy flags");
        CUDA_TRY(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    *flags_out = flags;
    return RPOOL_OK;
}

static int backward_det(const rpool_problem *p, void *ws, const Options &o, cudaStream_t st)
{
    if (!p->det_workspace) return fail(RPOOL_ERR_WORKSPACE, "deterministic backward: det_workspace is NULL "
                                       "(rpool_backward_det_bytes gives the exact size for these RoIs)");
    if (reinterpret_cast<uintptr_t>(p->det_workspace) & 15)
        return fail(RPOOL_ERR_INVALID, "det_workspace must be 16-byte aligned");
    KParams k;
    int threads;
    int rc = det_prepass(p, ws, o, st, k, threads);
    if (rc) return rc;
    const Workspace w = ws_split(ws, p);
    if (p->n_rois > 0) {
        int smem = fill_params(p, w, o, true, threads, k);
        while (smem > kMaxSmem && threads > 32) {
            threads -= 32;
            smem = fill_params(p, w, o, true, threads, k);
        }
        k.reverse = 0;
        k.det = 1;
        k.det_scratch = static_cast<float *>(p->det_workspace);
        k.det_scratch_floats = p->det_workspace_bytes / sizeof(float);
        rc = set_smem(rpool_backward_det_window_kernel, 3, smem);
        if (rc) return rc;
        // in the scan kernel's tail (waits before it reads the offsets)
        CUDA_TRY(launch_in_tail(rpool_backward_det_window_kernel, dim3(p->n_rois), dim3(threads), (size_t)smem, st, k),
                 "rpool_backward_det_window_kernel launch");
        g_launches++;
    }
    GatherParams g;
    memset(&g, 0, sizeof(g));
    long long ctas = 0;
    for (int li = 0; li < p->n_levels; ++li) {
        const int l = p->n_levels - 1 - li;       // launch order: coarse levels first
        g.lvl[l].data = static_cast<float *>(p->level[l].data);
        g.lvl[l].n_images = p->level[l].n_images;
        g.lvl[l].H = p->level[l].height;
        g.lvl[l].W = p->level[l].width;
        // at most 8 CTAs per map row: strips of 8, 16, 24 ... cells
        const int groups = (p->level[l].width + kGatherWarps - 1) / kGatherWarps;
        g.cells[l] = kGatherWarps * ((groups + kGatherRowCtas - 1) / kGatherRowCtas);
        g.strips[l] = (p->level[l].width + g.cells[l] - 1) / g.cells[l];
        g.strip_base[li] = ctas;
        ctas += (long long)p->level[l].n_images * p->level[l].height * g.strips[l];
    }
    g.strip_base[p->n_levels] = ctas;
    if (ctas > 2147483647ll) return fail(RPOOL_ERR_UNSUPPORTED, "pyramid too large for the gather launch");
    g.n_levels = p->n_levels;
    g.C = p->channels;
    g.accumulate = p->accumulate;
    g.gstart = w.gstart; g.rects = w.rects; g.woff = w.woff;
    g.scratch = static_cast<const float *>(p->det_workspace);
    g.scratch_floats = p->det_workspace_bytes / sizeof(float);
    if (p->n_rois == 0) {
        // no plan was made: every group is empty
        CUDA_TRY(cudaMemsetAsync(w.gstart, 0, kGstartInts * sizeof(int), st), "cudaMemsetAsync(gstart)");
    }
    // in the backward launch's tail when there is one: lists its windows, waits before it reads them
    const dim3 ggrid((unsigned)ctas), gblock(kGatherWarps * 32);
    if (p->n_rois > 0) {
        if (p->channels <= 256)
            CUDA_TRY(launch_in_tail(rpool_det_gather_kernel<2>, ggrid, gblock, 0, st, g), "rpool_det_gather_kernel launch");
        else
            CUDA_TRY(launch_in_tail(rpool_det_gather_kernel<kGatherSlabs>, ggrid, gblock, 0, st, g),
                     "rpool_det_gather_kernel launch");
    } else {
        if (p->channels <= 256) rpool_det_gather_kernel<2><<<ggrid, gblock, 0, st>>>(g);
        else rpool_det_gather_kernel<kGatherSlabs><<<ggrid, gblock, 0, st>>>(g);
        CUDA_TRY(cudaGetLastError(), "rpool_det_gather_kernel launch");
    }
    g_launches++;
    return RPOOL_OK;
}

// validation of the fields rpool_zero_fill reads
static int validate_levels(const rpool_problem *p)
{
    if (!p) return fail(RPOOL_ERR_INVALID, "problem is NULL");
    if (p->n_levels < 1 || p->n_levels > RPOOL_MAX_LEVELS)
        return fail(RPOOL_ERR_INVALID, "n_levels=%d outside [1,%d]", p->n_levels, RPOOL_MAX_LEVELS);
    if (p->channels < 1) return fail(RPOOL_ERR_INVALID, "channels=%d", p->channels);
    for (int l = 0; l < p->n_levels; ++l) {
        const rpool_level &L = p->level[l];
        if (!L.data || L.n_images < 1 || L.height < 1 || L.width < 1)
            return fail(RPOOL_ERR_INVALID, "level %d: data=%p n=%d h=%d w=%d", l, L.data,
                        L.n_images, L.height, L.width);
    }
    return RPOOL_OK;
}

static int launch_zero(const rpool_problem *p, cudaStream_t st, bool in_tail)
{
    ZeroParams z;
    memset(&z, 0, sizeof(z));
    z.n = p->n_levels;
    for (int l = 0; l < p->n_levels; ++l) {
        const unsigned long long n = (unsigned long long)p->level[l].n_images *
                                     p->level[l].height * p->level[l].width * p->channels;
        z.ptr[l] = static_cast<float *>(p->level[l].data);
        if (reinterpret_cast<uintptr_t>(z.ptr[l]) & 15) { z.n4[l] = 0; z.tail[l] = n; }
        else { z.n4[l] = n / 4; z.tail[l] = n % 4; }
        if (z.tail[l] > 1024ull * 148 * 8)
            return fail(RPOOL_ERR_UNSUPPORTED, "level %d gradient is not 16-byte aligned", l);
    }
    // 256-thread CTAs (18 registers): they fit beside the pooling CTAs that are still running when
    // the fill starts in the tail of the previous launch (opt.zero_fill_in_tail)
    if (in_tail) {
        CUDA_TRY(launch_in_tail(rpool_zero_kernel, dim3(148 * 32), dim3(256), 0, st, z), "rpool_zero_kernel launch");
    } else {
        rpool_zero_kernel<<<148 * 32, 256, 0, st>>>(z);
        CUDA_TRY(cudaGetLastError(), "rpool_zero_kernel launch");
    }
    g_launches++;
    return RPOOL_OK;
}

int rpool_zero_fill(const rpool_problem *p, void *stream)
{
    int rc = validate_levels(p);
    if (rc) return rc;
    if (p->opt.zero_fill_in_tail < 0 || p->opt.zero_fill_in_tail > 1)
        return fail(RPOOL_ERR_INVALID, "opt.zero_fill_in_tail=%d outside [0,1]", p->opt.zero_fill_in_tail);
    return launch_zero(p, static_cast<cudaStream_t>(stream), p->opt.zero_fill_in_tail != 0);
}

int rpool_backward(const rpool_problem *p, void *ws, size_t ws_size, void *stream)
{
    int rc = validate(p, ws, ws_size, true);
    if (rc) return rc;
    Options o;
    rc = read_options(p, o);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->deterministic) return backward_det(p, ws, o, st);
    if (!p->accumulate) {
        rc = launch_zero(p, st, o.fill_in_tail != 0);
        if (rc) return rc;
    }
    if (p->n_rois == 0) return RPOOL_OK;
    const Workspace w = ws_split(ws, p);
    // Two pooled sizes: the backward pass has nothing to share between them (each reads its
    // own gy; the reductions into the gradient are per size either way), and one launch per
    // size keeps the per-CTA strips small (more L1 for the gy re-reads): 2 launches unless
    // opt.fuse_heads_backward is set.
    const int parts = (p->n_heads > 1 && o.split_heads) ? p->n_heads : 1;
    // (launch order: the pooled size with the most bins first -- its CTAs are the long ones; the launch
    // that drains with nothing queued behind it is then the one with the short CTAs)
    int head_order[RPOOL_MAX_HEADS];
    for (int h = 0; h < RPOOL_MAX_HEADS; ++h) head_order[h] = h;
    if (parts > 1)
        for (int a = 0; a < parts; ++a)
            for (int b = a + 1; b < parts; ++b)
                if (p->out_h[head_order[b]] * p->out_w[head_order[b]] > p->out_h[head_order[a]] * p->out_w[head_order[a]]) {
                    const int t = head_order[a]; head_order[a] = head_order[b]; head_order[b] = t;
                }
    for (int pi = 0; pi < parts; ++pi) {
        const int part = parts > 1 ? head_order[pi] : 0;
        rpool_problem q = *p;
        if (parts > 1) {
            q.n_heads = 1;
            q.out_h[0] = p->out_h[part];
            q.out_w[0] = p->out_w[part];
            q.pooled[0] = p->pooled[part];
        }
        KParams k;
        int threads = o.threads;
        int smem = fill_params(&q, w, o, true, threads, k);
        while (smem > kMaxSmem && threads > 32) {  // two wide heads: fewer warps, same result
            threads -= 32;
            smem = fill_params(&q, w, o, true, threads, k);
        }
        k.rec_head = parts > 1 ? part : 0;
        k.rec_stride = w.rec_stride;       // records keep the layout of the plan's head count
        rc = set_smem(rpool_backward_kernel, 1, smem);
        if (rc) return rc;
        // the launch for the second pooled size depends on the zero fill only, like the first one:
        // it may start in the first one's tail (both add into the gradients with reductions)
        // ... and the first one, when this call's own zero fill is the kernel before it, may start in the
        // fill's tail: it waits before its first reduction (what it reads earlier -- the plan, gy --
        // was complete before the fill started, or is covered by opt.zero_fill_in_tail's contract)
        k.wait_fill = (pi == 0 && !p->accumulate) ? 1 : 0;
        int st_smem = 0;
        const int st_grid = staged_grid(&q, o, st_smem);
        if (st_grid > 0) {
            k.staged_slots = (p->n_rois + st_grid - 1) / st_grid;
            rc = set_smem(rpool_backward_staged_kernel, 2, st_smem);
            if (rc) return rc;
            if (pi > 0 || !p->accumulate) {
                CUDA_TRY(launch_in_tail(rpool_backward_staged_kernel, dim3(st_grid), dim3(kStThreads), (size_t)st_smem, st, k),
                         "rpool_backward_staged_kernel launch");
            } else {
                rpool_backward_staged_kernel<<<st_grid, kStThreads, st_smem, st>>>(k);
                CUDA_TRY(cudaGetLastError(), "rpool_backward_staged_kernel launch");
            }
            g_launches++;
            continue;
        }
        const int grid = split_tail(rpool_backward_kernel, threads, smem, kTailPartsBwd, k);
        if (pi > 0 || !p->accumulate) {
            CUDA_TRY(launch_in_tail(rpool_backward_kernel, dim3(grid), dim3(threads), (size_t)smem, st, k),
                     "rpool_backward_kernel launch");
        } else {
            rpool_backward_kernel<<<grid, threads, smem, st>>>(k);
            CUDA_TRY(cudaGetLastError(), "rpool_backward_kernel launch");
        }
        g_launches++;
    }
    return RPOOL_OK;
}

int rpool_read_plan(const void *ws, int32_t n_rois, int32_t *levels_host, int32_t *order_host,
                    void *stream)
{
    if (!ws || n_rois < 0) return fail(RPOOL_ERR_INVALID, "bad arguments");
    const Workspace w = ws_split(const_cast<void *>(ws), n_rois, 1, RPOOL_COORD_CAFFE2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (levels_host)
        CUDA_TRY(cudaMemcpyAsync(levels_host, w.levels, sizeof(int) * n_rois, cudaMemcpyDeviceToHost, st),
                 "copy levels");
    if (order_host)
        CUDA_TRY(cudaMemcpyAsync(order_host, w.order, sizeof(int) * n_rois, cudaMemcpyDeviceToHost, st),
                 "copy order");
    CUDA_TRY(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return RPOOL_OK;
}

static int transpose(const float *src, float *dst, int n, int rows, int cols, void *stream)
{
    if (!src || !dst || n < 1 || rows < 1 || cols < 1) return fail(RPOOL_ERR_INVALID, "bad arguments");
    if (n > 65535) return fail(RPOOL_ERR_UNSUPPORTED, "more than 65535 images");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, n), block(32, 8);
    if (grid.y > 65535) return fail(RPOOL_ERR_UNSUPPORTED, "plane too large");
    rpool_transpose_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
    CUDA_TRY(cudaGetLastError(), "rpool_transpose_kernel launch");
    g_launches++;
    return RPOOL_OK;
}

int rpool_nchw_to_nhwc(const float *src, float *dst, int32_t n, int32_t c, int32_t h, int32_t w,
                       void *stream)
{
    return transpose(src, dst, n, c, h * w, stream);  // (C, HW) -> (HW, C)
}

int rpool_nhwc_to_nchw(const float *src, float *dst, int32_t n, int32_t c, int32_t h, int32_t w,
                       void *stream)
{
    return transpose(src, dst, n, h * w, c, stream);  // (HW, C) -> (C, HW)
}

}  // extern "C"
