// oracle/refgpu_baseline.cu -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Same-box GPU baseline: the arithmetic of the reference's
// two CuPy ElementwiseKernels restated as plain grid-stride CUDA kernels, driven
// the way the reference's FPN heads drive them.  CuPy is not installed in this
// image, so the reference's own GPU path cannot run; this file restates
//   * roi_align_2d_fwd  (chainer_maskrcnn/functions/roi_align/roi_align_2d.py:100-144):
//       one thread per output element (r, c, ph, pw), NCHW, four scalar loads;
//   * roi_align_2d_bwd  (roi_align_2d.py:196-279): one thread per *input* element
//       (n, c, h, w) looping over every RoI and every bin (gather formulation;
//       coincident-cell taps are dropped by its if/continue chain, :256-272);
//   * the heads' dispatch (model/head/fpn_roi_mask_head.py:57-63,74-78): one op
//       call per RoI on x[level[r]], i.e. per RoI one forward launch, and in
//       backward per RoI a zero-filled dense map (cupy.zeros, :195), one kernel
//       over the whole level map and Chainer's accumulation of that dense
//       gradient into the level's gradient.
// Only bench.py (its "gpu_baseline" object) and tests/ use it.  The dispatch loop
// here is C (launch cost only), which flatters the reference: its loop is Python.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void ref_fwd_kernel(const float *__restrict__ x, float scale, int C, int H, int W, int PH,
                               int PW, const float *__restrict__ rois, float *__restrict__ top, long long total)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int pw = (int)(i % PW);
        const int ph = (int)((i / PW) % PH);
        const int c = (int)((i / PW / PH) % C);
        const int r = (int)(i / PW / PH / C);
        const float *q = rois + (size_t)r * 5;
        const int b = (int)q[0];
        const float sw = q[1] * scale, sh = q[2] * scale, ew = q[3] * scale, eh = q[4] * scale;
        const float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
        const float bh = rh / (float)PH, bw = rw / (float)PW;
        // (ph + 0.5) is a double in the reference's kernel source (:127-128)
        const float cy = (float)((ph + 0.5) * bh + sh);
        const float cx = (float)((pw + 0.5) * bw + sw);
        const float p = cy - floorf(cy), qq = cx - floorf(cx);
        const int y1 = max((int)floorf(cy), 0), x1 = max((int)floorf(cx), 0);
        const int y2 = min(y1 + 1, H - 1), x2 = min(x1 + 1, W - 1);
        const float *f = x + ((size_t)b * C + c) * H * W;
        float v = f[y1 * W + x1] * (1 - p) * (1 - qq);
        v += f[y2 * W + x1] * p * (1 - qq);
        v += f[y1 * W + x2] * (1 - p) * qq;
        v += f[y2 * W + x2] * p * qq;
        top[i] = v;
    }
}

__global__ void ref_bwd_kernel(const float *__restrict__ gy, int R, float scale, int C, int H, int W,
                               int PH, int PW, const float *__restrict__ rois, float *__restrict__ gx,
                               long long total)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const int c = (int)((i / ((long long)W * H)) % C);
        const int n = (int)(i / ((long long)W * H * C));
        float g = 0.f;
        for (int r = 0; r < R; ++r) {
            const float *q = rois + (size_t)r * 5;
            if (n != (int)q[0]) continue;
            const float sw = q[1] * scale, sh = q[2] * scale, ew = q[3] * scale, eh = q[4] * scale;
            const float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
            if (!(sw - 2 <= w && w <= ew + 2 && sh - 2 <= h && h <= eh + 2)) continue;
            const float bh = rh / (float)PH, bw = rw / (float)PW;
            const float *t = gy + ((size_t)r * C + c) * PH * PW;
            for (int row = 0; row < PH; ++row) {
                for (int col = 0; col < PW; ++col) {
                    const float cx = (col + 0.5f) * bw + sw, cy = (row + 0.5f) * bh + sh;
                    const float p = cy - floorf(cy), qq = cx - floorf(cx);
                    const int x0 = max(min((int)cx, W - 1), 0), y0 = max(min((int)cy, H - 1), 0);
                    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
                    const float tv = t[row * PW + col];
                    if (x0 == w && y0 == h) { g += (1 - p) * (1 - qq) * tv; continue; }
                    if (x1 == w && y0 == h) { g += (1 - p) * qq * tv; continue; }
                    if (x0 == w && y1 == h) { g += p * (1 - qq) * tv; continue; }
                    if (x1 == w && y1 == h) { g += p * qq * tv; }
                }
            }
        }
        gx[i] = g;
    }
}

__global__ void ref_add_kernel(float *__restrict__ acc, const float *__restrict__ g, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        acc[i] += g[i];
}

int grid_for(long long total)
{
    long long g = (total + 255) / 256;
    return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace

#define REFGPU_API extern "C" __attribute__((visibility("default")))

// One batched call of the restated forward kernel (the reference API allows this
// for a single map: light_roi_mask_head.py:91-92).
REFGPU_API int refgpu_forward(const float *x, int C, int H, int W, const float *rois_xy, int R, int PH,
                              int PW, float scale, float *top, void *stream)
{
    const long long total = (long long)R * C * PH * PW;
    if (total == 0) return 0;
    ref_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, scale, C, H, W, PH, PW, rois_xy,
                                                                      top, total);
    return (int)cudaGetLastError();
}

REFGPU_API int refgpu_backward(const float *gy, const float *rois_xy, int R, int N, int C, int H, int W,
                               int PH, int PW, float scale, float *gx, void *stream)
{
    const long long total = (long long)N * C * H * W;
    ref_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(gy, R, scale, C, H, W, PH, PW, rois_xy,
                                                                      gx, total);
    return (int)cudaGetLastError();
}

struct refgpu_level {
    const float *x;   // (N, C, H, W) features
    float *grad;      // (N, C, H, W) accumulated gradient of this level
    float *tmp;       // (N, C, H, W) the per-call dense gradient
    int N, H, W;
    float scale;
};

// The heads' per-RoI dispatch over a pyramid: levels_host[r] picks the map.
// top (R, C, P, P) and gy (R, C, P, P) are NCHW like the reference's arrays.
// do_backward: 0 forward only.  Returns the number of kernel launches + memsets.
REFGPU_API long long refgpu_fpn_step(const refgpu_level *lv, int n_levels, int C, const float *rois_xy,
                                     const int *levels_host, int R, int P, float *top, const float *gy,
                                     int do_backward, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    long long ops = 0;
    const long long per_roi = (long long)C * P * P;
    for (int r = 0; r < R; ++r) {
        const refgpu_level &L = lv[levels_host[r]];
        ref_fwd_kernel<<<grid_for(per_roi), 256, 0, st>>>(L.x, L.scale, C, L.H, L.W, P, P, rois_xy + 5 * (size_t)r,
                                                          top + (size_t)r * per_roi, per_roi);
        ++ops;
    }
    if (do_backward) {
        bool seen[16] = {false};
        for (int r = R - 1; r >= 0; --r) {   // backprop visits the concat's inputs last to first
            const int l = levels_host[r];
            const refgpu_level &L = lv[l];
            const long long n = (long long)L.N * C * L.H * L.W;
            cudaMemsetAsync(L.tmp, 0, (size_t)n * 4, st);                       // cupy.zeros (:195)
            ref_bwd_kernel<<<grid_for(n), 256, 0, st>>>(gy + (size_t)r * per_roi, 1, L.scale, C, L.H, L.W, P, P,
                                                        rois_xy + 5 * (size_t)r, L.tmp, n);
            if (!seen[l]) {                                                     // first gradient: kept as is
                cudaMemcpyAsync(L.grad, L.tmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
                seen[l] = true;
            } else {                                                            // later ones: accumulated
                ref_add_kernel<<<grid_for(n), 256, 0, st>>>(L.grad, L.tmp, n);
            }
            ops += 3;
        }
        for (int l = 0; l < n_levels; ++l)
            if (!seen[l]) {
                cudaMemsetAsync(lv[l].grad, 0, (size_t)lv[l].N * C * lv[l].H * lv[l].W * 4, st);
                ++ops;
            }
    }
    return cudaGetLastError() == cudaSuccess ? ops : -1;
}
