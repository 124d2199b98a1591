// oracle/ref_caffe2_wrapper.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Exposes the reference's own C++ forward (ROIAlignForward<float>,
// chainer_maskrcnn/functions/roi_align/caffe2_operation/caffe2_roi_align.cpp:115-226)
// with its sampling_ratio argument reachable, which the reference's pybind11
// binding hard-codes to 1 (caffe2_roi_align.cpp:240).  The reference source is
// NOT copied: it is #included from where it lies under /root/reference at
// build time (REF_CAFFE2_CPP is set by oracle/Makefile), and the result goes
// to oracle/_ref/ only.
#include <cmath>
#include <vector>
#include <algorithm>
#include <stdexcept>
using std::ceil;
#include REF_CAFFE2_CPP

extern "C" __attribute__((visibility("default")))
int ref_caffe2_roi_align_forward(const float* bottom_data, int N, int C, int H, int W,
                                 const float* bottom_rois, int R, int out_h, int out_w,
                                 float spatial_scale, int sampling_ratio, float* top_data)
{
    (void)N;
    try {
        const int nthreads = R * C * out_h * out_w;
        ROIAlignForward<float>(nthreads, bottom_data, spatial_scale, C, H, W, out_h, out_w,
                               sampling_ratio, bottom_rois, 5, top_data);
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}
