// Segmentation-based tracking error: ErrorCalculator.calculate_precise /
// calculate_segmentation, wtracker/eval/error_calculator.py:19-161.  One warp per row of the log:
//   discretize the worm and microscope boxes (BoxUtils.discretize, bbox_utils.py:119-167), intersect them,
//   mask = |view - background| > diff_thresh over the worm box (u8 grey), error = 1 - |mask ∩ mic| / |mask|
//   (0 when the mask is empty, NaN when the worm box is illegal).
// The worm view of a row is its frame cropped at the discretized worm box (what the reference saved as
// wrm_*.png and reads back through `worm_reader`), so nothing but the frames and the background is read: two
// bytes per pixel of the worm box.  float64 box arithmetic with explicit round-to-nearest, integer pixel counts.
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

struct IBox {
    int x, y, w, h;
    bool legal;
};

__device__ __forceinline__ int clip_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ IBox discretize_xywh(double x, double y, double w, double h, int H, int W) {
    if (!(isfinite(x) && isfinite(y) && isfinite(w) && isfinite(h))) x = y = w = h = 0.0;
    const int x1 = clip_i(int(floor(x)), 0, W), y1 = clip_i(int(floor(y)), 0, H);
    const int x2 = clip_i(int(ceil(__dadd_rn(x, w))), 0, W), y2 = clip_i(int(ceil(__dadd_rn(y, h))), 0, H);
    IBox b;
    b.legal = (x2 - x1) > 0 && (y2 - y1) > 0;
    b.x = b.legal ? x1 : 0;
    b.y = b.legal ? y1 : 0;
    b.w = b.legal ? x2 - x1 : 0;
    b.h = b.legal ? y2 - y1 : 0;
    return b;
}

__global__ void __launch_bounds__(256) precise_error_kernel(const uint8_t* __restrict__ frames, int fh, int fw,
                                                            const int32_t* __restrict__ frame_idx,
                                                            const long long* __restrict__ view_off,
                                                            const uint8_t* __restrict__ background,
                                                            const double* __restrict__ worm, const double* __restrict__ mic,
                                                            double thr, double* __restrict__ err, long long n) {
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double2 w01 = reinterpret_cast<const double2*>(worm)[2 * i], w23 = reinterpret_cast<const double2*>(worm)[2 * i + 1];
    const double2 m01 = reinterpret_cast<const double2*>(mic)[2 * i], m23 = reinterpret_cast<const double2*>(mic)[2 * i + 1];
    const IBox wb = discretize_xywh(w01.x, w01.y, w23.x, w23.y, fh, fw);
    const IBox mb = discretize_xywh(m01.x, m01.y, m23.x, m23.y, fh, fw);
    if (!wb.legal) {
        if (lane == 0) err[i] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    // intersection of the two boxes in worm-view coordinates (error_calculator.py:110-124)
    const int il = max(wb.x, mb.x), it = max(wb.y, mb.y);
    const int ir = min(wb.x + wb.w, mb.x + mb.w), ib = min(wb.y + wb.h, mb.y + mb.h);
    const int iw = max(0, ir - il), ih = max(0, ib - it);
    const int ix = il - wb.x, iy = it - wb.y;
    // the worm view: the frame cropped at the worm box, or (view_off) row i's own crop in a packed buffer
    const uint8_t* img = view_off ? frames + view_off[i] : frames + size_t(frame_idx[i]) * fh * fw + size_t(wb.y) * fw + wb.x;
    const int pitch = view_off ? wb.w : fw;
    const uint8_t* bg = background + size_t(wb.y) * fw + wb.x;
    int total = 0, inside = 0;
    const int npix = wb.w * wb.h;
    for (int p = lane; p < npix; p += 32) {
        const int r = p / wb.w, c = p - r * wb.w;
        const int d = abs(int(__ldg(img + size_t(r) * pitch + c)) - int(__ldg(bg + size_t(r) * fw + c)));
        if (double(d) > thr) {
            ++total;
            inside += (r >= iy && r < iy + ih && c >= ix && c < ix + iw) ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        total += __shfl_xor_sync(0xffffffffu, total, o);
        inside += __shfl_xor_sync(0xffffffffu, inside, o);
    }
    if (lane == 0) err[i] = total == 0 ? 0.0 : __dsub_rn(1.0, __ddiv_rn(double(inside), double(total)));
}

}  // namespace
}  // namespace wt

extern "C" int wt_precise_error(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* frame_idx,
                                const int64_t* view_off, const uint8_t* background, const double* worm_xywh, const double* mic_xywh,
                                double diff_thresh, double* err, int64_t n, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(frames && (frame_idx || view_off) && background && worm_xywh && mic_xywh && err, "null argument");
    WT_REQUIRE((view_off || n_frames >= 1) && frame_h >= 1 && frame_w >= 1, "frame geometry");
    const long long threads = n * 32;
    precise_error_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        frames, frame_h, frame_w, frame_idx, reinterpret_cast<const long long*>(view_off), background, worm_xywh, mic_xywh,
        diff_thresh, err, n);
    WT_LAUNCHED();
    return 0;
}
