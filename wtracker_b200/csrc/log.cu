// Rows of the tracking log `bboxes.csv`: the per-frame part of
// LoggingController._log_cycle (wtracker/sim/sim_controllers/logging_controller.py:145-185) and the crop
// dimensions of BoxUtils.discretize (wtracker/utils/bbox_utils.py:119-167), one thread per frame.
//
//   wrm = camera-relative worm box + camera-box origin (absolute frame coordinates)     :152-154
//   a row with a non-finite coordinate becomes 0, 0, 0, 0 — discretize() zeroes it in the caller's array
//   BEFORE the csv rows are written (bbox_utils.py:141), so that is what the reference logs
//   crop = (floor x1, floor y1, ceil x2, ceil y2) clipped to the frame, back to xywh, 0 if empty
//
// The arithmetic follows the dtype of the controller's prediction array: float64 (CsvController, or any
// YoloController cycle with a missed detection) or float32 (yolo_controller.py:85-90), because numpy adds the
// int64 camera origin in float64 and casts the result back to the array's dtype.
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

constexpr int kLogCols = 17;

struct LogParams {
    const void* worm_rel;
    int worm_is_f32;
    const int32_t* cam;
    const int32_t* mic;
    const int32_t* plt;
    long long n, first_frame;
    int cycle_frame_num, imaging_frame_num, frame_h, frame_w;
    double* table;
    int32_t* crop;
    uint8_t* legal;
};

__device__ __forceinline__ int clip_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void log_rows_kernel(const LogParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int4 cam = reinterpret_cast<const int4*>(p.cam)[i];
    const int4 mic = reinterpret_cast<const int4*>(p.mic)[i];
    const int2 plt = reinterpret_cast<const int2*>(p.plt)[i];
    double x, y, w, h, x2, y2;
    if (p.worm_is_f32) {
        const float4 r = reinterpret_cast<const float4*>(p.worm_rel)[i];
        // float32 array += int64 column: computed in float64, cast back to float32
        float fx = float(__dadd_rn(double(r.x), double(cam.x))), fy = float(__dadd_rn(double(r.y), double(cam.y)));
        float fw = r.z, fh = r.w;
        if (!(isfinite(fx) && isfinite(fy) && isfinite(fw) && isfinite(fh))) fx = fy = fw = fh = 0.f;
        x = fx; y = fy; w = fw; h = fh;
        x2 = double(__fadd_rn(fx, fw));
        y2 = double(__fadd_rn(fy, fh));
    } else {
        const double2 r01 = reinterpret_cast<const double2*>(p.worm_rel)[2 * i];
        const double2 r23 = reinterpret_cast<const double2*>(p.worm_rel)[2 * i + 1];
        x = __dadd_rn(r01.x, double(cam.x));
        y = __dadd_rn(r01.y, double(cam.y));
        w = r23.x;
        h = r23.y;
        if (!(isfinite(x) && isfinite(y) && isfinite(w) && isfinite(h))) x = y = w = h = 0.0;
        x2 = __dadd_rn(x, w);
        y2 = __dadd_rn(y, h);
    }
    const long long frame = p.first_frame + i;
    double* row = p.table + i * kLogCols;
    row[0] = double(frame);
    row[1] = double(frame / p.cycle_frame_num);
    row[2] = (frame % p.cycle_frame_num) < p.imaging_frame_num ? 0.0 : 1.0;   // 0 = "imaging", 1 = "moving"
    row[3] = plt.x; row[4] = plt.y;
    row[5] = cam.x; row[6] = cam.y; row[7] = cam.z; row[8] = cam.w;
    row[9] = mic.x; row[10] = mic.y; row[11] = mic.z; row[12] = mic.w;
    row[13] = x; row[14] = y; row[15] = w; row[16] = h;

    // BoxUtils.discretize(XYWH): floor / ceil to int32, clip to [0, W] x [0, H], empty boxes are illegal
    int ix1 = clip_i(int(floor(x)), 0, p.frame_w), iy1 = clip_i(int(floor(y)), 0, p.frame_h);
    int ix2 = clip_i(int(ceil(x2)), 0, p.frame_w), iy2 = clip_i(int(ceil(y2)), 0, p.frame_h);
    const bool ok = (ix2 - ix1) > 0 && (iy2 - iy1) > 0;
    reinterpret_cast<int4*>(p.crop)[i] = ok ? make_int4(ix1, iy1, ix2 - ix1, iy2 - iy1) : make_int4(0, 0, 0, 0);
    p.legal[i] = ok ? 1 : 0;
}

}  // namespace
}  // namespace wt

extern "C" int wt_log_rows(const void* worm_rel, int worm_is_f32, const int32_t* cam_xywh, const int32_t* mic_xywh,
                           const int32_t* plt_xy, int64_t n, int64_t first_frame, int cycle_frame_num,
                           int imaging_frame_num, int frame_h, int frame_w, double* table, int32_t* crop_xywh,
                           uint8_t* crop_legal, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(worm_rel && cam_xywh && mic_xywh && plt_xy && table && crop_xywh && crop_legal, "null argument");
    WT_REQUIRE(cycle_frame_num >= 1 && imaging_frame_num >= 0 && first_frame >= 0, "cycle geometry");
    auto aligned16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    WT_REQUIRE(aligned16(worm_rel) && aligned16(cam_xywh) && aligned16(mic_xywh) && aligned16(crop_xywh) &&
                   reinterpret_cast<uintptr_t>(plt_xy) % 8 == 0 && reinterpret_cast<uintptr_t>(table) % 8 == 0,
               "wt_log_rows: box arrays must be 16-byte aligned (rows are read / written as 128-bit words)");
    LogParams p;
    p.worm_rel = worm_rel;
    p.worm_is_f32 = worm_is_f32;
    p.cam = cam_xywh;
    p.mic = mic_xywh;
    p.plt = plt_xy;
    p.n = n;
    p.first_frame = first_frame;
    p.cycle_frame_num = cycle_frame_num;
    p.imaging_frame_num = imaging_frame_num;
    p.frame_h = frame_h;
    p.frame_w = frame_w;
    p.table = table;
    p.crop = crop_xywh;
    p.legal = crop_legal;
    log_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
