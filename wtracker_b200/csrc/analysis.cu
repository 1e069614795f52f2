// Derived columns of the analysed tracking log: DataAnalyzer.initialize,
// wtracker/eval/data_analyzer.py:54-107 — box centres, n-lag speed, worm-to-microscope deviation, bbox error and the
// final DataFrame.round(5) — one thread per log row, float64 with numpy's operation order (no FMA contraction).
//   in : f64 [n][17]  the bboxes.csv columns (wt_log_rows layout; column 2 = phase is not used)
//   out: f64 [n][30]  frame, cycle, plt_x, plt_y, cam_x..cam_h, mic_x..mic_h, wrm_x..wrm_h, time, cycle_step,
//                     wrm_center_x/y, mic_center_x/y, wrm_speed_x/y, wrm_speed, worm_deviation_x/y, worm_deviation,
//                     bbox_error, precise_error (NaN until calculate_precise fills it)
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

__device__ __forceinline__ double np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
// numpy.round(x, 5): rint(x * 10^5) / 10^5
__device__ __forceinline__ double round5(double x) { return __ddiv_rn(rint(__dmul_rn(x, 100000.0)), 100000.0); }

__global__ void analysis_columns_kernel(const double* __restrict__ in, long long n, int period, int cycle_frame_num,
                                        double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = in + i * 17;
    double* o = out + i * 30;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double wx = r[13], wy = r[14], ww = r[15], wh = r[16];
    const double mx = r[9], my = r[10], mw = r[11], mh = r[12];
    const double wcx = __dadd_rn(wx, __ddiv_rn(ww, 2.0)), wcy = __dadd_rn(wy, __ddiv_rn(wh, 2.0));
    const double mcx = __dadd_rn(mx, __ddiv_rn(mw, 2.0)), mcy = __dadd_rn(my, __ddiv_rn(mh, 2.0));
    double sx = nan, sy = nan;
    if (i >= period) {
        const double* p = in + (i - period) * 17;
        const double pcx = __dadd_rn(p[13], __ddiv_rn(p[15], 2.0)), pcy = __dadd_rn(p[14], __ddiv_rn(p[16], 2.0));
        const double dt = __dsub_rn(r[0], p[0]);
        sx = __ddiv_rn(__dsub_rn(wcx, pcx), dt);
        sy = __ddiv_rn(__dsub_rn(wcy, pcy), dt);
    }
    const double speed = __dsqrt_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)));
    const double dx = __dsub_rn(wcx, mcx), dy = __dsub_rn(wcy, mcy);
    const double dev = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    // ErrorCalculator.calculate_bbox_error (error_calculator.py:163-195)
    const double wr = __dadd_rn(wx, ww), wb = __dadd_rn(wy, wh), mr = __dadd_rn(mx, mw), mb = __dadd_rn(my, mh);
    const double il = np_max(wx, mx), it = np_max(wy, my), ir = np_min(wr, mr), ib = np_min(wb, mb);
    const double iw = np_max(0.0, __dsub_rn(ir, il)), ih = np_max(0.0, __dsub_rn(ib, it));
    const double total = __dmul_rn(ww, wh);
    double err = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(iw, ih), total));
    if (total == 0.0) err = 0.0;
    o[0] = r[0]; o[1] = r[1];
#pragma unroll
    for (int k = 0; k < 10; ++k) o[2 + k] = r[3 + k];                 // plt, cam, mic: integers, round is a no-op
    o[12] = round5(wx); o[13] = round5(wy); o[14] = round5(ww); o[15] = round5(wh);
    o[16] = r[0];                                                      // time = frame
    o[17] = double((long long)r[0] % cycle_frame_num);                 // cycle_step
    o[18] = round5(wcx); o[19] = round5(wcy); o[20] = round5(mcx); o[21] = round5(mcy);
    o[22] = round5(sx); o[23] = round5(sy); o[24] = round5(speed);
    o[25] = round5(dx); o[26] = round5(dy); o[27] = round5(dev);
    o[28] = round5(err);
    o[29] = nan;
}

// Row masks of DataAnalyzer.clean (keep) and DataAnalyzer.calc_anomalies (six flag bits) over the analysed table.
struct MaskParams {
    int imaging_only, has_bounds, no_preds;
    double b0, b1, b2, b3;
    double min_bbox_error, min_dist_error, min_speed, min_size;
};

__global__ void analysis_masks_kernel(const double* __restrict__ table30, const uint8_t* __restrict__ moving, long long n,
                                      const MaskParams m, uint8_t* __restrict__ keep, uint8_t* __restrict__ anomaly) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = table30 + i * 30;
    const double mx = r[8], my = r[9], mw = r[10], mh = r[11];
    const double wx = r[12], wy = r[13], ww = r[14], wh = r[15];
    const bool has_pred = isfinite(wx) && isfinite(wy) && isfinite(ww) && isfinite(wh);
    bool k = true;
    if (m.imaging_only && moving[i]) k = false;
    if (m.has_bounds) {
        // The reference means "a frame with a prediction is judged by its worm box, a frame without one by the
        // microscope box", but `mask_wrm = has_pred; mask_wrm &= <x condition>` narrows has_pred IN PLACE (one ndarray,
        // two names; the result comes back as a pandas Series, so the second `&=` with the y condition no longer
        // touches has_pred — data_analyzer.py:140-148 under numpy 2 / pandas 2).  Hence `mask_mic = ~has_pred` means
        // "no prediction, or the worm box fails the X range": such rows are judged by the microscope box.  Kept as is
        // (results must be the reference's; golden: tests/golden/reference_masks.npz).  NaN comparisons are false.
        const bool in_wx = has_pred && wx >= m.b0 && __dadd_rn(wx, ww) <= m.b2;
        const bool in_wrm = in_wx && wy >= m.b1 && __dadd_rn(wy, wh) <= m.b3;
        const bool in_mic = !in_wx && mx >= m.b0 && __dadd_rn(mx, mw) <= m.b2 && my >= m.b1 && __dadd_rn(my, mh) <= m.b3;
        k = k && (in_wrm || in_mic);
    }
    keep[i] = k ? 1 : 0;
    uint8_t a = 0;
    if (r[24] >= m.min_speed) a |= 1;         // wrm_speed
    if (r[28] >= m.min_bbox_error) a |= 2;    // bbox_error
    if (r[27] >= m.min_dist_error) a |= 4;    // worm_deviation
    if (ww >= m.min_size) a |= 8;
    if (wh >= m.min_size) a |= 16;
    if (m.no_preds && !has_pred) a |= 32;
    anomaly[i] = a;
}

}  // namespace
}  // namespace wt

extern "C" int wt_analysis_masks(const double* table30, const uint8_t* moving, int64_t n, int imaging_only,
                                 const double* h_bounds, int no_preds, double min_bbox_error, double min_dist_error,
                                 double min_speed, double min_size, uint8_t* keep, uint8_t* anomaly, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(table30 && keep && anomaly, "null argument");
    WT_REQUIRE(!imaging_only || moving, "imaging_only needs the phase column");
    MaskParams m;
    m.imaging_only = imaging_only;
    m.has_bounds = h_bounds != nullptr;
    m.no_preds = no_preds;
    m.b0 = m.b1 = m.b2 = m.b3 = 0.0;
    if (h_bounds) { m.b0 = h_bounds[0]; m.b1 = h_bounds[1]; m.b2 = h_bounds[2]; m.b3 = h_bounds[3]; }
    m.min_bbox_error = min_bbox_error;
    m.min_dist_error = min_dist_error;
    m.min_speed = min_speed;
    m.min_size = min_size;
    analysis_masks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(table30, moving, n, m,
                                                                                                 keep, anomaly);
    WT_LAUNCHED();
    return 0;
}

extern "C" int wt_analysis_columns(const double* table, int64_t n, int period, int cycle_frame_num, double* out,
                                   void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(table && out, "null argument");
    WT_REQUIRE(period >= 1 && cycle_frame_num >= 1, "period and cycle length must be positive");
    analysis_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, n, period,
                                                                                                   cycle_frame_num, out);
    WT_LAUNCHED();
    return 0;
}
