// K10: per-row bbox metrics in float64, matching numpy's IEEE arithmetic op for op
// (wtracker/eval/error_calculator.py:163-195 and 197-212).  NaN rows propagate like np.maximum /
// np.minimum do; no FMA contraction (explicit round-to-nearest intrinsics).
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

__device__ __forceinline__ double np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

__global__ void bbox_error_kernel(const double* __restrict__ wrm, const double* __restrict__ mic,
                                  double* __restrict__ err, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 w01 = reinterpret_cast<const double2*>(wrm)[2 * i], w23 = reinterpret_cast<const double2*>(wrm)[2 * i + 1];
    const double2 m01 = reinterpret_cast<const double2*>(mic)[2 * i], m23 = reinterpret_cast<const double2*>(mic)[2 * i + 1];
    const double wr = __dadd_rn(w01.x, w23.x), wb = __dadd_rn(w01.y, w23.y);
    const double mr = __dadd_rn(m01.x, m23.x), mb = __dadd_rn(m01.y, m23.y);
    const double il = np_max(w01.x, m01.x), it = np_max(w01.y, m01.y);
    const double ir = np_min(wr, mr), ib = np_min(wb, mb);
    const double iw = np_max(0.0, __dsub_rn(ir, il)), ih = np_max(0.0, __dsub_rn(ib, it));
    const double inter = __dmul_rn(iw, ih);
    const double total = __dmul_rn(w23.x, w23.y);
    double e = __dsub_rn(1.0, __ddiv_rn(inter, total));
    if (total == 0.0) e = 0.0;
    err[i] = e;
}

__global__ void mse_error_kernel(const double* __restrict__ wrm, const double* __restrict__ mic,
                                 double* __restrict__ err, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 w01 = reinterpret_cast<const double2*>(wrm)[2 * i], w23 = reinterpret_cast<const double2*>(wrm)[2 * i + 1];
    const double2 m01 = reinterpret_cast<const double2*>(mic)[2 * i], m23 = reinterpret_cast<const double2*>(mic)[2 * i + 1];
    const double wcx = __dadd_rn(w01.x, __ddiv_rn(w23.x, 2.0)), wcy = __dadd_rn(w01.y, __ddiv_rn(w23.y, 2.0));
    const double mcx = __dadd_rn(m01.x, __ddiv_rn(m23.x, 2.0)), mcy = __dadd_rn(m01.y, __ddiv_rn(m23.y, 2.0));
    const double dx = __dsub_rn(wcx, mcx), dy = __dsub_rn(wcy, mcy);
    // np.mean over 2 elements: (dx^2 + dy^2) / 2
    err[i] = __ddiv_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 2.0);
}

}  // namespace
}  // namespace wt

extern "C" int wt_bbox_error(const double* worm, const double* mic, double* err, int64_t n, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(worm && mic && err, "null argument");
    bbox_error_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(worm, mic, err, n);
    WT_LAUNCHED();
    return 0;
}

extern "C" int wt_mse_error(const double* worm, const double* mic, double* err, int64_t n, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(worm && mic && err, "null argument");
    mse_error_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(worm, mic, err, n);
    WT_LAUNCHED();
    return 0;
}
