// The tail of the hot path as ONE launch: tracking rows -> ResMLP input gather -> ResMLP -> bbox error.
//
// The four separate launches (wt_track_rows, wt_mlp_gather, wt_resmlp_forward, wt_bbox_error) cost 47 us per
// 64-frame step on a B200 — launch latency and, in resmlp_warp_kernel, every CTA re-transposing the 22 KB weight set
// with a division per element.  Here a warp owns one frame of the batch:
//   * lane j < k gathers box j of the ResMLP input: a row of THIS batch is computed straight from the detector
//     output (the same arithmetic track_rows_kernel uses), a row of an earlier batch is read from the tracking
//     table, so no thread depends on another thread's global write and the batch can span many CTAs;
//   * lane 0 writes the frame's own tracking row, its microscope box and the bbox error;
//   * the warp runs the ResMLP with lane o owning output neuron o (and o + 32), bias first and inputs in order —
//     the same accumulation order as resmlp_kernel / resmlp_warp_kernel / resmlp_pair_kernel, hence bit-identical.
// The weights come pre-transposed ([in][out] then bias[out] per layer, packed once on the host), so staging them is
// a straight 16-byte-vector copy.
//   replaces logging_controller.py:152-155 + view_controller.py:93-117 (rows), mlp_controllers.py:38-56 (gather),
//   neural/mlp.py:176-188 (ResMLP), eval/error_calculator.py:163-195 (bbox error) of the reference.
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

constexpr int kTailThreads = 256;
constexpr int kTailWarps = kTailThreads / 32;

struct TailParams {
    wt_tail_args a;
    int maxw;
    int n_weights_pad;   // floats of the staged blob, padded to a multiple of 4
};

__device__ __forceinline__ double np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// worm row (x, y, w, h) in frame pixels of batch item i; NaN row when nothing was detected (== track_rows_kernel)
__device__ __forceinline__ void worm_row(const wt_tail_args& a, long long i, double* w) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    w[0] = w[1] = w[2] = w[3] = nan;
    if (a.count[i] > 0) {
        const float* b = a.boxes + i * a.max_det * 6;
        w[0] = __dadd_rn(double(b[0]), double(a.crop_x[i]));
        w[1] = __dadd_rn(double(b[1]), double(a.crop_y[i]));
        w[2] = double(__fsub_rn(b[2], b[0]));
        w[3] = double(__fsub_rn(b[3], b[1]));
    }
}

__device__ __forceinline__ void dense_warp_t(const float* __restrict__ wt, const float* __restrict__ b, const float* in,
                                             float* out, int nin, int nout, bool relu, int lane) {
    for (int o = lane; o < nout; o += 32) {
        float acc = b[o];
#pragma unroll 8
        for (int i = 0; i < nin; ++i) acc = fmaf(in[i], wt[i * nout + o], acc);
        out[o] = relu ? fmaxf(acc, 0.f) : acc;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kTailThreads) hot_tail_kernel(const TailParams p) {
    extern __shared__ __align__(16) float tail_smem[];
    const wt_tail_args& a = p.a;
    const wt_resmlp_desc& d = a.mlp;
    float* sw = tail_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = sw + p.n_weights_pad + warp * 3 * p.maxw;
    float* t0 = xs + p.maxw;
    float* t1 = t0 + p.maxw;
    {
        const float4* src = reinterpret_cast<const float4*>(a.weights_t);
        float4* dst = reinterpret_cast<float4*>(sw);
        for (int i = threadIdx.x; i < p.n_weights_pad / 4; i += kTailThreads) dst[i] = __ldg(src + i);
    }
    const long long i = (long long)blockIdx.x * kTailWarps + warp;
    const bool live = i < a.n;
    bool ok = true;
    if (live) {
        // ---- gather: lane j holds input box j
        double b[4] = {0.0, 0.0, 0.0, 0.0};
        if (lane < a.k) {
            const long long r = a.first_row + i + a.offsets[lane];
            if (r >= a.first_row && r < a.first_row + a.n) {
                worm_row(a, r - a.first_row, b);
            } else if (r >= 0 && r < a.table_rows) {
                const double2 lo = reinterpret_cast<const double2*>(a.table)[2 * r], hi = reinterpret_cast<const double2*>(a.table)[2 * r + 1];
                b[0] = lo.x; b[1] = lo.y; b[2] = hi.x; b[3] = hi.y;
            } else {
                b[0] = b[1] = b[2] = b[3] = __longlong_as_double(0x7ff8000000000000LL);
            }
            ok = isfinite(b[0]) && isfinite(b[1]) && isfinite(b[2]) && isfinite(b[3]);
        }
        ok = __all_sync(0xffffffffu, ok);
        const double x0 = __shfl_sync(0xffffffffu, b[0], 0), y0 = __shfl_sync(0xffffffffu, b[1], 0);
        if (lane < a.k) {
            const float v0 = float(__dsub_rn(b[0], x0)), v1 = float(__dsub_rn(b[1], y0)), v2 = float(b[2]), v3 = float(b[3]);
            t0[4 * lane] = v0; t0[4 * lane + 1] = v1; t0[4 * lane + 2] = v2; t0[4 * lane + 3] = v3;
            if (a.x) reinterpret_cast<float4*>(a.x)[i * a.k + lane] = make_float4(v0, v1, v2, v3);
        }
        // ---- this frame's own row, microscope box and bbox error
        if (lane == 0) {
            double w[4];
            worm_row(a, i, w);
            const long long r = a.first_row + i;
            reinterpret_cast<double2*>(a.table)[2 * r] = make_double2(w[0], w[1]);
            reinterpret_cast<double2*>(a.table)[2 * r + 1] = make_double2(w[2], w[3]);
            const double cx = double(a.crop_x[i]), cy = double(a.crop_y[i]);
            const double mx = cx + double(a.cam_w / 2) - double(a.mic_w / 2), my = cy + double(a.cam_h / 2) - double(a.mic_h / 2);
            const double mw = double(a.mic_w), mh = double(a.mic_h);
            reinterpret_cast<double2*>(a.mic_table)[2 * r] = make_double2(mx, my);
            reinterpret_cast<double2*>(a.mic_table)[2 * r + 1] = make_double2(mw, mh);
            // ErrorCalculator.calculate_bbox_error, op for op (== bbox_error_kernel)
            const double wr = __dadd_rn(w[0], w[2]), wb = __dadd_rn(w[1], w[3]);
            const double mr = __dadd_rn(mx, mw), mb = __dadd_rn(my, mh);
            const double il = np_max(w[0], mx), it = np_max(w[1], my);
            const double ir = np_min(wr, mr), ib = np_min(wb, mb);
            const double iw = np_max(0.0, __dsub_rn(ir, il)), ih = np_max(0.0, __dsub_rn(ib, it));
            const double inter = __dmul_rn(iw, ih);
            const double total = __dmul_rn(w[2], w[3]);
            double e = __dsub_rn(1.0, __ddiv_rn(inter, total));
            if (total == 0.0) e = 0.0;
            a.err[i] = e;
            a.valid[i] = ok ? 1 : 0;
        }
    }
    __syncthreads();   // weights staged (all threads reach this: no early return above)
    if (!live) return;

    const int H = d.hidden, ind = d.in_dim;
    const float* w = sw;
    dense_warp_t(w, w + H * ind, t0, xs, ind, H, true, lane);
    w += H * ind + H;
    for (int blk = 0; blk < d.n_blocks; ++blk) {
        const float* in = xs;
        int nin = H;
        float* bufs[2] = {t0, t1};
        for (int l = 0; l < d.block_len; ++l) {
            const int nout = d.block_dims[l];
            float* out = bufs[l & 1];
            dense_warp_t(w, w + nout * nin, in, out, nin, nout, true, lane);
            w += nout * nin + nout;
            in = out;
            nin = nout;
        }
        for (int f = lane; f < H; f += 32) xs[f] += in[f];
        __syncwarp();
    }
    dense_warp_t(w, w + d.out_dim * H, xs, t0, H, d.out_dim, false, lane);
    for (int f = lane; f < d.out_dim; f += 32) a.y[i * d.out_dim + f] = t0[f];
}

// one row of the gathered per-frame result table: 8 x 32 bit
__global__ void result_rows_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ count, int max_det,
                                   long long first_frame, int32_t* __restrict__ rows, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float nan = __int_as_float(0x7fc00000);
    float x = nan, y = nan, w = nan, h = nan, conf = nan;
    int anchor = -1, valid = 0;
    if (count[i] > 0) {
        const float* b = boxes + i * max_det * 6;
        x = b[0];
        y = b[1];
        w = __fsub_rn(b[2], b[0]);   // BoxConverter.to_xywh on the fp32 result (yolo_controller.py:86-88)
        h = __fsub_rn(b[3], b[1]);
        conf = b[4];
        anchor = int(b[5]);
        valid = 1;
    }
    int4* o = reinterpret_cast<int4*>(rows + 8 * i);
    o[0] = make_int4(__float_as_int(x), __float_as_int(y), __float_as_int(w), __float_as_int(h));
    o[1] = make_int4(__float_as_int(conf), anchor, int(first_frame + i), valid);
}

}  // namespace
}  // namespace wt

extern "C" int wt_hot_tail(const wt_tail_args* args, void* stream) {
    using namespace wt;
    WT_REQUIRE(args, "null argument");
    const wt_tail_args& a = *args;
    if (a.n == 0) return 0;
    WT_REQUIRE(a.boxes && a.count && a.crop_x && a.crop_y && a.table && a.mic_table && a.weights_t && a.y && a.valid && a.err,
               "null argument");
    WT_REQUIRE(a.k >= 1 && a.k <= WT_TAIL_MAX_K && a.mlp.in_dim == 4 * a.k, "1..16 input boxes, in_dim = 4 k");
    WT_REQUIRE(a.first_row >= 0 && a.first_row + a.n <= a.table_rows, "rows outside the tracking table");
    WT_REQUIRE(a.max_det >= 1, "max_det");
    const wt_resmlp_desc& d = a.mlp;
    WT_REQUIRE(d.block_len >= 1 && d.block_len <= 8 && d.n_blocks >= 0, "block shape");
    WT_REQUIRE(d.block_dims[d.block_len - 1] == d.hidden, "a block must map hidden -> hidden");
    int maxw = d.in_dim > d.hidden ? d.in_dim : d.hidden;
    long long expect = (long long)d.hidden * d.in_dim + d.hidden;
    int nin = d.hidden;
    long long per_block = 0;
    for (int l = 0; l < d.block_len; ++l) {
        per_block += (long long)d.block_dims[l] * nin + d.block_dims[l];
        nin = d.block_dims[l];
        if (nin > maxw) maxw = nin;
    }
    expect += per_block * d.n_blocks + (long long)d.out_dim * d.hidden + d.out_dim;
    if (d.out_dim > maxw) maxw = d.out_dim;
    WT_REQUIRE(expect == d.n_weights, "weight blob size does not match the layer description");
    WT_REQUIRE(reinterpret_cast<uintptr_t>(a.weights_t) % 16 == 0, "weights_t must be 16-byte aligned");
    TailParams p;
    p.a = a;
    p.maxw = (maxw + 3) & ~3;
    p.n_weights_pad = (d.n_weights + 3) & ~3;   // (the host pads the transposed blob to a multiple of 4 floats)
    const size_t smem = (size_t(p.n_weights_pad) + size_t(kTailWarps) * 3 * p.maxw) * sizeof(float);
    WT_REQUIRE(smem <= 220 * 1024, "ResMLP too large for the shared-memory kernel");
    static SmemOptIn opt_in;
    WT_CHECK_CUDA(opt_in_smem(hot_tail_kernel, opt_in, smem));
    const long long blocks = (a.n + kTailWarps - 1) / kTailWarps;
    hot_tail_kernel<<<(unsigned)blocks, kTailThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}

extern "C" int wt_result_rows(const float* boxes, const int32_t* count, int max_det, int64_t first_frame, int32_t* rows,
                              int64_t n, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(boxes && count && rows && max_det >= 1, "null argument");
    WT_REQUIRE(first_frame >= 0 && first_frame + n <= 0x7fffffffLL, "frame indices are stored as int32");
    WT_REQUIRE(reinterpret_cast<uintptr_t>(rows) % 16 == 0, "rows must be 16-byte aligned");
    result_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(boxes, count, max_det,
                                                                                                 first_frame, rows, n);
    WT_LAUNCHED();
    return 0;
}
