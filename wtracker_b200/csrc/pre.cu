// K1-K4: camera-view crop (replicate border) + letterbox resize of u8 grey frames.
//
// The reference pads the WHOLE frame with cv.copyMakeBorder(BORDER_REPLICATE) and slices the view
// (wtracker/sim/view_controller.py:45-61,158-172); here the crop is a clamp-addressed gather, so
// the padded frame never exists.  Resampling follows cv2.resize(INTER_LINEAR) on u8: 11-bit fixed
// point coefficient tables (built on the host exactly as OpenCV builds them) and
//   dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2,  r = s0*a0 + s1*a1.
// Letterbox padding is 114.  Outputs: u8 grey (product path) and/or f32 NCHW x3 in [0,1].
#include "../../include/wtracker_b200.h"
#include "common.cuh"

#include <algorithm>

namespace wt {
namespace {

struct PreParams {
    const uint8_t* frames;
    int n_frames, fh, fw;
    const int32_t* frame_idx;
    const int32_t* crop_x;
    const int32_t* crop_y;
    wt_letterbox lb;
    uint8_t* out_u8;
    float* out_f32;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Each thread produces 16 horizontally adjacent output pixels of one output row.
template <bool RESIZE>
__global__ void __launch_bounds__(256) pre_kernel(const PreParams p) {
    const int groups_per_row = (p.lb.dst_w + 15) >> 4;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (gid >= groups_per_row * p.lb.dst_h) return;
    const int Y = gid / groups_per_row;
    const int X0 = (gid - Y * groups_per_row) << 4;

    const int f = clampi(p.frame_idx[img], 0, p.n_frames - 1);   // a stale index must not read outside `frames`
    const uint8_t* frame = p.frames + size_t(f) * p.fh * p.fw;
    const int cx = p.crop_x[img], cy = p.crop_y[img];
    const int ry = Y - p.lb.pad_top;
    const bool row_in = ry >= 0 && ry < p.lb.new_h;

    uint8_t px[16];
    if (!row_in) {
#pragma unroll
        for (int i = 0; i < 16; ++i) px[i] = 114;
    } else if (!RESIZE) {
        const uint8_t* srow = frame + size_t(clampi(cy + ry, 0, p.fh - 1)) * p.fw;
        const int rx0 = X0 - p.lb.pad_left;
        const uint8_t* src = srow + cx + rx0;
        // Interior fast path (all 16 pixels inside the view and the frame, one more group of slack to the right so
        // the second vector load stays inside the row): two aligned 128-bit loads, byte-aligned with funnel shifts.
        // The crop origin is arbitrary, so the misalignment m is only uniform per image row.
        if (rx0 >= 0 && rx0 + 16 <= p.lb.new_w && cx + rx0 >= 0 && cx + rx0 + 32 <= p.fw) {
            const uint32_t m = uint32_t(reinterpret_cast<uintptr_t>(src)) & 15u;
            const uint4* base = reinterpret_cast<const uint4*>(src - m);
            const uint4 lo = __ldg(base), hi = __ldg(base + 1);
            const uint32_t sh = (m & 3u) * 8u;
            uint32_t w0, w1, w2, w3, w4;
            switch (m >> 2) {
                case 0: w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; break;
                case 1: w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; break;
                case 2: w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; break;
                default: w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; break;
            }
            uint4 v;
            v.x = __funnelshift_r(w0, w1, sh);
            v.y = __funnelshift_r(w1, w2, sh);
            v.z = __funnelshift_r(w2, w3, sh);
            v.w = __funnelshift_r(w3, w4, sh);
            if (p.out_u8 && (p.lb.dst_w & 15) == 0 && !p.out_f32) {   // product path: straight to the 128-bit store
                *reinterpret_cast<uint4*>(p.out_u8 + (size_t(img) * p.lb.dst_h + Y) * p.lb.dst_w + X0) = v;
                return;
            }
            const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) px[i] = uint8_t(vw[i >> 2] >> ((i & 3) * 8));
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int rx = rx0 + i;
                px[i] = (rx >= 0 && rx < p.lb.new_w) ? __ldg(srow + clampi(cx + rx, 0, p.fw - 1)) : uint8_t(114);
            }
        }
    } else {
        // vertical taps: rows yofs, yofs+1 of the VIEW, clipped to the view, then to the frame
        const int sy = p.lb.yofs[ry];
        const int vy0 = clampi(sy, 0, p.lb.src_h - 1), vy1 = clampi(sy + 1, 0, p.lb.src_h - 1);
        const uint8_t* r0 = frame + size_t(clampi(cy + vy0, 0, p.fh - 1)) * p.fw;
        const uint8_t* r1 = frame + size_t(clampi(cy + vy1, 0, p.fh - 1)) * p.fw;
        const int b0 = p.lb.ycoef[2 * ry], b1 = p.lb.ycoef[2 * ry + 1];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int rx = X0 + i - p.lb.pad_left;
            if (rx >= 0 && rx < p.lb.new_w) {
                const int sx = p.lb.xofs[rx];
                const int a0 = p.lb.xcoef[2 * rx], a1 = p.lb.xcoef[2 * rx + 1];
                const int fx0 = clampi(cx + sx, 0, p.fw - 1);
                const int fx1 = clampi(cx + (sx + 1 < p.lb.src_w ? sx + 1 : sx), 0, p.fw - 1);
                const int h0 = int(__ldg(r0 + fx0)) * a0 + int(__ldg(r0 + fx1)) * a1;
                const int h1 = int(__ldg(r1 + fx0)) * a0 + int(__ldg(r1 + fx1)) * a1;
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                px[i] = uint8_t(clampi(v, 0, 255));
            } else {
                px[i] = 114;
            }
        }
    }

    const size_t plane = size_t(p.lb.dst_h) * p.lb.dst_w;
    const bool full = X0 + 16 <= p.lb.dst_w;
    if (p.out_u8) {
        uint8_t* o = p.out_u8 + size_t(img) * plane + size_t(Y) * p.lb.dst_w + X0;
        if (full && (p.lb.dst_w & 15) == 0) {
            uint4 v;
            v.x = px[0] | (px[1] << 8) | (px[2] << 16) | (uint32_t(px[3]) << 24);
            v.y = px[4] | (px[5] << 8) | (px[6] << 16) | (uint32_t(px[7]) << 24);
            v.z = px[8] | (px[9] << 8) | (px[10] << 16) | (uint32_t(px[11]) << 24);
            v.w = px[12] | (px[13] << 8) | (px[14] << 16) | (uint32_t(px[15]) << 24);
            *reinterpret_cast<uint4*>(o) = v;
        } else {
            for (int i = 0; i < 16 && X0 + i < p.lb.dst_w; ++i) o[i] = px[i];
        }
    }
    if (p.out_f32) {
        float* o = p.out_f32 + size_t(img) * 3 * plane + size_t(Y) * p.lb.dst_w + X0;
        for (int i = 0; i < 16 && X0 + i < p.lb.dst_w; ++i) {
            const float v = __fdiv_rn(float(px[i]), 255.0f);
            o[i] = v;
            o[plane + i] = v;
            o[2 * plane + i] = v;
        }
    }
}

// Pure crop (view == network input, no letterbox padding, u8 output only): the product path of the 640x640
// configuration.  One thread moves a 16-pixel x 4-row block: the three per-image index loads are amortised over 128
// bytes, and all eight 128-bit loads are issued before the first store.
constexpr int kCropRows = 4;

__device__ __forceinline__ uint4 load16_unaligned(const uint8_t* src) {
    const uint32_t m = uint32_t(reinterpret_cast<uintptr_t>(src)) & 15u;
    const uint4* base = reinterpret_cast<const uint4*>(src - m);
    const uint4 lo = __ldg(base), hi = __ldg(base + 1);
    const uint32_t sh = (m & 3u) * 8u;
    uint32_t w0, w1, w2, w3, w4;
    switch (m >> 2) {
        case 0: w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; break;
        case 1: w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; break;
        case 2: w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; break;
        default: w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                      __funnelshift_r(w3, w4, sh));
}

__global__ void __launch_bounds__(256) pre_crop_kernel(const PreParams p) {
    const int groups_per_row = p.lb.dst_w >> 4;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (gid >= groups_per_row * (p.lb.dst_h / kCropRows)) return;
    const int Yq = gid / groups_per_row;
    const int X0 = (gid - Yq * groups_per_row) << 4;
    const uint8_t* frame = p.frames + size_t(clampi(p.frame_idx[img], 0, p.n_frames - 1)) * p.fh * p.fw;
    const int cx = p.crop_x[img], cy = p.crop_y[img];
    const bool x_in = cx + X0 >= 0 && cx + X0 + 32 <= p.fw;   // one group of slack for the second vector load
    uint4 v[kCropRows];
#pragma unroll
    for (int r = 0; r < kCropRows; ++r) {
        const uint8_t* srow = frame + size_t(clampi(cy + Yq * kCropRows + r, 0, p.fh - 1)) * p.fw;
        if (x_in) {
            v[r] = load16_unaligned(srow + cx + X0);
        } else {   // the view hangs over the left / right frame border: replicate per byte
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i >> 2] |= uint32_t(__ldg(srow + clampi(cx + X0 + i, 0, p.fw - 1))) << ((i & 3) * 8);
            v[r] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    uint8_t* o = p.out_u8 + (size_t(img) * p.lb.dst_h + Yq * kCropRows) * p.lb.dst_w + X0;
#pragma unroll
    for (int r = 0; r < kCropRows; ++r) *reinterpret_cast<uint4*>(o + size_t(r) * p.lb.dst_w) = v[r];
}

// Resize path of the product (u8-only) output.  The kernel above gives every thread 16 adjacent output pixels, so a
// warp's byte gathers touch ~4 cache lines per load and the resize ran at 3-5 % of HBM bandwidth (L1 wavefront
// bound).  Here one CTA produces kResizeRows output rows of one image: the source rows it needs are staged ONCE in
// shared memory with coalesced, clamp-addressed loads (== BORDER_REPLICATE), a thread owns output COLUMNS (its
// horizontal taps and weights stay in registers) and walks the rows, so shared-memory reads and global stores of a
// warp are contiguous.  Same integer arithmetic as above: bit-exact against cv2.resize's fixed-point model.
constexpr int kResizeThreads = 256;
constexpr int kResizeRows = 16;

__global__ void __launch_bounds__(kResizeThreads) pre_resize_kernel(const PreParams p, int pitch) {
    extern __shared__ uint8_t s_rows[];   // [rows of this tile][pitch]
    __shared__ int s_r0[kResizeRows], s_r1[kResizeRows], s_b0[kResizeRows], s_b1[kResizeRows];
    const wt_letterbox& lb = p.lb;
    const int img = blockIdx.y, Y0 = blockIdx.x * kResizeRows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows_here = min(kResizeRows, lb.dst_h - Y0);
    const uint8_t* frame = p.frames + size_t(clampi(p.frame_idx[img], 0, p.n_frames - 1)) * p.fh * p.fw;
    const int cx = p.crop_x[img], cy = p.crop_y[img];

    // view rows this tile reads: [vbase, vtop]
    const int ry_lo = max(Y0 - lb.pad_top, 0), ry_hi = min(Y0 + rows_here - lb.pad_top, lb.new_h) - 1;
    int vbase = 0, nrows = 0;
    if (ry_lo <= ry_hi) {
        vbase = clampi(lb.yofs[ry_lo], 0, lb.src_h - 1);
        nrows = clampi(lb.yofs[ry_hi] + 1, 0, lb.src_h - 1) - vbase + 1;
    }
    if (tid < rows_here) {
        const int ry = Y0 + tid - lb.pad_top;
        int r0 = -1, r1 = -1, b0 = 0, b1 = 0;
        if (ry >= 0 && ry < lb.new_h) {
            const int sy = lb.yofs[ry];
            r0 = (clampi(sy, 0, lb.src_h - 1) - vbase) * pitch;
            r1 = (clampi(sy + 1, 0, lb.src_h - 1) - vbase) * pitch;
            b0 = lb.ycoef[2 * ry];
            b1 = lb.ycoef[2 * ry + 1];
        }
        s_r0[tid] = r0; s_r1[tid] = r1; s_b0[tid] = b0; s_b1[tid] = b1;
    }
    for (int r = warp; r < nrows; r += kResizeThreads / 32) {
        const uint8_t* srow = frame + size_t(clampi(cy + vbase + r, 0, p.fh - 1)) * p.fw;
        for (int c = lane; c < lb.src_w; c += 32) s_rows[r * pitch + c] = __ldg(srow + clampi(cx + c, 0, p.fw - 1));
    }
    __syncthreads();

    uint8_t* out = p.out_u8 + (size_t(img) * lb.dst_h + Y0) * lb.dst_w;
    for (int x = tid; x < lb.dst_w; x += kResizeThreads) {
        const int rx = x - lb.pad_left;
        const bool col_in = rx >= 0 && rx < lb.new_w;
        int sx = 0, sx1 = 0, a0 = 0, a1 = 0;
        if (col_in) {
            sx = lb.xofs[rx];
            sx1 = sx + 1 < lb.src_w ? sx + 1 : sx;
            a0 = lb.xcoef[2 * rx];
            a1 = lb.xcoef[2 * rx + 1];
        }
#pragma unroll 4
        for (int j = 0; j < rows_here; ++j) {
            const int r0 = s_r0[j];
            int v = 114;
            if (col_in && r0 >= 0) {
                const uint8_t* q0 = s_rows + r0;
                const uint8_t* q1 = s_rows + s_r1[j];
                const int h0 = int(q0[sx]) * a0 + int(q0[sx1]) * a1;
                const int h1 = int(q1[sx]) * a0 + int(q1[sx1]) * a1;
                v = clampi((((s_b0[j] * (h0 >> 4)) >> 16) + ((s_b1[j] * (h1 >> 4)) >> 16) + 2) >> 2, 0, 255);
            }
            out[size_t(j) * lb.dst_w + x] = uint8_t(v);
        }
    }
}

}  // namespace
}  // namespace wt

extern "C" int wt_preprocess(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* frame_idx,
                             const int32_t* crop_x, const int32_t* crop_y, int n, const wt_letterbox* lb,
                             uint8_t* out_u8, float* out_f32, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(frames && frame_idx && crop_x && crop_y && lb, "null argument");
    WT_REQUIRE(out_u8 || out_f32, "no output requested");
    WT_REQUIRE(n_frames >= 1 && frame_h >= 1 && frame_w >= 1, "empty frame buffer");
    // the 128-bit load paths align on absolute addresses and may touch up to 15 bytes in front of a row: legal inside
    // an allocation whose base is 16-byte aligned (every cudaMalloc / torch tensor start), not for an odd sub-view
    WT_REQUIRE(reinterpret_cast<uintptr_t>(frames) % 16 == 0, "frames must be 16-byte aligned");
    WT_REQUIRE(lb->new_w + lb->pad_left <= lb->dst_w && lb->new_h + lb->pad_top <= lb->dst_h, "letterbox geometry");
    const bool resize = !(lb->new_w == lb->src_w && lb->new_h == lb->src_h);
    if (resize) WT_REQUIRE(lb->xofs && lb->xcoef && lb->yofs && lb->ycoef, "resize tables missing");
    if (n == 0) return 0;
    PreParams p;
    p.frames = frames;
    p.n_frames = n_frames;
    p.fh = frame_h;
    p.fw = frame_w;
    p.frame_idx = frame_idx;
    p.crop_x = crop_x;
    p.crop_y = crop_y;
    p.lb = *lb;
    p.out_u8 = out_u8;
    p.out_f32 = out_f32;
    const int groups = ((lb->dst_w + 15) / 16) * lb->dst_h;
    dim3 grid((groups + 255) / 256, n);
    if (!resize && out_u8 && !out_f32 && lb->pad_left == 0 && lb->pad_top == 0 && lb->new_w == lb->dst_w &&
        lb->new_h == lb->dst_h && lb->dst_w % 16 == 0 && lb->dst_h % kCropRows == 0) {
        const int cgroups = (lb->dst_w / 16) * (lb->dst_h / kCropRows);
        pre_crop_kernel<<<dim3((cgroups + 255) / 256, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
        WT_LAUNCHED();
        return 0;
    }
    if (resize && out_u8 && !out_f32) {
        // rows of the view one 16-row output tile can touch (+1 for the second tap, +2 for rounding at both ends)
        const int max_rows = std::min(lb->src_h, ((kResizeRows - 1) * lb->src_h + lb->new_h - 1) / lb->new_h + 3);
        const int pitch = (lb->src_w + 15) & ~15;
        const int smem = max_rows * pitch;
        if (smem <= 200 * 1024) {
            static SmemOptIn opt_in;
            WT_CHECK_CUDA(opt_in_smem(pre_resize_kernel, opt_in, 200 * 1024));
            dim3 rgrid((lb->dst_h + kResizeRows - 1) / kResizeRows, n);
            pre_resize_kernel<<<rgrid, kResizeThreads, smem, static_cast<cudaStream_t>(stream)>>>(p, pitch);
            WT_LAUNCHED();
            return 0;
        }
    }
    if (resize) pre_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else pre_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
