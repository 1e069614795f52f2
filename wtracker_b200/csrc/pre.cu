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
    // the 16 pixels lie inside the frame row; the second vector load may run up to 16 bytes past them — into the next
    // row (bytes shifted out) — so it only has to end inside the `frames` allocation (tight windows: fw == view width)
    const bool x_in = cx + X0 >= 0 && cx + X0 + 16 <= p.fw;
    const uint8_t* frames_end = p.frames + size_t(p.n_frames) * p.fh * p.fw;
    uint4 v[kCropRows];
#pragma unroll
    for (int r = 0; r < kCropRows; ++r) {
        const uint8_t* srow = frame + size_t(clampi(cy + Yq * kCropRows + r, 0, p.fh - 1)) * p.fw;
        if (x_in && srow + cx + X0 + 32 <= frames_end) {
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

// Resize path of the product (u8-only) output.  One CTA produces kResizeRows output rows of one image.
//  * The source rows it needs are staged ONCE in shared memory, clamp-addressed (== BORDER_REPLICATE): 32-bit loads
//    re-aligned with a funnel shift where the view lies inside the frame, byte loads where it hangs over the border.
//  * A thread owns 4 adjacent output columns over kResizeSpan consecutive output rows.  Its horizontal taps live in
//    registers: per staged row two 32-bit shared-memory loads cover the <= 8 source bytes the four columns touch, one
//    PRMT per column puts its (left, right) byte pair in place and one DP2A applies the two 11-bit weights.  The
//    horizontal results of the previous output row's tap pair are kept: when shrinking, last row's bottom tap is this
//    row's top tap (one new row per output row); when enlarging, half of the rows repeat the pair (nothing new).
//    The vertical pass is two IMAD.HI per pixel (b << 16 pre-shifted per row) and one 32-bit store per row.
//  * ncu (round 2, profiles/r02_simt_resize_pool.summary.txt): the kernel is instruction-issue bound (75 % of the issue
//    slots); the first form of this design spent 105 instructions per 4 pixels on cache bookkeeping and 64-bit store
//    addresses, this one ~45.
// Same integer arithmetic as pre_kernel: bit-exact against cv2.resize's fixed-point model.
constexpr int kResizeRows = 32;
constexpr int kResizeSpan = 16;

__device__ __forceinline__ uint32_t dp2a_u(uint32_t a, uint32_t b, uint32_t c) {   // a.lo16 * b.byte0 + a.hi16 * b.byte1 + c
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct ResizeTaps {
    uint32_t coef[4];   // a0 | a1 << 16
    uint32_t sel[4];    // PRMT selector: byte 0 = left tap, byte 1 = right tap (offsets from the item's first word)
    int sx[4], sx1[4];  // byte path (taps further apart than two words)
    int word;           // first 32-bit word of the row the item reads
    bool packed;
    uint32_t in_mask;   // bit i: column i lies inside the resized image
};

// horizontal pass of one staged row for the item's four columns: (s0 * a0 + s1 * a1) >> 4, the form the vertical pass uses
__device__ __forceinline__ void resize_hrow(const uint8_t* s_rows, int r, const ResizeTaps& t, int h[4]) {
    if (t.packed) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(s_rows + r) + t.word;
        const uint32_t lo = q[0], hi = q[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = int(dp2a_u(t.coef[i], __byte_perm(lo, hi, t.sel[i]), 0u) >> 4);
    } else {
        const uint8_t* q = s_rows + r;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            h[i] = (int(q[t.sx[i]]) * int(t.coef[i] & 0xFFFFu) + int(q[t.sx1[i]]) * int(t.coef[i] >> 16)) >> 4;
    }
}

__device__ __forceinline__ ResizeTaps resize_taps(const wt_letterbox& lb, int X0) {
    ResizeTaps t;
    t.in_mask = 0u;
    int lo_b = 1 << 30, hi_b = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rx = X0 + i - lb.pad_left;
        t.sx[i] = t.sx1[i] = 0;
        t.coef[i] = 0u;
        if (rx >= 0 && rx < lb.new_w) {
            t.in_mask |= 1u << i;
            t.sx[i] = lb.xofs[rx];
            t.sx1[i] = t.sx[i] + 1 < lb.src_w ? t.sx[i] + 1 : t.sx[i];
            t.coef[i] = uint32_t(uint16_t(lb.xcoef[2 * rx])) | (uint32_t(uint16_t(lb.xcoef[2 * rx + 1])) << 16);
            lo_b = min(lo_b, t.sx[i]);
            hi_b = max(hi_b, t.sx1[i]);
        }
    }
    t.word = hi_b >= 0 ? lo_b >> 2 : 0;
    t.packed = hi_b < 0 || hi_b - (t.word << 2) < 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)   // (upper selector bytes: don't care; columns outside the image are overwritten with 114)
        t.sel[i] = t.packed ? (uint32_t(t.sx[i] - (t.word << 2)) & 7u) | ((uint32_t(t.sx1[i] - (t.word << 2)) & 7u) << 4) | 0x4400u : 0x4400u;
    return t;
}

__global__ void __launch_bounds__(256) pre_resize_kernel(const PreParams p, int pitch) {
    extern __shared__ __align__(16) uint8_t s_rows[];   // [rows of this tile][pitch]
    __shared__ int4 s_row[kResizeRows];                 // per output row: (top row offset | -1, bottom row offset, b0 << 16, b1 << 16)
    const wt_letterbox& lb = p.lb;
    const int img = blockIdx.y, Y0 = blockIdx.x * kResizeRows;
    // thread = (column group lane, row span): blockDim.y == kResizeRows / kResizeSpan
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warps = (blockDim.x * blockDim.y) >> 5;
    const int rows_here = min(kResizeRows, lb.dst_h - Y0);
    const uint8_t* frame = p.frames + size_t(clampi(p.frame_idx[img], 0, p.n_frames - 1)) * p.fh * p.fw;
    const int cx = p.crop_x[img], cy = p.crop_y[img];

    // view rows this tile reads: [vbase, vbase + nrows)
    const int ry_lo = max(Y0 - lb.pad_top, 0), ry_hi = min(Y0 + rows_here - lb.pad_top, lb.new_h) - 1;
    int vbase = 0, nrows = 0;
    if (ry_lo <= ry_hi) {
        vbase = clampi(lb.yofs[ry_lo], 0, lb.src_h - 1);
        nrows = clampi(lb.yofs[ry_hi] + 1, 0, lb.src_h - 1) - vbase + 1;
    }
    if (tid < rows_here) {
        const int ry = Y0 + tid - lb.pad_top;
        int4 e = make_int4(-1, -1, 0, 0);
        if (ry >= 0 && ry < lb.new_h) {
            const int sy = lb.yofs[ry];
            e.x = (clampi(sy, 0, lb.src_h - 1) - vbase) * pitch;
            e.y = (clampi(sy + 1, 0, lb.src_h - 1) - vbase) * pitch;
            e.z = int(lb.ycoef[2 * ry]) << 16;       // (b * t) >> 16 == umulhi(b << 16, t): 0 <= b <= 2048, t < 2^15
            e.w = int(lb.ycoef[2 * ry + 1]) << 16;
        }
        s_row[tid] = e;
    }
    // the taps of this thread's first column group: table loads in flight while the rows are staged
    const int groups = lb.dst_w >> 2;
    ResizeTaps t = resize_taps(lb, min(int(threadIdx.x), groups - 1) << 2);
    // word loads where the view's columns lie inside the frame row and the staged span (pitch bytes + one word of slack,
    // which may run into the next row: those bytes are never used) ends inside the `frames` allocation
    const bool x_in = cx >= 0 && cx + lb.src_w <= p.fw;
    const uint8_t* frames_end = p.frames + size_t(p.n_frames) * p.fh * p.fw;
    for (int r = warp; r < nrows; r += warps) {
        const uint8_t* srow = frame + size_t(clampi(cy + vbase + r, 0, p.fh - 1)) * p.fw;
        if (x_in && srow + cx + pitch + 4 <= frames_end) {
            const uint8_t* src = srow + cx;
            const uint32_t m = uint32_t(reinterpret_cast<uintptr_t>(src)) & 3u;
            const uint32_t* base = reinterpret_cast<const uint32_t*>(src - m);
            uint32_t* dst = reinterpret_cast<uint32_t*>(s_rows + r * pitch);
            for (int w = lane; w < (pitch >> 2); w += 32) dst[w] = __funnelshift_r(__ldg(base + w), __ldg(base + w + 1), m * 8u);
        } else {
            for (int c = lane; c < lb.src_w; c += 32) s_rows[r * pitch + c] = __ldg(srow + clampi(cx + c, 0, p.fw - 1));
        }
    }
    __syncthreads();

    const int j0 = threadIdx.y * kResizeSpan, j_end = min(rows_here, j0 + kResizeSpan);
    const int out_pitch = lb.dst_w >> 2;    // in 32-bit words
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        if (g != int(threadIdx.x)) t = resize_taps(lb, g << 2);
        uint32_t* o = reinterpret_cast<uint32_t*>(p.out_u8 + (size_t(img) * lb.dst_h + Y0 + j0) * lb.dst_w) + g;
        // (ka, ha) / (kb, hb): staged-row offset and horizontal result of the previous output row's top / bottom tap
        int ka = -1, kb = -1, ha[4] = {0, 0, 0, 0}, hb[4] = {0, 0, 0, 0};
        for (int j = j0; j < j_end; ++j, o += out_pitch) {
            const int4 e = s_row[j];
            uint32_t v = 0x72727272u;   // 114 x 4
            if (e.x >= 0 && t.in_mask) {
                if (e.x != ka || e.y != kb) {      // (equal: an enlarging resize repeats the tap pair)
                    if (e.x == kb) {               // the usual step: last row's bottom tap is this row's top tap
#pragma unroll
                        for (int i = 0; i < 4; ++i) ha[i] = hb[i];
                    } else {
                        resize_hrow(s_rows, e.x, t, ha);
                    }
                    if (e.y == e.x) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) hb[i] = ha[i];
                    } else {
                        resize_hrow(s_rows, e.y, t, hb);
                    }
                    ka = e.x; kb = e.y;
                }
                // weights are >= 0 and sum to 2048 +- 1 per axis, so ((b0 t0 >> 16) + (b1 t1 >> 16) + 2) >> 2 <= 255: no clamp
                uint32_t px[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t acc = __umulhi(uint32_t(e.z), uint32_t(ha[i])) + 2u;
                    acc = __umulhi(uint32_t(e.w), uint32_t(hb[i])) + acc;
                    px[i] = acc >> 2;
                }
                if (t.in_mask != 0xFu) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (!((t.in_mask >> i) & 1u)) px[i] = 114u;
                }
                v = __byte_perm(__byte_perm(px[0], px[1], 0x0040u), __byte_perm(px[2], px[3], 0x0040u), 0x5410u);
            }
            *o = v;
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
    if (resize && out_u8 && !out_f32 && lb->dst_w % 4 == 0) {
        // rows of the view one output tile can touch (+1 for the second tap, +2 for rounding at both ends)
        const int max_rows = std::min(lb->src_h, ((kResizeRows - 1) * lb->src_h + lb->new_h - 1) / lb->new_h + 3);
        const int pitch = (lb->src_w + 4 + 15) & ~15;     // (a word of slack: an item reads two words from its first tap)
        const int smem = max_rows * pitch;
        if (smem <= 200 * 1024) {
            static SmemOptIn opt_in;
            WT_CHECK_CUDA(opt_in_smem(pre_resize_kernel, opt_in, 200 * 1024));
            // thread = (column group lane, 16-row span): 384 wide -> 96 x 2 threads with one column group each, 640 wide ->
            // 96 x 2 threads with one or two (taps are loaded per column group and amortised over the 16 rows of a span)
            const int groups = lb->dst_w / 4;
            const int iters = (groups + 127) / 128;
            dim3 block(std::min(128, (((groups + iters - 1) / iters) + 31) & ~31), kResizeRows / kResizeSpan);
            dim3 rgrid((lb->dst_h + kResizeRows - 1) / kResizeRows, n);
            pre_resize_kernel<<<rgrid, block, smem, static_cast<cudaStream_t>(stream)>>>(p, pitch);
            WT_LAUNCHED();
            return 0;
        }
    }
    if (resize) pre_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else pre_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
