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

    const int f = p.frame_idx[img];
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
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int rx = X0 + i - p.lb.pad_left;
            px[i] = (rx >= 0 && rx < p.lb.new_w) ? __ldg(srow + clampi(cx + rx, 0, p.fw - 1)) : uint8_t(114);
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

}  // namespace
}  // namespace wt

extern "C" int wt_preprocess(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* frame_idx,
                             const int32_t* crop_x, const int32_t* crop_y, int n, const wt_letterbox* lb,
                             uint8_t* out_u8, float* out_f32, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(frames && frame_idx && crop_x && crop_y && lb, "null argument");
    WT_REQUIRE(out_u8 || out_f32, "no output requested");
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
    if (resize) pre_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else pre_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
