// CUDA-core ops of the detector program: the first layer (u8 grey -> 32 ch, direct 3x3/s2 conv),
// the SPPF max-pool chain, nearest 2x upsampling into a concat slice, and a scalar convolution used
// only to validate the tcgen05 kernel (tests / WT conv_impl=1).
#include "../../include/wtracker_b200.h"
#include "conv.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace wt {

namespace {

__device__ __forceinline__ float silu_f(float v) {   // h * tanh(h) + h, h = v / 2 (one MUFU op)
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// ------------------------------------------------------------------ scalar validation conv
__global__ void conv_simt_kernel(const __nv_bfloat16* __restrict__ src, int sh, int sw, int sct, int scoff,
                                 void* __restrict__ dst, int dh, int dw, int dct, int dcoff, int dst_f32,
                                 const __nv_bfloat16* __restrict__ res, int rct, int rcoff,
                                 const __nv_bfloat16* __restrict__ wgt, const float* __restrict__ bias, int cin,
                                 int cout, int k, int stride, int act, const float* __restrict__ add, int act_ct,
                                 int acoff, long long total) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int co = int(idx % cout);
    long long pix = idx / cout;
    const int x = int(pix % dw);
    pix /= dw;
    const int y = int(pix % dh);
    const int n = int(pix / dh);
    const int pad = k / 2;
    float acc = 0.f;
    for (int kh = 0; kh < k; ++kh) {
        const int iy = y * stride + kh - pad;
        if (iy < 0 || iy >= sh) continue;
        for (int kw = 0; kw < k; ++kw) {
            const int ix = x * stride + kw - pad;
            if (ix < 0 || ix >= sw) continue;
            const __nv_bfloat16* ip = src + ((size_t(n) * sh + iy) * sw + ix) * sct + scoff;
            const __nv_bfloat16* wp = wgt + ((size_t(co) * k + kh) * k + kw) * cin;
            for (int c = 0; c < cin; ++c) acc = fmaf(__bfloat162float(ip[c]), __bfloat162float(wp[c]), acc);
        }
    }
    float v = acc + bias[co];
    if (add) v += add[((size_t(n) * (dh / 2) + (y >> 1)) * (dw / 2) + (x >> 1)) * act_ct + acoff + co];
    if (act == WT_ACT_SILU) v = silu_f(v);
    const size_t opix = (size_t(n) * dh + y) * dw + x;
    if (res) v += __bfloat162float(res[opix * rct + rcoff + co]);
    if (dst_f32) static_cast<float*>(dst)[opix * dct + dcoff + co] = v;
    else static_cast<__nv_bfloat16*>(dst)[opix * dct + dcoff + co] = __float2bfloat16_rn(v);
}

// scalar validation of a conv with a fused 1-channel head: one thread per pixel walks all channels
__global__ void conv_dot_simt_kernel(const __nv_bfloat16* __restrict__ src, int sh, int sw, int sct, int scoff,
                                     float* __restrict__ dst, int dh, int dw, const __nv_bfloat16* __restrict__ wgt,
                                     const float* __restrict__ bias, const float* __restrict__ dot_w, int cin, int cout,
                                     int k, int stride, int act, long long total) {
    long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const long long opix = pix;
    const int x = int(pix % dw);
    pix /= dw;
    const int y = int(pix % dh);
    const int n = int(pix / dh);
    const int pad = k / 2;
    float dot = 0.f;
    for (int co = 0; co < cout; ++co) {
        float acc = 0.f;
        for (int kh = 0; kh < k; ++kh) {
            const int iy = y * stride + kh - pad;
            if (iy < 0 || iy >= sh) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int ix = x * stride + kw - pad;
                if (ix < 0 || ix >= sw) continue;
                const __nv_bfloat16* ip = src + ((size_t(n) * sh + iy) * sw + ix) * sct + scoff;
                const __nv_bfloat16* wp = wgt + ((size_t(co) * k + kh) * k + kw) * cin;
                for (int c = 0; c < cin; ++c) acc = fmaf(__bfloat162float(ip[c]), __bfloat162float(wp[c]), acc);
            }
        }
        float v = acc + bias[co];
        if (act == WT_ACT_SILU) v = silu_f(v);
        dot = fmaf(v, dot_w[co], dot);
    }
    dst[opix] = dot + dot_w[cout];
}

// ------------------------------------------------------------------ first layer
// u8 grey -> COUT channels, 3x3 stride 2 pad 1, fp32 math.  One thread = 4 horizontally adjacent
// output pixels x 8 output channels, walking kConv0Rows output rows: the 72 weights of its channel
// group are loaded once and stay in registers, the 3-row input window slides (one 8-byte + one 1-byte
// load per new input row), and the four lanes of a pixel quad together store 64 contiguous bytes per
// pixel.  Issue-bound on the FMA/MUFU pipes (~14 instructions per output), not on HBM.
constexpr int kConv0Rows = 16;
constexpr int kConv0Threads = 128;

struct Conv0Raw {   // one input row segment as loaded: columns 2*x0 .. 2*x0+7 and column 2*x0-1
    uint2 v;
    uint32_t left;
};

__device__ __forceinline__ Conv0Raw conv0_load_row(const uint8_t* __restrict__ img, int w, int h, int iy, int x0) {
    Conv0Raw r;
    r.v = make_uint2(0u, 0u);
    r.left = 0u;
    if (iy >= 0 && iy < h) {
        const uint8_t* rowp = img + size_t(iy) * w + 2 * x0;
        r.v = __ldg(reinterpret_cast<const uint2*>(rowp));       // 8-byte aligned: w % 8 == 0, x0 % 4 == 0
        if (x0 > 0) r.left = __ldg(rowp - 1);                    // zero padding at x = -1
    }
    return r;
}

// u8 -> f32 without the conversion pipe: byte k of v placed in the mantissa of 2^23 (one PRMT), minus 2^23
// (one FADD).  I2F runs on the 16-lane XU pipe shared with MUFU, which bounded this kernel.
template <int K>
__device__ __forceinline__ float byte_to_float(uint32_t v) {
    return __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440u | K)) - 8388608.0f;
}

__device__ __forceinline__ void conv0_unpack(const Conv0Raw& r, float (&in)[9]) {
    in[0] = byte_to_float<0>(r.left);
    in[1] = byte_to_float<0>(r.v.x);
    in[2] = byte_to_float<1>(r.v.x);
    in[3] = byte_to_float<2>(r.v.x);
    in[4] = byte_to_float<3>(r.v.x);
    in[5] = byte_to_float<0>(r.v.y);
    in[6] = byte_to_float<1>(r.v.y);
    in[7] = byte_to_float<2>(r.v.y);
    in[8] = byte_to_float<3>(r.v.y);
}

// packed fp32 pair FMA (sm_100 FFMA2): d = a * (x, x) + d on two channels at once; the scalar operand is broadcast
// by the instruction itself, so the 288 FMAs of a thread's row step are 144 issue slots
__device__ __forceinline__ void ffma2_bcast(uint64_t& acc, float x, uint64_t w) {
    uint64_t xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(w));
}
template <int COUT, int MINB>
__global__ void __launch_bounds__(kConv0Threads, MINB) conv0_kernel(const uint8_t* __restrict__ src, int h, int w,
                                                              const float* __restrict__ w9,
                                                              const float* __restrict__ bias, int act,
                                                              __nv_bfloat16* __restrict__ dst, int dct, int dcoff,
                                                              long long total_threads) {
    constexpr int G = COUT / 8;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= total_threads) return;
    const int g = int(idx % G);
    long long q = idx / G;
    const int ho = h / 2, wo = w / 2, qw = wo / 4;
    const int strips = (ho + kConv0Rows - 1) / kConv0Rows;
    const int xq = int(q % qw);
    q /= qw;
    const int ys = int(q % strips);
    const int n = int(q / strips);
    const int x0 = xq * 4;
    const int y_begin = ys * kConv0Rows;
    const int y_end = min(y_begin + kConv0Rows, ho);
    const uint8_t* img = src + size_t(n) * h * w;

    // weights are constants: loaded while the previous kernel (the crop / letterbox) may still be draining.
    // SiLU(v) = hh * tanh(hh) + hh with hh = v / 2: weights and bias are halved here (exact), so the FMA chain
    // yields hh directly.  Channel pairs (2j, 2j + 1) live in one 64-bit register pair for FFMA2.
    ptx::grid_launch_dependents();
    const float ws = act == WT_ACT_SILU ? 0.5f : 1.0f;
    uint64_t wreg[9][4], breg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        breg[j] = ptx::pack_f32x2(ws * __ldg(bias + g * 8 + 2 * j), ws * __ldg(bias + g * 8 + 2 * j + 1));
#pragma unroll
        for (int t = 0; t < 9; ++t)
            wreg[t][j] = ptx::pack_f32x2(ws * __ldg(w9 + (g * 8 + 2 * j) * 9 + t), ws * __ldg(w9 + (g * 8 + 2 * j + 1) * 9 + t));
    }
    ptx::grid_dependency_wait();
    Conv0Raw raw0 = conv0_load_row(img, w, h, 2 * y_begin - 1, x0);
    Conv0Raw raw1 = conv0_load_row(img, w, h, 2 * y_begin, x0);
    Conv0Raw raw2 = conv0_load_row(img, w, h, 2 * y_begin + 1, x0);
    Conv0Raw raw3 = conv0_load_row(img, w, y_begin + 1 < y_end ? h : 0, 2 * y_begin + 2, x0);
    Conv0Raw raw4 = conv0_load_row(img, w, y_begin + 1 < y_end ? h : 0, 2 * y_begin + 3, x0);
    float r0[9], r1[9], r2[9];          // input rows 2y-1, 2y, 2y+1
    conv0_unpack(raw0, r0);
    for (int y = y_begin; y < y_end; ++y) {
        conv0_unpack(raw1, r1);
        conv0_unpack(raw2, r2);
        // the two new input rows of an output row are loaded TWO iterations ahead: one iteration (~1.2 us of this
        // warp's time at 12 warps / SM) is about one DRAM latency under load, and ncu showed 20 % of the warp time
        // waiting for the one-ahead prefetch.  (Rows past the strip are only read inside the image, rows past the
        // image return zeros without a load.)
        raw1 = raw3;
        raw2 = raw4;
        raw3 = conv0_load_row(img, w, y + 2 < y_end ? h : 0, 2 * y + 4, x0);
        raw4 = conv0_load_row(img, w, y + 2 < y_end ? h : 0, 2 * y + 5, x0);
        uint64_t acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[p][j] = breg[j];
        // same summation order per channel as the scalar form: rows 2y-1, 2y, 2y+1, taps left to right
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) ffma2_bcast(acc[p][j], r0[2 * p + kw], wreg[kw][j]);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) ffma2_bcast(acc[p][j], r1[2 * p + kw], wreg[3 + kw][j]);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) ffma2_bcast(acc[p][j], r2[2 * p + kw], wreg[6 + kw][j]);
        __nv_bfloat16* out = dst + ((size_t(n) * ho + y) * wo + x0) * dct + dcoff + g * 8;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo, hi;
                if (act == WT_ACT_SILU) {
                    ptx::unpack_f32x2(acc[p][j], lo, hi);
                    float tl, th;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(tl) : "f"(lo));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hi));
                    uint64_t o = acc[p][j];
                    const uint64_t t2 = ptx::pack_f32x2(tl, th);
                    asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(o) : "l"(acc[p][j]), "l"(t2));
                    ptx::unpack_f32x2(o, lo, hi);
                } else {
                    ptx::unpack_f32x2(acc[p][j], lo, hi);
                }
                __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                pk[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(out + size_t(p) * dct) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) r0[c] = r2[c];
    }
}

// ------------------------------------------------------------------ SPPF pooling chain
// One CTA per (image, 8-channel group): the 8 channels of a pixel are one 16-byte word, the slice
// (h*w words) lives in shared memory, and each 5x5 max-pool is a horizontal then a vertical 5-tap max
// (out-of-image taps ignored == -inf padding).  Small CTAs (128 threads, ~13 KB at 20x20) so that the
// whole grid is resident at once: the kernel is a chain of six barrier-separated passes, i.e. latency-bound.
constexpr int kPoolCg = 8;
constexpr int kPoolThreads = 128;
constexpr int kPoolIter = 4;          // pixels per thread: feature maps up to 512 pixels (20x20 at 640, 12x12 at 384)

__device__ __forceinline__ uint4 max_bf16x8(const uint4 a, const uint4 b) {
    uint4 r;
    __nv_bfloat162 t;
    t = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.x), *reinterpret_cast<const __nv_bfloat162*>(&b.x));
    r.x = *reinterpret_cast<uint32_t*>(&t);
    t = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.y), *reinterpret_cast<const __nv_bfloat162*>(&b.y));
    r.y = *reinterpret_cast<uint32_t*>(&t);
    t = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.z), *reinterpret_cast<const __nv_bfloat162*>(&b.z));
    r.z = *reinterpret_cast<uint32_t*>(&t);
    t = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.w), *reinterpret_cast<const __nv_bfloat162*>(&b.w));
    r.w = *reinterpret_cast<uint32_t*>(&t);
    return r;
}

// G channel groups (8 channels = one 16-byte word each) per CTA, thread = (pixel lane, group): the G words of a pixel
// are contiguous in global memory (full 32-byte sectors instead of half ones), and at G = 4 the whole grid (c / 32
// x images CTAs of 512 threads) is resident in ONE wave on 148 SMs — with G = 1 the last 15 % of the CTAs cost a second
// round of this latency-bound chain.
template <int G>
__global__ void __launch_bounds__(kPoolThreads * G) sppf_pool_kernel(const __nv_bfloat16* __restrict__ src, int sct,
                                                                     int scoff, __nv_bfloat16* __restrict__ dst, int dct,
                                                                     int dcoff, int c, int h, int w) {
    extern __shared__ uint4 pool_smem[];
    const int hw = h * w;
    uint4* cur = pool_smem;              // [hw][G]
    uint4* tmp = pool_smem + hw * G;     // [hw][G]
    const int n = blockIdx.y;
    const int g = threadIdx.x % G, lane_px = threadIdx.x / G;
    const int c0 = (blockIdx.x * G + g) * kPoolCg;
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    for (int i = lane_px; i < hw; i += kPoolThreads)
        cur[i * G + g] = __ldg(reinterpret_cast<const uint4*>(src + (size_t(n) * hw + i) * sct + scoff + c0));
    // a thread owns at most kPoolIter pixels (hw <= kPoolIter * 128): their coordinates are computed once, not with
    // a division in every one of the six passes (the kernel is instruction-issue bound at full occupancy)
    int px[kPoolIter], py[kPoolIter];
#pragma unroll
    for (int k = 0; k < kPoolIter; ++k) {
        const int i = lane_px + k * kPoolThreads;
        py[k] = i < hw ? i / w : -1000;      // (out-of-range slots fail every tap test below)
        px[k] = i < hw ? i - py[k] * w : -1000;
    }
    __syncthreads();
    for (int round = 0; round < 3; ++round) {
#pragma unroll
        for (int k = 0; k < kPoolIter; ++k) {
            const int i = lane_px + k * kPoolThreads, x = px[k];
            if (x < 0) continue;
            uint4 m = cur[i * G + g];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && x + d >= 0 && x + d < w) m = max_bf16x8(m, cur[(i + d) * G + g]);
            tmp[i * G + g] = m;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPoolIter; ++k) {
            const int i = lane_px + k * kPoolThreads, y = py[k];
            if (y < 0) continue;
            uint4 m = tmp[i * G + g];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && y + d >= 0 && y + d < h) m = max_bf16x8(m, tmp[(i + d * w) * G + g]);
            *reinterpret_cast<uint4*>(dst + (size_t(n) * hw + i) * dct + dcoff + round * c + c0) = m;
            cur[i * G + g] = m;     // only this thread touches element (i, g) before the barrier
        }
        __syncthreads();
    }
}

// Sliding-window form for maps up to 32 x 32 (the SPPF of every supported network size): 4 channel groups per CTA as above,
// but a thread is a whole ROW of one group in the horizontal passes and a whole COLUMN in the vertical ones and keeps the
// 5-wide window in registers — one shared-memory load and one store per output and pass instead of five loads (the
// kernel above is LDS-wavefront bound: ncu, round 2).  m[x] = max(p[x-2], v[x], p[x+1]) with the pair maxima
// p[x] = max(v[x], v[x+1]) carried along; outside the map the window holds -inf.  Rows are padded to w + 1 pixels so that
// the two rows (columns) of a quarter-warp fall into different halves of the 32 banks.
constexpr uint32_t kNegInf2 = 0xFF80FF80u;   // bf16x2 (-inf, -inf)

// Outputs [s0, s1) of one line (row or column) of length `len`: chunks of kPoolChunk outputs whose kPoolChunk + 4 inputs
// sit in registers with static indices (no window rotation: ncu showed the rotating form spending most of its issue slots
// on register moves and per-step bounds checks).
constexpr int kPoolChunk = 7;

template <bool kToGlobal>
__device__ __forceinline__ void pool_slide(const uint4* in, uint4* out, int len, int pitch, __nv_bfloat16* gdst,
                                           size_t gpitch, int s0, int s1) {
    const uint4 neg = make_uint4(kNegInf2, kNegInf2, kNegInf2, kNegInf2);
    for (int c = s0; c < s1; c += kPoolChunk) {
        uint4 v[kPoolChunk + 4];
#pragma unroll
        for (int i = 0; i < kPoolChunk + 4; ++i) {
            const int x = c - 2 + i;
            v[i] = (x >= 0 && x < len && x < s1 + 2) ? in[x * pitch] : neg;
        }
        uint4 pr[kPoolChunk + 3];                                       // pr[i] = max(v[i], v[i + 1])
#pragma unroll
        for (int i = 0; i < kPoolChunk + 3; ++i) pr[i] = max_bf16x8(v[i], v[i + 1]);
#pragma unroll
        for (int o = 0; o < kPoolChunk; ++o) {
            const int x = c + o;
            if (x < s1) {
                const uint4 m = max_bf16x8(max_bf16x8(pr[o], v[o + 2]), pr[o + 3]);
                out[x * pitch] = m;
                if (kToGlobal) *reinterpret_cast<uint4*>(gdst + size_t(x) * gpitch) = m;
            }
        }
    }
}

// 256 threads = 64 (line, segment) units x 4 channel groups: a row (column) is walked in `segs` pieces by different
// threads (each starts its window two pixels early), so 60 of the 64 units work on a 20 x 20 map.
__global__ void __launch_bounds__(256, 4) sppf_pool_slide_kernel(const __nv_bfloat16* __restrict__ src, int sct, int scoff,
                                                              __nv_bfloat16* __restrict__ dst, int dct, int dcoff, int c,
                                                              int h, int w, int segs) {
    extern __shared__ uint4 pool_smem[];
    constexpr int G = 4;
    const int pw = w + 1, hw = h * w;
    uint4* cur = pool_smem;                   // [h][pw][G]
    uint4* tmp = pool_smem + h * pw * G;
    const int n = blockIdx.y;
    const int g = threadIdx.x & (G - 1), u = threadIdx.x >> 2, units = blockDim.x >> 2;
    const int c0 = (blockIdx.x * G + g) * kPoolCg;
    const int line = u / segs, seg = u - line * segs;
    const int wseg = (w + segs - 1) / segs, hseg = (h + segs - 1) / segs;
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    // every thread of the CTA loads: independent loads, four in flight per thread
#pragma unroll 4
    for (int i = u; i < hw; i += units) {
        const int y = i / w;
        cur[(y * pw + (i - y * w)) * G + g] = __ldg(reinterpret_cast<const uint4*>(src + (size_t(n) * hw + i) * sct + scoff + c0));
    }
    __syncthreads();
    for (int round = 0; round < 3; ++round) {
        if (line < h)
            pool_slide<false>(cur + line * pw * G + g, tmp + line * pw * G + g, w, G, nullptr, 0, seg * wseg,
                              min(w, (seg + 1) * wseg));
        __syncthreads();
        if (line < w)
            pool_slide<true>(tmp + line * G + g, cur + line * G + g, h, pw * G,
                             dst + (size_t(n) * hw + line) * dct + dcoff + round * c + c0, size_t(w) * dct, seg * hseg,
                             min(h, (seg + 1) * hseg));
        __syncthreads();
    }
}

// Any map size (one channel group per CTA, coordinates recomputed per pass): used above 512 pixels.
__global__ void __launch_bounds__(kPoolThreads) sppf_pool_generic_kernel(const __nv_bfloat16* __restrict__ src, int sct,
                                                                         int scoff, __nv_bfloat16* __restrict__ dst,
                                                                         int dct, int dcoff, int c, int h, int w) {
    extern __shared__ uint4 pool_smem[];
    const int hw = h * w;
    uint4* cur = pool_smem;
    uint4* tmp = pool_smem + hw;
    const int n = blockIdx.y;
    const int c0 = blockIdx.x * kPoolCg;
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    for (int i = threadIdx.x; i < hw; i += kPoolThreads)
        cur[i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t(n) * hw + i) * sct + scoff + c0));
    __syncthreads();
    for (int round = 0; round < 3; ++round) {
        for (int i = threadIdx.x; i < hw; i += kPoolThreads) {
            const int x = i % w;
            uint4 m = cur[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && x + d >= 0 && x + d < w) m = max_bf16x8(m, cur[i + d]);
            tmp[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < hw; i += kPoolThreads) {
            const int y = i / w;
            uint4 m = tmp[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && y + d >= 0 && y + d < h) m = max_bf16x8(m, tmp[i + d * w]);
            *reinterpret_cast<uint4*>(dst + (size_t(n) * hw + i) * dct + dcoff + round * c + c0) = m;
            cur[i] = m;
        }
        __syncthreads();
    }
}

}  // namespace

int conv_simt_launch(const ConvDesc& d, int n_images, cudaStream_t stream) {
    WT_REQUIRE(d.src.dtype == WT_DT_BF16, "conv input must be bf16");
    if (d.dot_w) {
        const long long pixels = (long long)n_images * d.dst.h * d.dst.w;
        if (pixels == 0) return 0;
        conv_dot_simt_kernel<<<(unsigned)((pixels + 127) / 128), 128, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(d.src.base), d.src.h, d.src.w, d.src.ctot, d.src.coff,
            static_cast<float*>(d.dst.base), d.dst.h, d.dst.w, d.w, d.bias, d.dot_w, d.cin, d.cout, d.k, d.stride, d.act,
            pixels);
        WT_LAUNCHED();
        return 0;
    }
    const long long total = (long long)n_images * d.dst.h * d.dst.w * d.cout;
    if (total == 0) return 0;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    conv_simt_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(d.src.base), d.src.h, d.src.w, d.src.ctot, d.src.coff, d.dst.base, d.dst.h,
        d.dst.w, d.dst.ctot, d.dst.coff, d.dst.dtype == WT_DT_F32 ? 1 : 0,
        static_cast<const __nv_bfloat16*>(d.res.base), d.res.ctot, d.res.coff, d.w, d.bias, d.cin, d.cout, d.k,
        d.stride, d.act, static_cast<const float*>(d.add.base), d.add.ctot, d.add.coff, total);
    WT_LAUNCHED();
    return 0;
}

int conv0_launch(const uint8_t* src, int h, int w, const float* w9, const float* bias, int cout, int act,
                 const TensorView& dst, int n_images, cudaStream_t stream) {
    WT_REQUIRE(cout == 32 || cout == 16 || cout == 64, "conv0 supports 16/32/64 output channels");
    WT_REQUIRE(h % 2 == 0 && w % 8 == 0, "conv0 needs an even height and a width that is a multiple of 8");
    WT_REQUIRE(dst.dtype == WT_DT_BF16 && dst.h == h / 2 && dst.w == w / 2, "conv0 destination shape");
    WT_REQUIRE(dst.ctot % 8 == 0 && dst.coff % 8 == 0, "conv0 destination channel alignment");
    const int strips = (h / 2 + kConv0Rows - 1) / kConv0Rows;
    const long long total = (long long)n_images * strips * (w / 8) * (cout / 8);   // row strips x pixel quads x channel groups
    if (total == 0) return 0;
    const int threads = kConv0Threads;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(dst.base);
    auto* kernel = cout == 32 ? conv0_kernel<32, 3> : (cout == 16 ? conv0_kernel<16, 3> : conv0_kernel<64, 3>);
    WT_CHECK_CUDA(launch_pdl(kernel, dim3(blocks), dim3(threads), 0, stream, src, h, w, w9, bias, act, out, dst.ctot,
                             dst.coff, total));
    WT_LAUNCHED();
    return 0;
}

int sppf_pool_launch(const TensorView& src, const TensorView& dst, int c, int n_images, cudaStream_t stream) {
    WT_REQUIRE(c % kPoolCg == 0 && src.ctot % 8 == 0 && dst.ctot % 8 == 0 && src.coff % 8 == 0 && dst.coff % 8 == 0,
               "SPPF channels and slices must be multiples of 8");
    WT_REQUIRE(src.h == dst.h && src.w == dst.w, "SPPF keeps the spatial size");
    WT_REQUIRE(src.dtype == WT_DT_BF16 && dst.dtype == WT_DT_BF16, "SPPF works on bf16");
    if (n_images == 0) return 0;
    if (src.h <= 32 && src.w <= 32 && c % (4 * kPoolCg) == 0) {   // every YOLOv8 SPPF up to 1024 x 1024 inputs
        const size_t smem = size_t(src.h) * (src.w + 1) * 4 * sizeof(uint4) * 2;
        static SmemOptIn opt_in;
        WT_CHECK_CUDA(opt_in_smem(sppf_pool_slide_kernel, opt_in, smem));
        WT_CHECK_CUDA(launch_pdl(sppf_pool_slide_kernel, dim3(c / (kPoolCg * 4), n_images), dim3(256), smem,
                                 stream, static_cast<const __nv_bfloat16*>(src.base), src.ctot, src.coff,
                                 static_cast<__nv_bfloat16*>(dst.base), dst.ctot, dst.coff, c, src.h, src.w,
                                 64 / (src.h > src.w ? src.h : src.w)));
        WT_LAUNCHED();
        return 0;
    }
    const bool small = src.h * src.w <= kPoolIter * kPoolThreads;
    const int G = (small && c % (4 * kPoolCg) == 0) ? 4 : 1;
    const size_t smem = size_t(src.h) * src.w * sizeof(uint4) * 2 * G;
    WT_REQUIRE(smem <= 200 * 1024, "SPPF feature map too large for the shared-memory pool kernel");
    auto* kernel = !small ? sppf_pool_generic_kernel : (G == 4 ? sppf_pool_kernel<4> : sppf_pool_kernel<1>);
    static SmemOptIn opt_in[3];
    WT_CHECK_CUDA(opt_in_smem(kernel, opt_in[!small ? 2 : (G == 4)], smem));
    dim3 grid(c / (kPoolCg * G), n_images);
    WT_CHECK_CUDA(launch_pdl(kernel, grid, dim3(kPoolThreads * G), smem, stream,
                             static_cast<const __nv_bfloat16*>(src.base), src.ctot, src.coff,
                             static_cast<__nv_bfloat16*>(dst.base), dst.ctot, dst.coff, c, src.h, src.w));
    WT_LAUNCHED();
    return 0;
}

}  // namespace wt
