// K6-K8: class confidence, confidence filter, DFL box decode, greedy NMS, scale_boxes.
//
// One CTA per image.  (a) every warp walks anchors and computes the class logit (a 1x1 conv of the
// cls-branch features, fp32, or a ready-made logit), keeping anchors with sigmoid(logit) > conf;
// (b) candidates are ordered by (confidence desc, anchor index asc) with a shared-memory bitonic
// sort of 64-bit keys (or a single arg-max when max_det == 1); (c) only candidates are DFL-decoded;
// (d) warp 0 runs the greedy NMS against the list of boxes kept so far (a box survives iff no kept
// box overlaps it with IoU > thr — identical to torchvision.ops.nms on score-sorted boxes);
// (e) kept boxes are mapped back to the camera view like ultralytics scale_boxes + clip_boxes.
// Arithmetic is written with explicit round-to-nearest intrinsics so no FMA contraction changes
// the comparisons the reference makes in fp32.
#include "../../include/wtracker_b200.h"
#include "common.cuh"

namespace wt {
namespace {

constexpr int kMaxLevels = 4;
constexpr int kPostThreads = 512;
constexpr int kMaxKeep = 1024;
constexpr int kMaxClsC = 1024;

struct PostParams {
    wt_head_level lv[kMaxLevels];
    int level_start[kMaxLevels + 1];
    int n_levels;
    int total_anchors;
    int sort_cap;   // power of two >= total_anchors
    wt_post_params pp;
    float* out_boxes;
    int32_t* out_count;
    int box_from_feat;   // box logits are computed per surviving anchor (wt_head_level.box_feat)
    float* sc_logit;     // [n][A] class logits computed from cls_feat
    float* sc_box;       // [n][A][4] candidate boxes in sorted order (xyxy, letterboxed px)
    float* sc_conf;      // [n][A]
    int32_t* sc_idx;     // [n][A]
};

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ int level_of(const PostParams& p, int a) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < p.n_levels && a >= p.level_start[i]) l = i;
    return l;
}

// expectation of softmax(logits[16]) over bins 0..15
__device__ __forceinline__ float dfl_side(const float* lg) {
    float m = lg[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, lg[i]);
    float e[16], s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        e[i] = expf(__fsub_rn(lg[i], m));
        s = __fadd_rn(s, e[i]);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(e[i], s), float(i)));
    return acc;
}

// anchor-centre distances (l, t, r, b) in stride units -> xyxy in letterboxed pixels
__device__ __forceinline__ void box_from_sides(const wt_head_level& L, int local, float dl, float dt, float dr, float db,
                                               float* xyxy) {
    const float ax = float(local % L.w) + 0.5f, ay = float(local / L.w) + 0.5f;
    const float s = float(L.stride);
    // dist2bbox(xywh=True) * stride, then xywh2xyxy — same op order as the reference decode
    const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt), x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
    const float cxs = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.f), s), cys = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.f), s);
    const float ws = __fmul_rn(__fsub_rn(x2, x1), s), hs = __fmul_rn(__fsub_rn(y2, y1), s);
    const float hw2 = __fdiv_rn(ws, 2.f), hh2 = __fdiv_rn(hs, 2.f);
    xyxy[0] = __fsub_rn(cxs, hw2);
    xyxy[1] = __fsub_rn(cys, hh2);
    xyxy[2] = __fadd_rn(cxs, hw2);
    xyxy[3] = __fadd_rn(cys, hh2);
}

// One WARP per candidate when the box logits are not materialised: the final 1x1 conv of the box branch
// (64 outputs, box_c inputs, bf16 x bf16 products in fp32 like the tensor-core path) is evaluated for this
// anchor only — lane o computes outputs o and o + 32 — then four lanes run the DFL expectation of one side each.
__device__ void decode_box_from_feat(const PostParams& p, int img, int a, float* s_lg /* [64] per warp */, int lane,
                                     float* xyxy) {
    const int l = level_of(p, a);
    const wt_head_level& L = p.lv[l];
    const int local = a - p.level_start[l];
    const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(L.box_feat) + (size_t(img) * L.h * L.w + local) * L.box_c;
    const __nv_bfloat16* w0 = static_cast<const __nv_bfloat16*>(L.box_w) + size_t(lane) * L.box_c;
    const __nv_bfloat16* w1 = w0 + size_t(32) * L.box_c;
    float acc0 = 0.f, acc1 = 0.f;
    for (int c = 0; c < L.box_c; c += 8) {
        const uint4 fv = __ldg(reinterpret_cast<const uint4*>(f + c));       // same address in all lanes: broadcast
        const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(w0 + c));
        const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(w1 + c));
        const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w}, u0[4] = {a0.x, a0.y, a0.z, a0.w}, u1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc0 = fmaf(bf16_lo(fw[j]), bf16_lo(u0[j]), acc0);
            acc0 = fmaf(bf16_hi(fw[j]), bf16_hi(u0[j]), acc0);
            acc1 = fmaf(bf16_lo(fw[j]), bf16_lo(u1[j]), acc1);
            acc1 = fmaf(bf16_hi(fw[j]), bf16_hi(u1[j]), acc1);
        }
    }
    s_lg[lane] = acc0 + __ldg(L.box_b + lane);
    s_lg[lane + 32] = acc1 + __ldg(L.box_b + lane + 32);
    __syncwarp();
    float side = 0.f;
    if (lane < 4) side = dfl_side(s_lg + 16 * lane);
    const float dl = __shfl_sync(0xffffffffu, side, 0), dt = __shfl_sync(0xffffffffu, side, 1);
    const float dr = __shfl_sync(0xffffffffu, side, 2), db = __shfl_sync(0xffffffffu, side, 3);
    if (lane == 0) box_from_sides(L, local, dl, dt, dr, db, xyxy);
    __syncwarp();
}

__device__ void decode_box(const PostParams& p, int img, int a, float* xyxy) {
    const int l = level_of(p, a);
    const wt_head_level& L = p.lv[l];
    const int local = a - p.level_start[l];
    const int hw = L.h * L.w;
    float lg[64];
    if (L.box_dtype == WT_DT_F32) {
        const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(L.box) + (size_t(img) * hw + local) * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 v = __ldg(src + i);
            lg[4 * i] = v.x; lg[4 * i + 1] = v.y; lg[4 * i + 2] = v.z; lg[4 * i + 3] = v.w;
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(L.box) + (size_t(img) * hw + local) * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lg[8 * i + 2 * j] = bf16_lo(w[j]);
                lg[8 * i + 2 * j + 1] = bf16_hi(w[j]);
            }
        }
    }
    const float dl = dfl_side(lg), dt = dfl_side(lg + 16), dr = dfl_side(lg + 32), db = dfl_side(lg + 48);
    box_from_sides(L, local, dl, dt, dr, db, xyxy);
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr;
}

// Class logit of every anchor: the final 1x1 convolution of the cls branch (cout = nc = 1) as a dot
// product, 16 lanes per anchor (one 16-byte chunk of the bf16 feature vector each), fp32.
__global__ void __launch_bounds__(256) cls_logit_kernel(const PostParams p, int n) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long ag = t >> 4;
    const int sub = int(t & 15);
    const bool live = ag < (long long)n * p.total_anchors;
    float acc = 0.f;
    int img = 0, a = 0, l = 0;
    if (live) {
        img = int(ag / p.total_anchors);
        a = int(ag - (long long)img * p.total_anchors);
        l = level_of(p, a);
        const wt_head_level& L = p.lv[l];
        if (L.cls_feat) {
            const int local = a - p.level_start[l];
            const __nv_bfloat16* f = static_cast<const __nv_bfloat16*>(L.cls_feat) + (size_t(img) * L.h * L.w + local) * L.cls_c;
            const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(L.cls_w);
            for (int c = sub * 8; c < L.cls_c; c += 128) {
                const uint4 fv = __ldg(reinterpret_cast<const uint4*>(f + c));
                const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + c));
                const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc = fmaf(bf16_lo(fw[j]), bf16_lo(ww[j]), acc);
                    acc = fmaf(bf16_hi(fw[j]), bf16_hi(ww[j]), acc);
                }
            }
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && sub == 0 && p.lv[l].cls_feat) p.sc_logit[size_t(img) * p.total_anchors + a] = acc + p.lv[l].cls_b;
}

__global__ void __launch_bounds__(kPostThreads) post_kernel(const PostParams p) {
    extern __shared__ unsigned long long keys[];            // [sort_cap]
    __shared__ int s_count;
    __shared__ unsigned long long s_red[kPostThreads / 32];
    __shared__ float4 s_keep[kMaxKeep];
    __shared__ int s_keep_src[kMaxKeep];
    __shared__ int s_nkeep;
    __shared__ float s_lg[kPostThreads / 32][64];   // box logits of the candidate a warp is decoding

    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.total_anchors;

    if (tid == 0) { s_count = 0; s_nkeep = 0; }
    __syncthreads();

    // ---- (a) confidence + filter (logits are ready: either given or computed by cls_logit_kernel)
    // eight anchors per thread and trip: all eight logit loads are in flight before the first one is used (one
    // dependent load per trip made this loop a chain of ~17 DRAM latencies per image)
    constexpr int kPostUnroll = 8;
    for (int a0 = tid; a0 < A; a0 += kPostUnroll * kPostThreads) {
        float lg[kPostUnroll];
#pragma unroll
        for (int u = 0; u < kPostUnroll; ++u) {
            const int a = a0 + u * kPostThreads;
            lg[u] = 0.f;
            if (a < A) {
                const int l = level_of(p, a);
                const wt_head_level& L = p.lv[l];
                lg[u] = L.cls_logit ? __ldg(L.cls_logit + size_t(img) * L.h * L.w + (a - p.level_start[l]))
                                    : p.sc_logit[size_t(img) * A + a];
            }
        }
#pragma unroll
        for (int u = 0; u < kPostUnroll; ++u) {
            const int a = a0 + u * kPostThreads;
            const float conf = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-lg[u])));
            if (a < A && conf > p.pp.conf_thres) {
                const int pos = atomicAdd(&s_count, 1);
                keys[pos] = (static_cast<unsigned long long>(__float_as_uint(conf)) << 32) | (0xFFFFFFFFu - unsigned(a));
            }
        }
    }
    __syncthreads();
    const int n = s_count;
    if (n == 0) {
        if (tid == 0) p.out_count[img] = 0;
        return;
    }

    float* cbox = p.sc_box + size_t(img) * A * 4;
    float* cconf = p.sc_conf + size_t(img) * A;
    int32_t* cidx = p.sc_idx + size_t(img) * A;
    int n_sorted;

    if (p.pp.max_det == 1) {
        // ---- (b') arg-max: highest confidence, lowest anchor index on ties
        unsigned long long best = 0;
        for (int i = tid; i < n; i += kPostThreads) best = keys[i] > best ? keys[i] : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kPostThreads / 32; ++w) best = s_red[w] > best ? s_red[w] : best;
            keys[0] = best;
        }
        __syncthreads();
        n_sorted = 1;
    } else {
        // ---- (b) bitonic sort, descending
        int P = 1;
        while (P < n) P <<= 1;
        for (int i = n + tid; i < P; i += kPostThreads) keys[i] = 0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += kPostThreads) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = keys[i], b = keys[ixj];
                        const bool desc = (i & k) == 0;
                        if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        n_sorted = n;
    }

    // ---- (c) decode candidates in sorted order
    if (p.box_from_feat) {
        for (int i = warp; i < n_sorted; i += kPostThreads / 32) {       // one warp per candidate
            const unsigned long long k = keys[i];
            const int a = int(0xFFFFFFFFu - unsigned(k & 0xFFFFFFFFull));
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            decode_box_from_feat(p, img, a, s_lg[warp], lane, b);
            if (lane == 0) {
                cbox[4 * i] = b[0]; cbox[4 * i + 1] = b[1]; cbox[4 * i + 2] = b[2]; cbox[4 * i + 3] = b[3];
                cconf[i] = __uint_as_float(unsigned(k >> 32));
                cidx[i] = a;
            }
        }
    } else {
        for (int i = tid; i < n_sorted; i += kPostThreads) {
            const unsigned long long k = keys[i];
            const int a = int(0xFFFFFFFFu - unsigned(k & 0xFFFFFFFFull));
            float b[4];
            decode_box(p, img, a, b);
            cbox[4 * i] = b[0]; cbox[4 * i + 1] = b[1]; cbox[4 * i + 2] = b[2]; cbox[4 * i + 3] = b[3];
            cconf[i] = __uint_as_float(unsigned(k >> 32));
            cidx[i] = a;
        }
    }
    __syncthreads();

    // ---- (d) greedy NMS, warp 0
    if (warp == 0) {
        int nk = 0;
        const int max_det = p.pp.max_det < kMaxKeep ? p.pp.max_det : kMaxKeep;
        for (int i = 0; i < n_sorted && nk < max_det; ++i) {
            const float4 bi = make_float4(cbox[4 * i], cbox[4 * i + 1], cbox[4 * i + 2], cbox[4 * i + 3]);
            bool sup = false;
            for (int j = lane; j < nk; j += 32) sup |= iou_gt(s_keep[j], bi, p.pp.iou_thres);
            if (!__any_sync(0xffffffffu, sup)) {
                if (lane == 0) { s_keep[nk] = bi; s_keep_src[nk] = i; }
                ++nk;
                __syncwarp();
            }
        }
        if (lane == 0) s_nkeep = nk;
    }
    __syncthreads();

    // ---- (e) scale back to the camera view, clip, emit
    const int nk = s_nkeep;
    for (int k = tid; k < nk; k += kPostThreads) {
        const float4 b = s_keep[k];
        const int src = s_keep_src[k];
        float x1 = __fdiv_rn(__fsub_rn(b.x, p.pp.pad_x), p.pp.gain), y1 = __fdiv_rn(__fsub_rn(b.y, p.pp.pad_y), p.pp.gain);
        float x2 = __fdiv_rn(__fsub_rn(b.z, p.pp.pad_x), p.pp.gain), y2 = __fdiv_rn(__fsub_rn(b.w, p.pp.pad_y), p.pp.gain);
        const float W = float(p.pp.img_w), H = float(p.pp.img_h);
        x1 = fminf(fmaxf(x1, 0.f), W); x2 = fminf(fmaxf(x2, 0.f), W);
        y1 = fminf(fmaxf(y1, 0.f), H); y2 = fminf(fmaxf(y2, 0.f), H);
        float* o = p.out_boxes + (size_t(img) * p.pp.max_det + k) * 6;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        o[4] = cconf[src];
        o[5] = float(cidx[src]);
    }
    if (tid == 0) p.out_count[img] = nk;
}

// max_det == 1 (the reference's setting, yolo_controller.py:72-78): the kept box is the arg-max of the confidence —
// no key array, no sort, no suppression loop — so one image can be split over several CTAs.  CTA (split, img) scans its
// share of the anchors and folds its best key (confidence bits << 32 | ~anchor: highest confidence, lowest index on
// ties, exactly the order of the sort above) into best[img] with atomicMax; the CTA that arrives last at done[img]
// decodes that one box and resets both words, so the scratch stays zeroed between calls.
constexpr int kTop1Threads = 256;

__global__ void __launch_bounds__(kTop1Threads) post_top1_kernel(const PostParams p, int splits, unsigned long long* best,
                                                                 unsigned int* done) {
    __shared__ unsigned long long s_red[kTop1Threads / 32];
    __shared__ float s_lg[64];
    __shared__ int s_last;
    const int img = blockIdx.y, split = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.total_anchors;
    const int per = (A + splits - 1) / splits;
    const int a_lo = split * per, a_hi = min(A, a_lo + per);
    constexpr int kUnroll = 8;
    unsigned long long key = 0ull;
    for (int a0 = a_lo + tid; a0 < a_hi; a0 += kUnroll * kTop1Threads) {
        float lg[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {   // all loads in flight before the first use
            const int a = a0 + u * kTop1Threads;
            lg[u] = 0.f;
            if (a < a_hi) {
                const int l = level_of(p, a);
                const wt_head_level& L = p.lv[l];
                lg[u] = L.cls_logit ? __ldg(L.cls_logit + size_t(img) * L.h * L.w + (a - p.level_start[l]))
                                    : p.sc_logit[size_t(img) * A + a];
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = a0 + u * kTop1Threads;
            const float conf = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-lg[u])));
            if (a < a_hi && conf > p.pp.conf_thres) {
                const unsigned long long k = (static_cast<unsigned long long>(__float_as_uint(conf)) << 32) | (0xFFFFFFFFu - unsigned(a));
                key = k > key ? k : key;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane == 0) s_red[warp] = key;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kTop1Threads / 32; ++w) key = s_red[w] > key ? s_red[w] : key;
        if (key) atomicMax(best + img, key);
        __threadfence();
        s_last = atomicAdd(done + img, 1u) == unsigned(splits - 1);
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    unsigned long long k = 0ull;
    if (lane == 0) {
        k = atomicExch(best + img, 0ull);
        done[img] = 0u;
    }
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k == 0ull) {
        if (lane == 0) p.out_count[img] = 0;
        return;
    }
    const int a = int(0xFFFFFFFFu - unsigned(k & 0xFFFFFFFFull));
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.box_from_feat) decode_box_from_feat(p, img, a, s_lg, lane, b);
    else if (lane == 0) decode_box(p, img, a, b);
    if (lane == 0) {
        float x1 = __fdiv_rn(__fsub_rn(b[0], p.pp.pad_x), p.pp.gain), y1 = __fdiv_rn(__fsub_rn(b[1], p.pp.pad_y), p.pp.gain);
        float x2 = __fdiv_rn(__fsub_rn(b[2], p.pp.pad_x), p.pp.gain), y2 = __fdiv_rn(__fsub_rn(b[3], p.pp.pad_y), p.pp.gain);
        const float W = float(p.pp.img_w), H = float(p.pp.img_h);
        x1 = fminf(fmaxf(x1, 0.f), W); x2 = fminf(fmaxf(x2, 0.f), W);
        y1 = fminf(fmaxf(y1, 0.f), H); y2 = fminf(fmaxf(y2, 0.f), H);
        float* o = p.out_boxes + size_t(img) * 6;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        o[4] = __uint_as_float(unsigned(k >> 32));
        o[5] = float(a);
        p.out_count[img] = 1;
    }
}

__global__ void track_rows_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ count, int max_det,
                                  const int32_t* __restrict__ crop_x, const int32_t* __restrict__ crop_y, int cam_w,
                                  int cam_h, int mic_w, int mic_h, double* __restrict__ worm, double* __restrict__ mic,
                                  long long n, int none_as_zero) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cx = double(crop_x[i]), cy = double(crop_y[i]);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double w0 = nan, w1 = nan, w2 = nan, w3 = nan;
    if (none_as_zero) w0 = w1 = w2 = w3 = 0.0;   // what bboxes.csv holds for such a frame (logging_controller.py:153-158)
    if (count[i] > 0) {
        const float* b = boxes + i * max_det * 6;
        // xyxy -> xywh in float32 (BoxConverter.to_xywh on the fp32 result), then shifted by the
        // camera origin in float64 (numpy promotes when the log adds the integer camera offset)
        w0 = __dadd_rn(double(b[0]), cx);
        w1 = __dadd_rn(double(b[1]), cy);
        w2 = double(__fsub_rn(b[2], b[0]));
        w3 = double(__fsub_rn(b[3], b[1]));
    }
    reinterpret_cast<double2*>(worm)[2 * i] = make_double2(w0, w1);
    reinterpret_cast<double2*>(worm)[2 * i + 1] = make_double2(w2, w3);
    // camera view origin = pos - cam//2  =>  pos = origin + cam//2 ; microscope origin = pos - mic//2
    const double mx = cx + double(cam_w / 2) - double(mic_w / 2), my = cy + double(cam_h / 2) - double(mic_h / 2);
    reinterpret_cast<double2*>(mic)[2 * i] = make_double2(mx, my);
    reinterpret_cast<double2*>(mic)[2 * i + 1] = make_double2(double(mic_w), double(mic_h));
}

}  // namespace
}  // namespace wt

extern "C" int wt_track_rows(const float* boxes, const int32_t* count, int max_det, const int32_t* crop_x,
                             const int32_t* crop_y, int cam_w, int cam_h, int mic_w, int mic_h, double* worm_xywh,
                             double* mic_xywh, int64_t n, int none_as_zero, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(boxes && count && crop_x && crop_y && worm_xywh && mic_xywh, "null argument");
    track_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        boxes, count, max_det, crop_x, crop_y, cam_w, cam_h, mic_w, mic_h, worm_xywh, mic_xywh, n, none_as_zero);
    WT_LAUNCHED();
    return 0;
}

extern "C" int64_t wt_post_scratch_bytes(int n, int total_anchors) {
    // candidate boxes / confidences / indices / logits, then (16-byte aligned) best[n] u64 + done[n] u32 of the arg-max form
    return ((int64_t(n) * total_anchors * (4 * 4 + 4 + 4 + 4) + 15) / 16) * 16 + int64_t(n) * 12 + 256;
}

extern "C" int wt_decode_nms(const wt_head_level* levels, int n_levels, int n, const wt_post_params* pp,
                             float* out_boxes, int32_t* out_count, void* scratch, void* stream) {
    using namespace wt;
    WT_REQUIRE(levels && pp && out_boxes && out_count && scratch, "null argument");
    WT_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "1..4 head levels");
    WT_REQUIRE(pp->max_det >= 1 && pp->max_det <= kMaxKeep, "max_det must be in [1, 1024]");
    PostParams p;
    int total = 0;
    for (int l = 0; l < n_levels; ++l) {
        p.lv[l] = levels[l];
        p.level_start[l] = total;
        total += levels[l].h * levels[l].w;
        WT_REQUIRE((levels[l].box != nullptr) != (levels[l].box_feat != nullptr), "give box or box_feat");
        if (levels[l].box_feat)
            WT_REQUIRE(levels[l].box_c % 8 == 0 && levels[l].box_c >= 8 && levels[l].box_w && levels[l].box_b,
                       "box feature channels / weights");
        WT_REQUIRE((levels[l].box_feat != nullptr) == (levels[0].box_feat != nullptr), "all levels give box or all give box_feat");
        WT_REQUIRE((levels[l].cls_feat != nullptr) != (levels[l].cls_logit != nullptr), "give cls_feat or cls_logit");
        if (levels[l].cls_feat)
            WT_REQUIRE(levels[l].cls_c % 8 == 0 && levels[l].cls_c <= kMaxClsC && levels[l].cls_w, "cls feature channels");
    }
    for (int l = n_levels; l <= kMaxLevels; ++l) p.level_start[l] = total;
    p.n_levels = n_levels;
    p.box_from_feat = levels[0].box_feat != nullptr;
    p.total_anchors = total;
    int cap = 1;
    while (cap < total) cap <<= 1;
    p.sort_cap = cap;
    const size_t smem = size_t(cap) * 8;
    WT_REQUIRE(smem <= 160 * 1024, "too many anchors for the shared-memory candidate sort (max 20480)");
    p.pp = *pp;
    p.out_boxes = out_boxes;
    p.out_count = out_count;
    uint8_t* sc = static_cast<uint8_t*>(scratch);
    p.sc_box = reinterpret_cast<float*>(sc);
    p.sc_conf = reinterpret_cast<float*>(sc + size_t(n) * total * 16);
    p.sc_idx = reinterpret_cast<int32_t*>(sc + size_t(n) * total * 20);
    p.sc_logit = reinterpret_cast<float*>(sc + size_t(n) * total * 24);
    if (n == 0) return 0;
    static SmemOptIn opt_in;
    WT_CHECK_CUDA(opt_in_smem(post_kernel, opt_in, smem));
    bool any_feat = false;
    for (int l = 0; l < n_levels; ++l) any_feat |= levels[l].cls_feat != nullptr;
    if (any_feat) {
        const long long threads = (long long)n * total * 16;
        cls_logit_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n);
        WT_LAUNCHED();
    }
    if (pp->max_det == 1) {
        // arg-max form: enough CTAs per image to fill the GPU twice over (n = 64 -> 5 per image, n = 1 -> 16)
        int sm_count = 148;
        wt_device_info(&sm_count, nullptr, nullptr);
        int splits = (2 * sm_count + n - 1) / n;
        splits = splits < 1 ? 1 : (splits > 16 ? 16 : splits);
        uint8_t* tail = sc + ((size_t(n) * total * 28 + 15) / 16) * 16;
        post_top1_kernel<<<dim3(splits, n), kTop1Threads, 0, static_cast<cudaStream_t>(stream)>>>(
            p, splits, reinterpret_cast<unsigned long long*>(tail), reinterpret_cast<unsigned int*>(tail + size_t(n) * 8));
        WT_LAUNCHED();
        return 0;
    }
    post_kernel<<<n, kPostThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
