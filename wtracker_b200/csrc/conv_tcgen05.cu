// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA, fused epilogue (bias + SiLU + residual + bf16/f32 cast) stored by TMA
// straight into the (concat-)destination channel slice.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel @ tap, cin] * W[cout, tap, cin]
//
//   M tile  = 128 output pixels = a (tn x th x tw) patch of the NHWC output   (UMMA M = 128)
//   N tile  = BN output channels                                              (UMMA N = BN)
//   K block = BK input channels of one filter tap                             (UMMA K = 16)
//
// For every (tap, cin-block) the producer issues ONE 4-D TMA box load of the input shifted by the
// tap offset; out-of-image pixels are zero-filled by TMA, which is exactly the conv zero padding,
// and the box lands in smem as [128 pixels][BK channels] with the 128B/64B swizzle the UMMA
// K-major descriptor expects.  Stride-2 convs use four parity views of the input (even/odd rows x
// even/odd columns), so each tap is again a dense box.
//
// Warp roles (608 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; it
// also owns the TMEM allocation), warp 2 = spare (addend loader, second UMMA chain of the chained forms), warps 3..10 / 11..18 = two epilogue groups of eight warps (TMEM -> registers -> swizzled smem -> TMA
// store).  The TMEM accumulator is double-buffered: epilogue group g owns buffer g, i.e. every second tile of the CTA.
//
// Kernels in this file: conv_tc_kernel (generic: 1x1, stride-2 3x3, small maps), conv_halo_kernel (3x3 stride 1 with
// shared-memory halo reuse; stride-2 pixel-pair form for 32 input channels), conv0_tc_kernel (the first layer: im2col
// rows built by threads).  Fused forms: dot head (wt_op.dot_off), upsampled addend (wt_op.add_buf), chained 1x1 conv
// (wt_op.chain_w_off), concat chain at a C2f exit (wt_op.cat_buf).
#include "../../include/wtracker_b200.h"
#include "conv.cuh"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace wt {

namespace {

constexpr int kTileM = 128;
constexpr int kEpiGroups = 2;           // epilogue warp groups; group g drains TMEM accumulator g (tiles it % 2 == g)
constexpr int kEpiThreads = 256;        // threads per epilogue group (8 warps = 4 TMEM lane quadrants x 2 column halves)
constexpr int kFirstEpiWarp = 3;        // warp 0 TMA producer, warp 1 MMA issuer, warp 2 spare (addend loader / second UMMA chain)
constexpr int kThreads = kFirstEpiWarp * 32 + kEpiGroups * kEpiThreads;
constexpr int kEpiBarrier = 1;          // named barrier ids kEpiBarrier + group
constexpr int kStageBufBytes = 16384;   // one epilogue staging buffer: 128 rows x 128 B
constexpr int kSmemBudget = 232448;   // 227 KB opt-in maximum per CTA
constexpr int kBarrierBytes = 512;
constexpr int kMaxCout = 512;           // bias vector kept in shared memory
constexpr int kMaxStages = 16;

struct ConvTcParams {
    CUtensorMap tmA[4];   // input views (index = row parity * 2 + col parity for stride 2, else [0])
    CUtensorMap tmB;      // weights [cout][k*k*cin]  (halo kernel: 3-D view [cout][tap][cin])
    CUtensorMap tmD;      // output slice
    CUtensorMap tmR;      // residual slice (bf16)
    const float* bias;
    int tiles_x, tiles_y, tiles_n;   // pixel-tile grid (CTA pairs: tiles_x counts PAIRS of x-adjacent tiles)
    int n_blocks;                    // cout / BN
    int tw, th, tn;                  // pixel patch, tw*th*tn == 128
    int ksize, stride;
    int cin, cin_blocks;             // cin / BK
    int cout;
    int src_coff, dst_coff, res_coff;
    int act, has_res, out_f32;
    int num_tiles;
    // half-resolution f32 pre-activation addend (wt_op.add_buf): per tile the spare warp 2 TMA-loads the
    // (tw/2 x th/2 x tn)-pixel patch of the addend map into shared memory, one buffer per epilogue group
    CUtensorMap tmP;
    int has_add, add_coff;
    // fused 1-channel 1x1 head (wt_op.dot_off): out pixel = sum_c act(conv)[c] * dot_w[c] + dot_w[cout]
    const float* dot_w;
    float* dot_out;                  // f32 [n][out_h][out_w]
    int out_w, out_h, n_images;
    // shared-memory plan (host-chosen): pipeline depth and epilogue staging buffers per group (1 | 2).
    // HBM-bound layers (1x1, narrow N) want two staging buffers per epilogue group, MMA-bound layers
    // want the bytes as pipeline stages instead.
    int stages;        // generic kernel: A+B stages; halo kernel: weight (B) stages
    int a_stages;      // halo kernel: halo-tile stages
    int epi_bufs;
    // halo kernel: the layer's whole weight set (9 taps x one 64-channel block x all cout) fits in the B
    // stages, so it is loaded ONCE per CTA and stays resident: no weight re-streaming from L2 per tile
    int resident;
    // chained 1x1 conv (ConvDesc.chain_w): W2 [BN][BN] stays resident in shared memory as BN / 64 K blocks, the
    // epilogue's bf16 staging tiles are the A operand of a second UMMA chain (issued by warp 2) into accumulators
    // 2, 3 (TMEM columns 2 * BN ...), and a second epilogue pass applies bias2 / act2 and stores
    CUtensorMap tmB2;
    const float* bias2;
    int chain, act2;   // chain: 0 none | 1 chained 1x1 (cout -> cout) | 2 concat chain (below)
    // concat chain (ConvDesc.cat, C2f exit): the second GEMM is a 1x1 conv over concat(cat slice [64 ch], this conv's
    // output [32 ch]) -> 64 channels.  The cat tile of the pixel patch is TMA-loaded by warp 2 (tmY; it also supplies
    // the residual, which is its upper 32 channels), W2 = [64][96] is resident as a 64-channel K block (tmB2,
    // SWIZZLE_128B) + a 32-channel K block (tmB2b, SWIZZLE_64B like the staging tile it multiplies).
    CUtensorMap tmY, tmB2b;
    int cat_coff;
    // first layer on the tensor cores (conv0_tc_kernel): u8 grey input map [n][in_h][in_w]
    CUtensorMap tmIn;
    uint32_t mg_nb, mg_tx, mg_ty;   // multipliers for the tile-index divisions by n_blocks, tiles_x, tiles_y (fast_div)
    // halo kernel: pixel tiles per weight pass (NT template parameter).  A work item is a GROUP of nt tiles at the same
    // patch position of nt consecutive images; every weight tile streamed from L2 feeds the MMAs of all of them (the
    // 3x3 layers are bound by the L2 -> shared-memory fill, ~86 % of it weights: DESIGN.md), nt accumulators sit side
    // by side in TMEM and an epilogue group drains all of them.
    int nt;
    // Split epilogue (plain / residual / addend convs whose work item has >= 2 staging units): BOTH epilogue groups drain
    // every work item, unit su of item `it` going to group (su + it) & 1, instead of alternating whole items.  Same
    // epilogue throughput, half the latency per item: the drain after a CTA's last MMA — tensor pipe idle — halves.
    // tempty / add_empty then collect 16 warp arrivals and accumulator set it & 1 belongs to the item, not to a group.
    int split;
    // tuning builds: per-CTA event trace (wt_debug_conv_trace): [launch][CTA][4 roles][trace_cap] u64, or nullptr
    unsigned long long* trace;
    int trace_cap;
    // halo kernel: images interleaved per tile (IL template parameter, 1 | 2).  il == 2: the output / residual tensor maps
    // list (channel, x, image, row), the tile is tw x th x 2 with th = 8.
    int il;
};

template <int BN, int BK>
struct SmemLayout {
    static constexpr int kRowBytes = BK * 2;
    static constexpr int kABytes = kTileM * kRowBytes;
    static constexpr int kBBytes = BN * kRowBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
};

// bytes after the pipeline stages: epilogue staging + bias vector + barriers
__host__ __device__ constexpr int fixed_smem_bytes(int epi_bufs) {
    return kEpiGroups * epi_bufs * kStageBufBytes + kMaxCout * 4 + kBarrierBytes;
}
// addend patches between the staging buffers and the bias vector (generic kernel only): per epilogue group
// 32 low-res pixels x BN f32
__host__ __device__ constexpr int add_smem_bytes(int bn) { return kEpiGroups * bn * 128; }

// SiLU(v) = v * sigmoid(v) = h * tanh(h) + h with h = v / 2: ONE MUFU op (tanh.approx) per element
// instead of two (ex2 + rcp) — the epilogue warps are MUFU/issue bound, not the tensor pipe.  Measured on
// B200 against the fp32 oracle (tests/tools/gpu_detector_check.py): feature and logit errors are identical to
// the ex2+rcp form (both are dominated by the bf16 rounding of the stored activations) and the whole
// forward is 7 % faster.  WT_SILU_EXACT=1 selects the ex2+rcp form for A/B runs.
constexpr int kActSiluTanh = 2;
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Event trace of one warp role of one CTA (tuning builds only; compiled to nothing in the product library).  Slot 0 holds
// %globaltimer and slot 1 the SM's clock64 at the same moment (aligns the CTAs of a launch and the launches of a
// forward); every later slot is tag << 56 | value << 40 | clock64 (40 bits).  tools/gpu_conv_trace.py captures one forward,
// tools/conv_trace_report.py turns it into a per-launch table (profiles/r02_conv_trace.txt).
#ifdef WT_TUNING_KNOBS
struct Tracer {
    unsigned long long* p;
    int n, cap;
    __device__ __forceinline__ Tracer(const ConvTcParams& prm, int role, bool writer) {
        p = (prm.trace && writer) ? prm.trace + (size_t(blockIdx.x) * 4 + role) * prm.trace_cap : nullptr;
        n = 0;
        cap = prm.trace_cap;
        if (p) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p[0] = gt;
            p[1] = static_cast<unsigned long long>(clock64());
            n = 2;
        }
    }
    __device__ __forceinline__ void ev(uint32_t tag, uint32_t val = 0) {
        if (p && n < cap)
            p[n++] = (static_cast<unsigned long long>(tag) << 56) | (static_cast<unsigned long long>(val & 0xFFFFu) << 40) |
                     (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFull);
    }
    __device__ __forceinline__ long long now() const { return p ? clock64() : 0; }
    __device__ __forceinline__ void ev_smid(uint32_t tag) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        ev(tag, smid);
    }
};
#else
struct Tracer {
    __device__ __forceinline__ Tracer(const ConvTcParams&, int, bool) {}
    __device__ __forceinline__ void ev(uint32_t, uint32_t = 0) {}
    __device__ __forceinline__ long long now() const { return 0; }
    __device__ __forceinline__ void ev_smid(uint32_t) {}
};
#endif

// Work item t -> (N block, pixel patch origin).
struct TileCoord {
    int nblk, x0, y0, n0;
};
// n / d with the host-computed multiplier floor(2^32 / d) + 1 (exact while n * d < 2^32; 0 = fall back to a division).
// A runtime integer division is ~35 instructions including XU-pipe conversions and a MUFU.RCP; the small-tile kernels
// decode a tile every ~1500 clk in each role's loop, where three of them were 10-20 % of the critical path.
__device__ __forceinline__ int fast_div(int n, int d, uint32_t magic) {
    if (magic) return int(__umulhi(uint32_t(n), magic));
    return d == 1 ? n : n / d;
}
__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, int t) {
    TileCoord c;
    int m = fast_div(t, p.n_blocks, p.mg_nb);
    c.nblk = t - m * p.n_blocks;
    int q = fast_div(m, p.tiles_x, p.mg_tx);
    int xb = m - q * p.tiles_x;
    const int nb = fast_div(q, p.tiles_y, p.mg_ty);
    const int yb = q - nb * p.tiles_y;
    c.x0 = xb * p.tw;
    c.y0 = yb * p.th;
    c.n0 = nb * p.tn * p.nt;
    return c;
}

// Epilogue: two groups of 8 warps; group g owns TMEM accumulator g and therefore every second tile of this
// CTA, so the latency chain of one tile (TMEM load -> bias/SiLU/residual -> swizzled staging smem -> TMA
// store) overlaps the chain of the next tile as well as the MMA main loop.  Inside a group, warp pair
// (q, h) drains TMEM lane quadrant q (accumulator rows 32q..32q+31 = 32 pixels of the tile) and column
// half h of every staging unit: the epilogue is latency bound (ncu, round 1: 17 % issue utilisation with
// one warp per quadrant, 'wait' the top stall, every 1x1 / narrow-N layer limited by it and not by HBM or
// the tensor pipe), so each accumulator row is split over two threads.  sBias holds the layer's whole bias
// vector, pre-multiplied by 1/2 for the tanh form of SiLU (h = acc/2 + bias/2 is one FFMA).

// Fused class-logit head (wt_op.dot_off): thread (row, h) sums its half of the pixel's channels, the halves meet
// in shared memory.
template <int BN, int NT>
__device__ __forceinline__ void conv_epilogue_dot(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                                  uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base,
                                                  int warp, int lane) {
    const int ew = warp - kFirstEpiWarp;
    const int g = ew >> 3, h = (ew >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* part = reinterpret_cast<float*>(sStageAll + g * p.epi_bufs * kStageBufBytes);   // [2 slots][128]
    const float* dw = sBias + p.cout;
    const int bar_id = kEpiBarrier + g;
    int it = g;
    const int first = blockIdx.x, step = gridDim.x;
    ptx::grid_dependency_wait();
    for (int tile = first + g * step; tile < p.num_tiles; tile += 2 * step, it += 2) {
        const TileCoord tc = decode_tile(p, tile);
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tfull_bar[g], aphase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < NT; ++sub) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (g * NT + sub) * BN + h * (BN / 2);
        const float* bias = sBias + h * (BN / 2);
        const float* dwh = dw + h * (BN / 2);
        float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {   // 16 columns per step, BN / 2 columns per thread
            uint32_t acc[16];
            ptx::tmem_ld_32x16(t_row + c * 16, acc);
            ptx::tmem_ld_wait();
            if (c == BN / 32 - 1 && sub == NT - 1) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&tempty_bar[g]);
                }
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b = *reinterpret_cast<const float4*>(bias + c * 16 + 4 * j4);
                const float4 w = *reinterpret_cast<const float4*>(dwh + c * 16 + 4 * j4);
                const float bb[4] = {b.x, b.y, b.z, b.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float v;
                    if (p.act == kActSiluTanh) {
                        const float hh = fmaf(__uint_as_float(acc[4 * j4 + e]), 0.5f, bb[e]);
                        v = fmaf(hh, tanh_fast(hh), hh);
                    } else {
                        v = __uint_as_float(acc[4 * j4 + e]) + bb[e];
                        if (p.act == WT_ACT_SILU) v = __fdividef(v, 1.0f + __expf(-v));
                    }
                    dot = fmaf(v, ww[e], dot);
                }
            }
        }
        float* slot = part + (aphase * NT + sub) * kTileM;
        if (h == 1) slot[row] = dot;
        ptx::bar_sync(bar_id, kEpiThreads);   // (the slot is rewritten two work items of this group later: one more barrier in between)
        if (h == 0) {
            const int px = tc.x0 + row % p.tw;
            int py, pn;
            if (p.il == 2) {   // row = (ty * 2 + img) * 8 + tx
                py = tc.y0 + (row >> 4);
                pn = tc.n0 + 2 * sub + ((row >> 3) & 1);
            } else {
                py = tc.y0 + (row / p.tw) % p.th;
                pn = tc.n0 + sub * p.tn + row / (p.tw * p.th);
            }
            if (px < p.out_w && py < p.out_h && pn < p.n_images)
                p.dot_out[(size_t(pn) * p.out_h + py) * p.out_w + px] = (dot + slot[row]) + dw[p.cout];
        }
        }
    }
}

// CW = accumulator columns per thread and staging unit: 32 (bf16 output, units of 64 channels = 128-byte staging
// rows) or 16 (f32 output: units of 32 channels = 128-byte rows; bf16 with BN == 32: one unit of 64-byte rows).
template <int BN, int CW, int NT>
__device__ __forceinline__ void conv_epilogue_cw(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                                 uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_bar_all,
                                                 uint32_t tmem_base, int warp, int lane,
                                                 const uint8_t* sAddAll, uint64_t* add_full, uint64_t* add_empty) {
    const int ew = warp - kFirstEpiWarp;
    const int g = ew >> 3;                  // epilogue group == accumulator buffer
    const int h = (ew >> 2) & 1;            // column half of every unit
    const int et = threadIdx.x - kFirstEpiWarp * 32 - g * kEpiThreads;   // 0..255 inside the group
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // accumulator row == pixel index inside the tile
    const bool store_thread = (et == 0);
    const bool two_bufs = p.epi_bufs == 2;
    uint8_t* sStage = sStageAll + g * p.epi_bufs * kStageBufBytes;
    uint64_t* res_bar = res_bar_all + 2 * g;
    const int bar_id = kEpiBarrier + g;
    constexpr int kUnitCh = 2 * CW;                       // channels per staging unit
    constexpr int kUnits = BN / kUnitCh;
    const bool rows64 = CW == 16 && !p.out_f32;           // bf16, BN == 32: 64-byte staging rows (SWIZZLE_64B)
    const uint32_t unit_bytes = rows64 ? kTileM * 64 : kTileM * 128;
    uint32_t unit_counter = 0;
    const bool split = p.split != 0;
    int it = split ? 0 : g;
    const int first = blockIdx.x, step = gridDim.x;
    const int tile_step = split ? step : 2 * step, it_step = split ? 1 : 2;
    Tracer tr(p, 2 + g, store_thread);
    ptx::grid_dependency_wait();   // residual loads and output stores come after the previous kernel
    tr.ev(1);
    for (int tile = first + (split ? 0 : g * step); tile < p.num_tiles; tile += tile_step, it += it_step) {
        const TileCoord tc = decode_tile(p, tile);
        const int nblk = tc.nblk, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
        const uint32_t aphase = (it >> 1) & 1;
        const int ab = split ? (it & 1) : g;      // accumulator set / addend buffer of this work item
        // last unit of this work item that this group handles (its TMEM reads and addend reads end there)
        const int su_last = !split ? NT * kUnits - 1
                                   : ((((NT * kUnits - 1) + it) & 1) == g ? NT * kUnits - 1 : NT * kUnits - 2);
        const float* bias = sBias + nblk * BN + h * CW;

        // Upsampled addend (wt_op.add_buf): warp 2 has TMA-loaded this tile's half-resolution patch into this group's
        // buffer as [BN / 32 slices][32 low-res pixels][32 f32] with the 128-byte swizzle; this thread's pixel
        // (lx, ly, ln) of the (tw, th, tn) patch reads low-res pixel (lx / 2, ly / 2, ln).
        const uint8_t* add_row = nullptr;
        int add_xr = 0;
        if (p.has_add) {
            const int lx = row % p.tw, ly = (row / p.tw) % p.th, ln = row / (p.tw * p.th);
            const int prow = (ln * (p.th >> 1) + (ly >> 1)) * (p.tw >> 1) + (lx >> 1);
            add_row = sAddAll + ab * (BN * 128) + prow * 128;
            add_xr = prow & 7;
            ptx::mbar_wait(&add_full[ab], aphase);
        }

        tr.ev(19, it);
        ptx::mbar_wait(&tfull_bar[ab], aphase);
        ptx::tc_fence_after();
        tr.ev(20, it);

#pragma unroll 1
        for (int su = 0; su < NT * kUnits; ++su) {
            if (split && ((su + it) & 1) != g) continue;        // the other group's unit
            const int sub = NT == 1 ? 0 : su / kUnits;          // pixel tile of the group (image n0 + sub)
            const int unit = NT == 1 ? su : su - sub * kUnits;
            const int n0s = n0 + sub * p.tn;
            const int cy = p.il == 2 ? n0s : y0, cn = p.il == 2 ? y0 : n0s;   // tensor-map order of the last two coordinates
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (ab * NT + sub) * BN + h * CW;
            const int sb = two_bufs ? (unit_counter & 1) : 0;
            uint8_t* stage_buf = sStage + sb * kStageBufBytes;
            uint32_t acc[CW];
            ptx::tmem_ld_cols(t_row + unit * kUnitCh, acc);
            // the TMA store that last read this staging buffer must have finished reading
            if (store_thread) {
                if (two_bufs) ptx::tma_store_wait_read<1>();
                else ptx::tma_store_wait_read<0>();
            }
            ptx::bar_sync(bar_id, kEpiThreads);
            if (p.has_res && store_thread) {
                ptx::mbar_expect_tx(&res_bar[sb], unit_bytes);
                ptx::tma_load_4d(stage_buf, &p.tmR, &res_bar[sb], p.res_coff + nblk * BN + unit * kUnitCh, x0, cy, cn);
            }
            ptx::tmem_ld_wait();
            if (su == su_last) {
                // all TMEM reads of this warp for this tile are done: hand the accumulator back to the MMA warp
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&tempty_bar[ab]);
                }
            }
            float v[CW];
            if (p.act == kActSiluTanh) {
                // SiLU(x) = hh * tanh(hh) + hh with hh = x / 2; sBias holds bias / 2.  Packed f32x2 FMAs (FFMA2): the
                // accumulator registers of tcgen05.ld and the LDS.128 bias words are already adjacent pairs.
                uint64_t hh[CW / 2];
                const uint64_t half2 = ptx::pack_f32x2(0.5f, 0.5f);
#pragma unroll
                for (int j4 = 0; j4 < CW / 4; ++j4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + unit * kUnitCh + 4 * j4);
                    hh[2 * j4 + 0] = ptx::ffma2(ptx::pack_u32x2(acc[4 * j4 + 0], acc[4 * j4 + 1]), half2, ptx::pack_f32x2(b.x, b.y));
                    hh[2 * j4 + 1] = ptx::ffma2(ptx::pack_u32x2(acc[4 * j4 + 2], acc[4 * j4 + 3]), half2, ptx::pack_f32x2(b.z, b.w));
                }
                if (p.has_add) {
                    const int col0 = unit * kUnitCh + h * CW;             // first channel of this thread inside the N block
                    const uint8_t* arow = add_row + (col0 >> 5) * 4096;   // 32-channel slice
                    const int chunk0 = (col0 & 31) >> 2;
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const float4 a = *reinterpret_cast<const float4*>(arow + (((chunk0 + j4) ^ add_xr) << 4));
                        hh[2 * j4 + 0] = ptx::ffma2(ptx::pack_f32x2(a.x, a.y), half2, hh[2 * j4 + 0]);
                        hh[2 * j4 + 1] = ptx::ffma2(ptx::pack_f32x2(a.z, a.w), half2, hh[2 * j4 + 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) {
                    float lo, hi;
                    ptx::unpack_f32x2(hh[j], lo, hi);
                    const uint64_t t2 = ptx::pack_f32x2(tanh_fast(lo), tanh_fast(hi));
                    ptx::unpack_f32x2(ptx::ffma2(hh[j], t2, hh[j]), v[2 * j], v[2 * j + 1]);
                }
            } else {
#pragma unroll
                for (int j4 = 0; j4 < CW / 4; ++j4) {   // accumulator + bias (bias read as LDS.128 broadcasts)
                    const float4 b = *reinterpret_cast<const float4*>(bias + unit * kUnitCh + 4 * j4);
                    v[4 * j4 + 0] = __uint_as_float(acc[4 * j4 + 0]) + b.x;
                    v[4 * j4 + 1] = __uint_as_float(acc[4 * j4 + 1]) + b.y;
                    v[4 * j4 + 2] = __uint_as_float(acc[4 * j4 + 2]) + b.z;
                    v[4 * j4 + 3] = __uint_as_float(acc[4 * j4 + 3]) + b.w;
                }
                if (p.has_add) {
                    const int col0 = unit * kUnitCh + h * CW;
                    const uint8_t* arow = add_row + (col0 >> 5) * 4096;
                    const int chunk0 = (col0 & 31) >> 2;
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const float4 a = *reinterpret_cast<const float4*>(arow + (((chunk0 + j4) ^ add_xr) << 4));
                        v[4 * j4 + 0] += a.x;
                        v[4 * j4 + 1] += a.y;
                        v[4 * j4 + 2] += a.z;
                        v[4 * j4 + 3] += a.w;
                    }
                }
                if (p.act == WT_ACT_SILU) {
                    // v * sigmoid(v) with ex2.approx + rcp.approx (2 MUFU): relative error ~1e-6 everywhere.
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
                }
            }
            if (p.has_add && su == su_last) {   // this warp is done with the patch: warp 2 may load the next one into it
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&add_empty[ab]);
            }
            if (p.has_res) ptx::mbar_wait(&res_bar[sb], (two_bufs ? (unit_counter >> 1) : unit_counter) & 1);

            if (CW == 16 && p.out_f32) {
                // 32 f32 = 128 B per row, 8 chunks of 16 B, SWIZZLE_128B; this thread's half = 4 chunks
                uint8_t* rowp = stage_buf + row * 128;
#pragma unroll
                for (int c = 0; c < CW / 4; ++c) {
                    float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    *reinterpret_cast<float4*>(rowp + (((h * 4 + c) ^ (row & 7)) << 4)) = o;
                }
            } else {
                // CW bf16 = CW / 8 chunks of 16 B at chunk h * CW / 8 of the row
                uint8_t* rowp;
                int xr;
                if (CW == 16) {
                    rowp = stage_buf + row * 64;
                    xr = (row >> 1) & 3;
                } else {
                    rowp = stage_buf + row * 128;
                    xr = row & 7;
                }
#pragma unroll
                for (int c = 0; c < CW / 8; ++c) {
                    uint4* dstp = reinterpret_cast<uint4*>(rowp + (((h * (CW / 8) + c) ^ xr) << 4));
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = v[8 * c + j];
                    if (p.has_res) {
                        const uint4 r = *dstp;
                        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            f[2 * j] += __uint_as_float(rw[j] << 16);
                            f[2 * j + 1] += __uint_as_float(rw[j] & 0xFFFF0000u);
                        }
                    }
                    uint4 o;
                    o.x = pack_bf16(f[0], f[1]);
                    o.y = pack_bf16(f[2], f[3]);
                    o.z = pack_bf16(f[4], f[5]);
                    o.w = pack_bf16(f[6], f[7]);
                    *dstp = o;
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::bar_sync(bar_id, kEpiThreads);
            if (store_thread) {
                ptx::tma_store_4d(&p.tmD, stage_buf, p.dst_coff + nblk * BN + unit * kUnitCh, x0, cy, cn);
                ptx::tma_store_commit();
            }
            ++unit_counter;
        }
        tr.ev(21, it);
    }
    if (store_thread) ptx::tma_store_wait<0>();
    tr.ev(22);
}

// Chained form (ConvTcParams.chain, n_blocks == 1, bf16 in / out, BN = 64 | 128).  Per tile the group
//   1. drains accumulator g, applies bias / act and writes the bf16 tile into its BN / 64 staging buffers, which
//      are exactly the K-major SWIZZLE_128B A operand of a 128 x BN x BN GEMM;
//   2. signals a2_full[g]; warp 2 issues the second UMMA chain (B = W2, resident) into accumulator 2 + g;
//   3. waits t2full[g], drains accumulator 2 + g with bias2 / act2 into the same staging buffers (the MMAs that
//      read them have completed) and TMA-stores them.
// cb = chain barriers: a2_full[2], t2full[2], t2empty[2].
template <int BN>
__device__ __forceinline__ void conv_epilogue_chain(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                                    uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* cb,
                                                    uint32_t tmem_base, int warp, int lane) {
    constexpr int CW = 32;
    constexpr int kUnits = BN / 64;
    const int ew = warp - kFirstEpiWarp;
    const int g = ew >> 3, h = (ew >> 2) & 1;
    const int et = threadIdx.x - kFirstEpiWarp * 32 - g * kEpiThreads;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool store_thread = (et == 0);
    uint8_t* sStage = sStageAll + g * p.epi_bufs * kStageBufBytes;
    uint64_t* a2_full = cb + g;
    uint64_t* t2full = cb + 2 + g;
    uint64_t* t2empty = cb + 4 + g;
    const int bar_id = kEpiBarrier + g;
    const int xr = row & 7;
    int it = g;
    const int first = blockIdx.x, step = gridDim.x;
    ptx::grid_dependency_wait();
    for (int tile = first + g * step; tile < p.num_tiles; tile += 2 * step, it += 2) {
        const TileCoord tc = decode_tile(p, tile);
        const uint32_t aphase = (it >> 1) & 1;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * CW;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const float* bias = sBias + pass * BN + h * CW;
            const int act = pass ? p.act2 : p.act;
            if (pass == 0) {
                ptx::mbar_wait(&tfull_bar[g], aphase);
                // the stores of this group's previous tile must have finished reading the staging buffers
                if (store_thread) ptx::tma_store_wait_read<0>();
            } else {
                ptx::mbar_wait(t2full, aphase);
            }
            ptx::tc_fence_after();
            ptx::bar_sync(bar_id, kEpiThreads);
#pragma unroll 1
            for (int unit = 0; unit < kUnits; ++unit) {
                uint32_t acc[CW];
                ptx::tmem_ld_cols(t_row + (pass ? 2 * BN : 0) + g * BN + unit * 64, acc);
                ptx::tmem_ld_wait();
                if (unit == kUnits - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(pass ? t2empty : &tempty_bar[g]);
                }
                float v[CW];
                if (act == kActSiluTanh) {
                    const uint64_t half2 = ptx::pack_f32x2(0.5f, 0.5f);
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + unit * 64 + 4 * j4);
                        uint64_t h0 = ptx::ffma2(ptx::pack_u32x2(acc[4 * j4 + 0], acc[4 * j4 + 1]), half2, ptx::pack_f32x2(b.x, b.y));
                        uint64_t h1 = ptx::ffma2(ptx::pack_u32x2(acc[4 * j4 + 2], acc[4 * j4 + 3]), half2, ptx::pack_f32x2(b.z, b.w));
                        float l0, l1, l2, l3;
                        ptx::unpack_f32x2(h0, l0, l1);
                        ptx::unpack_f32x2(h1, l2, l3);
                        ptx::unpack_f32x2(ptx::ffma2(h0, ptx::pack_f32x2(tanh_fast(l0), tanh_fast(l1)), h0), v[4 * j4 + 0], v[4 * j4 + 1]);
                        ptx::unpack_f32x2(ptx::ffma2(h1, ptx::pack_f32x2(tanh_fast(l2), tanh_fast(l3)), h1), v[4 * j4 + 2], v[4 * j4 + 3]);
                    }
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + unit * 64 + 4 * j4);
                        v[4 * j4 + 0] = __uint_as_float(acc[4 * j4 + 0]) + b.x;
                        v[4 * j4 + 1] = __uint_as_float(acc[4 * j4 + 1]) + b.y;
                        v[4 * j4 + 2] = __uint_as_float(acc[4 * j4 + 2]) + b.z;
                        v[4 * j4 + 3] = __uint_as_float(acc[4 * j4 + 3]) + b.w;
                    }
                    if (act == WT_ACT_SILU) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
                    }
                }
                uint8_t* rowp = sStage + unit * kStageBufBytes + row * 128;
#pragma unroll
                for (int c = 0; c < CW / 8; ++c) {
                    uint4 o;
                    o.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
                    o.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
                    o.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
                    o.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
                    *reinterpret_cast<uint4*>(rowp + (((h * 4 + c) ^ xr) << 4)) = o;
                }
            }
            ptx::fence_proxy_async_smem();   // the tile is read by the async proxy: UMMA (pass 0) or the TMA store (pass 1)
            ptx::bar_sync(bar_id, kEpiThreads);
            if (store_thread) {
                if (pass == 0) {
                    ptx::mbar_arrive(a2_full);
                } else {
                    for (int unit = 0; unit < kUnits; ++unit)
                        ptx::tma_store_4d(&p.tmD, sStage + unit * kStageBufBytes, p.dst_coff + unit * 64, tc.x0, tc.y0, tc.n0);
                    ptx::tma_store_commit();
                }
            }
        }
    }
    if (store_thread) ptx::tma_store_wait<0>();
}

// Second UMMA chain of the chained form, one elected lane of warp 2: A = the epilogue group's staging tiles,
// B = W2 (resident, BN / 64 K blocks of [BN rows][128 B]), D = accumulator 2 + g.
template <int BN>
__device__ __forceinline__ void conv_chain_issuer(const ConvTcParams& p, const uint8_t* sStageAll, const uint8_t* sW2,
                                                  uint64_t* cb, uint64_t* w2_full, uint32_t tmem_base) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
    const int first = blockIdx.x, step = gridDim.x;
    const uint64_t b_desc0 = ptx::make_kmajor_desc<128>(ptx::smem_u32(sW2));
    const uint32_t b_hi = uint32_t(b_desc0 >> 32), b_lo0 = uint32_t(b_desc0);
    ptx::mbar_wait(w2_full, 0);
    ptx::tc_fence_after();
    int it = 0;
    for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
        const int g = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const uint64_t a_desc0 = ptx::make_kmajor_desc<128>(ptx::smem_u32(sStageAll + g * p.epi_bufs * kStageBufBytes));
        const uint32_t a_hi = uint32_t(a_desc0 >> 32), a_lo0 = uint32_t(a_desc0);
        ptx::mbar_wait(cb + 4 + g, ph ^ 1);   // t2empty: the group drained accumulator 2 + g (two tiles ago)
        ptx::mbar_wait(cb + g, ph);           // a2_full: the bf16 tile is in the staging buffers
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + 2 * BN + g * BN;
#pragma unroll
        for (int kb = 0; kb < BN / 64; ++kb)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ptx::umma_bf16_lohi(d_tmem, a_lo0 + kb * (kStageBufBytes >> 4) + 2 * kk, a_hi,
                                    b_lo0 + kb * ((BN * 128) >> 4) + 2 * kk, b_hi, idesc, (kb | kk) != 0);
        ptx::umma_commit(cb + 2 + g);         // t2full
    }
}

// ------------------------------------------------------------------------------------------------------------
// Concat chain (ConvTcParams.chain == 2): the exit of a C2f block with one bottleneck and 32 hidden channels,
//   b   = y1 + SiLU(conv3x3(b1))                  (this launch's main GEMM, BN = 32, residual y1)
//   out = SiLU(W2 * concat(y0, y1, b) + bias2)    (C2f.cv2, 96 -> 64)
// as one launch: b never leaves shared memory and y = [y0 | y1] is read once (it is both the residual and two
// thirds of the second GEMM's A operand).  Shared memory behind the staging buffers: Y[2] (one 128 px x 64 ch tile
// per epilogue group), W2a [64][64], W2b [64][32].  Staging buffer 0 of a group holds the bf16 b tile (64-byte rows,
// SWIZZLE_64B = K-major A operand of a K = 32 GEMM), buffer 1 the output tile.  cb: a2_full[2], t2full[2],
// t2empty[2]; yb: y_full[2].
constexpr int kCatC = 64, kCatOut = 64;
constexpr int kCatSmemBytes = 2 * kStageBufBytes + kCatOut * kCatC * 2 + kCatOut * 32 * 2;   // Y[2] + W2a + W2b

__device__ __forceinline__ void conv_epilogue_cat(const ConvTcParams& p, uint8_t* sStageAll, const uint8_t* sY,
                                                  const float* sBias, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                  uint64_t* cb, uint64_t* yb, uint32_t tmem_base, int warp, int lane) {
    constexpr int BN = 32;
    const int ew = warp - kFirstEpiWarp;
    const int g = ew >> 3, h = (ew >> 2) & 1;
    const int et = threadIdx.x - kFirstEpiWarp * 32 - g * kEpiThreads;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool store_thread = (et == 0);
    uint8_t* sB8 = sStageAll + g * 2 * kStageBufBytes;          // b tile, 64-byte rows
    uint8_t* sOut = sB8 + kStageBufBytes;                       // output tile, 128-byte rows
    const uint8_t* yrow = sY + g * kStageBufBytes + row * 128;
    const int bar_id = kEpiBarrier + g;
    int it = g;
    const int first = blockIdx.x, step = gridDim.x;
    ptx::grid_dependency_wait();
    for (int tile = first + g * step; tile < p.num_tiles; tile += 2 * step, it += 2) {
        const TileCoord tc = decode_tile(p, tile);
        const uint32_t aphase = (it >> 1) & 1;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // ---- pass 0: b = y1 + act(acc + bias), bf16, into the K = 32 staging tile
        ptx::mbar_wait(&tfull_bar[g], aphase);
        ptx::mbar_wait(&yb[g], aphase);
        if (store_thread) ptx::tma_store_wait_read<0>();   // the previous tile's output store has left sOut
        ptx::tc_fence_after();
        {
            uint32_t acc[16];
            ptx::tmem_ld_32x16(t_lane + g * BN + h * 16, acc);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[g]);
            const float* bias = sBias + h * 16;
            float v[16];
            if (p.act == kActSiluTanh) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float hh = fmaf(__uint_as_float(acc[j]), 0.5f, bias[j]);
                    v[j] = fmaf(hh, tanh_fast(hh), hh);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = __uint_as_float(acc[j]) + bias[j];
                    if (p.act == WT_ACT_SILU) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
                }
            }
            uint8_t* rowp = sB8 + row * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                // residual = channels 32 + 16 h + 8 c .. of the cat tile: 16-byte chunk 4 + 2 h + c of its 128-byte row
                const uint4 r = *reinterpret_cast<const uint4*>(yrow + (((4 + 2 * h + c) ^ (row & 7)) << 4));
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
                float f[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    f[2 * j] = v[8 * c + 2 * j] + __uint_as_float(rw[j] << 16);
                    f[2 * j + 1] = v[8 * c + 2 * j + 1] + __uint_as_float(rw[j] & 0xFFFF0000u);
                }
                uint4 o;
                o.x = pack_bf16(f[0], f[1]);
                o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]);
                o.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4*>(rowp + (((2 * h + c) ^ ((row >> 1) & 3)) << 4)) = o;
            }
        }
        ptx::fence_proxy_async_smem();
        ptx::bar_sync(bar_id, kEpiThreads);
        if (store_thread) ptx::mbar_arrive(&cb[g]);          // a2_full
        // ---- pass 1: out = act2(acc2 + bias2)
        ptx::mbar_wait(&cb[2 + g], aphase);                  // t2full
        ptx::tc_fence_after();
        {
            uint32_t acc[32];
            ptx::tmem_ld_32x32(t_lane + 2 * BN + g * kCatOut + h * 32, acc);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&cb[4 + g]);     // t2empty
            const float* bias = sBias + BN + h * 32;
            float v[32];
            if (p.act2 == kActSiluTanh) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float hh = fmaf(__uint_as_float(acc[j]), 0.5f, bias[j]);
                    v[j] = fmaf(hh, tanh_fast(hh), hh);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = __uint_as_float(acc[j]) + bias[j];
                    if (p.act2 == WT_ACT_SILU) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
                }
            }
            uint8_t* rowp = sOut + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 o;
                o.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
                o.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
                o.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
                o.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
                *reinterpret_cast<uint4*>(rowp + (((4 * h + c) ^ (row & 7)) << 4)) = o;
            }
        }
        ptx::fence_proxy_async_smem();
        ptx::bar_sync(bar_id, kEpiThreads);
        if (store_thread) {
            ptx::tma_store_4d(&p.tmD, sOut, p.dst_coff, tc.x0, tc.y0, tc.n0);
            ptx::tma_store_commit();
        }
    }
    if (store_thread) ptx::tma_store_wait<0>();
}

// Warp 2, one elected lane: loads the cat tiles one tile ahead and issues the second UMMA chain
// (4 K slices of the cat tile x W2a + 2 K slices of the b tile x W2b) into accumulator 2 + g.
__device__ __forceinline__ void conv_cat_issuer(const ConvTcParams& p, const uint8_t* sStageAll, const uint8_t* sY,
                                                const uint8_t* sW2a, const uint8_t* sW2b, uint64_t* cb, uint64_t* yb,
                                                uint64_t* w2_full, uint32_t tmem_base) {
    constexpr int BN = 32;
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, kCatOut);
    const int first = blockIdx.x, step = gridDim.x;
    const uint64_t ba = ptx::make_kmajor_desc<128>(ptx::smem_u32(sW2a));
    const uint64_t bb = ptx::make_kmajor_desc<64>(ptx::smem_u32(sW2b));
    ptx::prefetch_tmap(&p.tmY);
    ptx::grid_dependency_wait();           // the cat buffer is written by earlier kernels
    if (first < p.num_tiles) {
        const TileCoord tc = decode_tile(p, first);
        ptx::mbar_expect_tx(&yb[0], kStageBufBytes);
        ptx::tma_load_4d(const_cast<uint8_t*>(sY), &p.tmY, &yb[0], p.cat_coff, tc.x0, tc.y0, tc.n0);
    }
    ptx::mbar_wait(w2_full, 0);
    int it = 0;
    for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
        const int g = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        if (tile + step < p.num_tiles) {   // next tile's cat tile into the other buffer, once its previous user is done
            if (it >= 1) ptx::mbar_wait(&cb[2 + (g ^ 1)], ((it - 1) >> 1) & 1);   // t2full of tile it - 1
            const TileCoord tn = decode_tile(p, tile + step);
            ptx::mbar_expect_tx(&yb[g ^ 1], kStageBufBytes);
            ptx::tma_load_4d(const_cast<uint8_t*>(sY) + (g ^ 1) * kStageBufBytes, &p.tmY, &yb[g ^ 1], p.cat_coff, tn.x0,
                             tn.y0, tn.n0);
        }
        ptx::mbar_wait(&cb[4 + g], ph ^ 1);   // t2empty
        ptx::mbar_wait(&yb[g], ph);           // cat tile landed
        ptx::mbar_wait(&cb[g], ph);           // a2_full: b tile staged
        ptx::tc_fence_after();
        const uint64_t ay = ptx::make_kmajor_desc<128>(ptx::smem_u32(sY + g * kStageBufBytes));
        const uint64_t ab = ptx::make_kmajor_desc<64>(ptx::smem_u32(sStageAll + g * 2 * kStageBufBytes));
        const uint32_t d_tmem = tmem_base + 2 * BN + g * kCatOut;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            ptx::umma_bf16_lohi(d_tmem, uint32_t(ay) + 2 * kk, uint32_t(ay >> 32), uint32_t(ba) + 2 * kk, uint32_t(ba >> 32),
                                idesc, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
            ptx::umma_bf16_lohi(d_tmem, uint32_t(ab) + 2 * kk, uint32_t(ab >> 32), uint32_t(bb) + 2 * kk, uint32_t(bb >> 32),
                                idesc, true);
        ptx::umma_commit(&cb[2 + g]);         // t2full
    }
}

template <int BN, int NT = 1>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_bar_all,
                                              uint32_t tmem_base, int warp, int lane,
                                              const uint8_t* sAddAll = nullptr, uint64_t* add_full = nullptr,
                                              uint64_t* add_empty = nullptr, uint64_t* chain_bars = nullptr) {
    if (p.chain) {
        if constexpr ((BN == 64 || BN == 128) && NT == 1)
            conv_epilogue_chain<BN>(p, sStageAll, sBias, tfull_bar, tempty_bar, chain_bars, tmem_base, warp, lane);
        return;
    }
    if (p.dot_w) {
        conv_epilogue_dot<BN, NT>(p, sStageAll, sBias, tfull_bar, tempty_bar, tmem_base, warp, lane);
        return;
    }
    if constexpr (BN >= 64) {
        if (!p.out_f32) {
            conv_epilogue_cw<BN, 32, NT>(p, sStageAll, sBias, tfull_bar, tempty_bar, res_bar_all, tmem_base, warp, lane, sAddAll, add_full, add_empty);
            return;
        }
    }
    conv_epilogue_cw<BN, 16, NT>(p, sStageAll, sBias, tfull_bar, tempty_bar, res_bar_all, tmem_base, warp, lane,
                                     sAddAll, add_full, add_empty);
}

template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    using L = SmemLayout<BN, BK>;
    Tracer tr0(p, 0, threadIdx.x == 0);   // (tuning builds) slot 0 / 1 = kernel entry
    tr0.ev_smid(32);
    const int first = blockIdx.x, step = gridDim.x;
    const int kStages = p.stages;
    constexpr int kRowBytes = L::kRowBytes;
    constexpr uint32_t kTmemCols = BN > 128 ? 512 : 2 * BN;   // double-buffered accumulator, a power of two >= 64
    static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 192 || BN == 256, "BN");

    // 128-byte swizzle atoms need a 1024-byte aligned base: declared on the array (the dynamic window then
    // starts aligned, so no slack bytes are reserved) and checked once
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sA = smem;                                  // [stages][128][BK] bf16 (swizzled)
    uint8_t* sB = smem + kStages * L::kABytes;           // [stages][BN][BK]  bf16 (swizzled)
    uint8_t* sStage = smem + kStages * L::kStageBytes;   // 2 groups x epi_bufs x 16 KB epilogue staging
    // addend patches [2 groups][BN / 32][32 px][128 B], 1024-byte aligned like everything before them (128-byte swizzle)
    const uint8_t* sAdd = sStage + kEpiGroups * p.epi_bufs * kStageBufBytes;
    float* sBias = reinterpret_cast<float*>(const_cast<uint8_t*>(sAdd) + (p.has_add ? add_smem_bytes(BN) : 0) +
                                            (p.chain ? BN * BN * 2 : 0));   // [kMaxCout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);
    uint64_t* full_bar = bars;                          // [stages]  TMA -> MMA
    uint64_t* empty_bar = bars + kMaxStages;            // [stages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [2]       MMA -> epilogue group
    uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;   // [2]       epilogue group -> MMA
    uint64_t* res_bar = bars + 2 * kMaxStages + 4;      // [2][2]    residual TMA -> epilogue group
    uint64_t* add_full = bars + 2 * kMaxStages + 8;     // [2]       addend patch TMA -> epilogue group
    uint64_t* add_empty = bars + 2 * kMaxStages + 10;   // [2]       epilogue group -> addend loader
    uint64_t* chain_bars = bars + 2 * kMaxStages + 12;  // [6]       a2_full[2], t2full[2], t2empty[2] (chained form)
    uint64_t* w2_full = bars + 2 * kMaxStages + 18;     //           W2 resident
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 19);
    const uint8_t* sW2 = sAdd;                          // chained form: W2 where the addend patches would be

    // shfl makes the warp index provably warp-uniform, so the role branches below are uniform branches
    // and the producer / MMA loops can live on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.tmA[0]);
        ptx::prefetch_tmap(&p.tmB);
        ptx::prefetch_tmap(&p.tmD);
        if (p.has_res) ptx::prefetch_tmap(&p.tmR);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], p.split ? 16 : 8);   // the 8 warps of an epilogue group (split: of both)
        }
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&res_bar[i], 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&add_full[i], 1);
            ptx::mbar_init(&add_empty[i], p.split ? 16 : 8);
        }
        for (int i = 0; i < 6; ++i) ptx::mbar_init(&chain_bars[i], i < 4 ? 1 : 8);
        ptx::mbar_init(w2_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, p.chain ? 4u * BN : kTmemCols);
        ptx::tmem_relinquish();
    }
    // whole bias vector -> smem once per CTA
    {
        const float bscale = p.act == kActSiluTanh ? 0.5f : 1.0f;   // tanh form of SiLU works on (acc + bias) / 2
        for (int i = threadIdx.x; i < p.cout; i += kThreads) sBias[i] = bscale * __ldg(p.bias + i);
    }
    if (p.dot_w)   // dot weights + bias behind the conv bias (host checks 2 * cout + 1 <= kMaxCout)
        for (int i = threadIdx.x; i <= p.cout; i += kThreads) sBias[p.cout + i] = __ldg(p.dot_w + i);
    if (p.chain) {   // bias of the chained 1x1 conv behind the conv bias
        const float bscale2 = p.act2 == kActSiluTanh ? 0.5f : 1.0f;
        for (int i = threadIdx.x; i < p.cout; i += kThreads) sBias[p.cout + i] = bscale2 * __ldg(p.bias2 + i);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tr0.ev(30);
    // Programmatic dependent launch: everything above (barriers, TMEM, bias = constants) overlapped the tail of the
    // previous kernel; the next kernel may start its own prologue now.  Activations are only touched after
    // grid_dependency_wait() (producer and epilogue warps; the MMA warps never touch global memory).
    ptx::grid_launch_dependents();

    if (warp == 2 && p.has_add) {
        // ------------------------------------------------------------------ addend loader (the spare issuer warp)
        // tile it -> epilogue group it & 1; the group's buffer is refilled as soon as its four warps released it
        if (ptx::elect_one()) {
            ptx::prefetch_tmap(&p.tmP);
            ptx::grid_dependency_wait();
            int it = 0;
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int g = it & 1;
                const TileCoord tc = decode_tile(p, tile);
                ptx::mbar_wait(&add_empty[g], ((it >> 1) & 1) ^ 1);
                ptx::mbar_expect_tx(&add_full[g], BN * 128);
                for (int sub = 0; sub < BN / 32; ++sub)
                    ptx::tma_load_4d(const_cast<uint8_t*>(sAdd) + g * (BN * 128) + sub * 4096, &p.tmP, &add_full[g],
                                     p.add_coff + tc.nblk * BN + sub * 32, tc.x0 >> 1, tc.y0 >> 1, tc.n0);
            }
        }
        __syncwarp();
    }

    if (warp == 2 && p.chain == 1) {
        // ------------------------------------------------------------------ second UMMA chain (chained 1x1 conv)
        if constexpr ((BN == 64 || BN == 128)) {
            if (ptx::elect_one()) conv_chain_issuer<BN>(p, sStage, sW2, chain_bars, w2_full, tmem_base);
        }
        __syncwarp();
    }

    const int taps = p.ksize * p.ksize;
    const int num_kb = taps * p.cin_blocks;
    const int pad = p.ksize >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // The whole warp walks the loop (uniform control flow); one elected lane issues the copies.
        if (p.chain == 1 && ptx::elect_one()) {   // W2 is a constant: loaded before the grid dependency resolves
            ptx::prefetch_tmap(&p.tmB2);
            ptx::mbar_expect_tx(w2_full, BN * BN * 2);
            for (int kb = 0; kb < BN / 64; ++kb)
                ptx::tma_load_2d(const_cast<uint8_t*>(sW2) + kb * (BN * 128), &p.tmB2, w2_full, kb * 64, 0);
        }
        __syncwarp();
        Tracer& tr = tr0;
        ptx::grid_dependency_wait();
        tr.ev(1);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = first; tile < p.num_tiles; tile += step) {
            const TileCoord tc = decode_tile(p, tile);
            const int nblk = tc.nblk, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
            tr.ev(2);
            for (int tap = 0; tap < taps; ++tap) {
                const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
                int ax, ay, mapi;
                if (p.stride == 1) {
                    ax = x0 + kw - pad;
                    ay = y0 + kh - pad;
                    mapi = 0;
                } else {   // stride 2, 3x3, pad 1: input row 2*oy+kh-1 -> parity view + offset
                    ax = x0 + (kw == 0 ? -1 : 0);
                    ay = y0 + (kh == 0 ? -1 : 0);
                    mapi = ((kh != 1) ? 2 : 0) + ((kw != 1) ? 1 : 0);
                }
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                        ptx::tma_load_4d(sA + stage * L::kABytes, &p.tmA[mapi], &full_bar[stage],
                                         p.src_coff + cb * BK, ax, ay, n0);
                        ptx::tma_load_2d(sB + stage * L::kBBytes, &p.tmB, &full_bar[stage], tap * p.cin + cb * BK,
                                         nblk * BN);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            tr.ev(3);
        }
    } else if (warp < kFirstEpiWarp) {
        // ------------------------------------------------------------------ MMA issuer (warp 1; warp 2 idles here)
        // ONE elected lane per issuer warp runs the whole loop (waits included).  elect.sync tells the
        // compiler that a single thread is active, so descriptors and barrier addresses stay on the
        // uniform datapath; the descriptor low words advance by plain 32-bit adds (stage, K slice).
        // Issuer w handles tiles it = w, w + 2, ...; both walk the same smem ring, whose slot for the
        // g-th K block of the CTA is g % stages.
        const int w = warp - 1;
        if (w == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));
            const uint64_t b_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            int it = 0;
            Tracer tr(p, 1, true);
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);
                int stage = int(g0 % uint32_t(kStages));
                uint32_t phase = (g0 / uint32_t(kStages)) & 1u;
                uint32_t a_lo = a_lo0 + stage * (L::kABytes >> 4), b_lo = b_lo0 + stage * (L::kBBytes >> 4);
                tr.ev(10, it);
                ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                tr.ev(11, it);
                long long wf = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const long long t0 = tr.now();
                    ptx::mbar_wait(&full_bar[stage], phase);
                    wf += tr.now() - t0;
                    ptx::tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        // advance 16 bf16 = 32 B along K inside the swizzle span: start address field += 2
                        ptx::umma_bf16_lohi(d_tmem, a_lo + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc, (kb | kk) != 0);
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs finish
                    ptx::umma_commit(&empty_bar[stage]);
                    a_lo += L::kABytes >> 4;
                    b_lo += L::kBBytes >> 4;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                        a_lo = a_lo0;
                        b_lo = b_lo0;
                    }
                }
                ptx::umma_commit(&tfull_bar[ab]);
                tr.ev(13, uint32_t(wf >> 4));
                tr.ev(14, it);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 3..18)
        conv_epilogue<BN>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane, sAdd,
                              add_full, add_empty, chain_bars);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, p.chain ? 4u * BN : kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 / stride-1 variant with shared-memory halo reuse.
//
// The generic kernel above re-reads the input once per filter tap (9 TMA boxes per channel block),
// which makes the 3x3 layers L2-bandwidth bound.  Here the tile is 16 rows x 8 columns of one image
// and the producer loads the (16+2) x (8+2) halo of a 64-channel block ONCE; the nine taps are nine
// UMMA descriptors into that buffer: tap (kh, kw) starts (kh*10 + kw) pixels in, the 8 pixels of a
// tile row are 8 consecutive 128-byte rows (one swizzle group) and consecutive tile rows are one
// halo row (10 pixels = 1280 B) apart, which is the descriptor's stride-byte-offset.  The 128-byte
// swizzle is a function of the shared-memory address (verified on B200: no descriptor base offset
// is needed), so TMA (writer) and UMMA (reader) agree on it for any start pixel.  Layers with 32
// input channels use a 32-wide K block: 64-byte rows and SWIZZLE_64B for both operands (half the
// shared-memory traffic of zero-padding K to 64 — these layers are shared-memory-bandwidth bound).
// Weights stream through their own, deeper pipeline (one BN x BK tile per tap) or stay resident.
constexpr int kHaloW = 10, kHaloH = 18;
// Stride-2 form (S2 = 1, 32 input channels, e.g. layer 1): the input is viewed as PAIRS of x-adjacent pixels
// (2 x 32 channels = one 128-byte SWIZZLE_128B row).  The tile's 33 x 9 pair-row patch is loaded once; output
// pixel (ty, tx) at tap (kh, kw) reads input row 2*ty + kh of the patch and input x = 2*ox + kw - 1, i.e. the
// second half of pair tx (kw = 0), the first half of pair tx + 1 (kw = 1) or its second half (kw = 2): a
// descriptor start offset of (kh * 9 + (kw != 0)) rows + (kw != 1) * 64 bytes, 8 consecutive rows per tile row,
// and a stride of two patch rows (18 rows) between tile rows.  One fill instead of nine strided ones.
constexpr int kS2HaloW = 9, kS2HaloH = 33;
// bytes of one halo stage (1024-byte aligned): S2 = 0: K block of BK channels (rows of 2 * BK bytes)
// Interleaved form (IL = 2): the tile is 8 columns x 8 rows of TWO images, held in shared memory as [row][image][x]
// (the tensor maps list the image dimension before the row dimension), so accumulator row group 2 * ty + img is still
// kHaloW pixels after the previous one — the same descriptor stride — and tap (kh, kw) starts (kh * 2 * kHaloW + kw)
// pixels in.  Maps whose height is a multiple of 8 but not of 16 (40 x 40) tile exactly instead of wasting a fifth of
// every third tile row.
constexpr int kIlHaloRows = 10 * 2 * kHaloW;   // (8 + 2) rows x 2 images x 10 pixels
__host__ __device__ constexpr int halo_rows(int s2, int il = 1) {
    return s2 ? kS2HaloW * kS2HaloH : (il == 2 ? kIlHaloRows : kHaloW * kHaloH);
}
__host__ __device__ constexpr int halo_a_bytes(int bk, int s2, int il = 1) {
    return ((halo_rows(s2, il) * (s2 ? 128 : bk * 2) + 1023) / 1024) * 1024;
}

template <int BN, int BK, int S2, int IL = 1>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // weight rows: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kARowBytes = S2 ? 128 : BK * 2;     // halo rows (S2: a pixel pair)
    static constexpr int kABytes = halo_a_bytes(BK, S2, IL);
    static constexpr int kATxBytes = halo_rows(S2, IL) * kARowBytes;
    static constexpr int kBBytes = (BN) * kRowBytes;    // a CTA of a pair holds BN / 2 weight rows
};
constexpr int kMaxAStages = 4;

template <int BN, int BK, int S2, int NT, int IL>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    static_assert(!S2 || (BK == 32), "the stride-2 pair form is written for 32 input channels");
    static_assert(IL == 1 || (IL == 2 && !S2), "interleaved tiles: stride 1, single-CTA MMAs");
    static_assert(NT == 1 || (2 * NT * BN <= 512), "tile groups: single-CTA MMAs, 2 x NT accumulators in TMEM");
    using L = HaloSmem<BN, BK, S2, IL>;
    Tracer tr0(p, 0, threadIdx.x == 0);   // (tuning builds) slot 0 / 1 = kernel entry
    tr0.ev_smid(32);
    constexpr int kARowBytes = L::kARowBytes;
    constexpr int kTapRowPitch = kHaloW * IL;        // pixels between vertically adjacent taps in the halo tile
    const int first = blockIdx.x, step = gridDim.x;
    constexpr int kHaloABytes = L::kABytes;          // one halo tile
    constexpr int kAStageBytes = NT * kHaloABytes;   // one A stage = the halo tiles of a whole tile group
    constexpr int kRowBytes = L::kRowBytes;
    const int kAStages = p.a_stages, kBStages = p.stages;
    constexpr uint32_t kAccCols = 2 * NT * BN;       // two groups of NT accumulators (BN = 192: 384 columns used of 512)
    constexpr uint32_t kTmemCols = kAccCols > 256 ? 512 : (kAccCols > 128 ? 256 : (kAccCols > 64 ? 128 : 64));

    // 128-byte swizzle atoms need a 1024-byte aligned base: declared on the array (the dynamic window then
    // starts aligned, so no slack bytes are reserved) and checked once
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sA = smem;                                        // [kAStages][NT] halo tiles (180 px x 128 B, swizzled)
    uint8_t* sB = smem + kAStages * kAStageBytes;              // [kBStages][BN][64] bf16
    uint8_t* sStage = sB + kBStages * L::kBBytes;
    const uint8_t* sW2 = sStage + kEpiGroups * p.epi_bufs * kStageBufBytes;   // chained form: W2 [BN / 64][BN][64] bf16
    // concat chain: Y[2] cat tiles, then W2a [64][64] (128-byte rows), then W2b [64][32] (64-byte rows)
    const uint8_t* sY = sW2;
    const uint8_t* sW2a = sY + 2 * kStageBufBytes;
    const uint8_t* sW2b = sW2a + kCatOut * kCatC * 2;
    float* sBias = reinterpret_cast<float*>(const_cast<uint8_t*>(sW2) +
                                            (p.chain == 2 ? kCatSmemBytes : (p.chain ? BN * BN * 2 : 0)));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + kMaxAStages;
    uint64_t* bfull = aempty + kMaxAStages;
    uint64_t* bempty = bfull + kMaxStages;
    uint64_t* tfull_bar = bempty + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_bar = tempty_bar + 2;
    uint64_t* chain_bars = res_bar + 4;   // [6] a2_full[2], t2full[2], t2empty[2] (chained form)
    uint64_t* w2_full = res_bar + 10;
    uint64_t* y_full = res_bar + 11;      // [2] cat tile TMA -> epilogue group / second MMA chain (concat chain)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 13);
    constexpr bool kCatCapable = (BN == 32 && BK == 32 && S2 == 0 && NT == 1 && IL == 1);

    // shfl makes the warp index provably warp-uniform, so the role branches below are uniform branches
    // and the producer / MMA loops can live on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.tmA[0]);
        ptx::prefetch_tmap(&p.tmB);
        ptx::prefetch_tmap(&p.tmD);
        if (p.has_res) ptx::prefetch_tmap(&p.tmR);
        for (int s = 0; s < kAStages; ++s) {
            ptx::mbar_init(&afull[s], 1);
            ptx::mbar_init(&aempty[s], 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            ptx::mbar_init(&bfull[s], 1);
            ptx::mbar_init(&bempty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], p.split ? 16 : 8);   // the 8 warps of an epilogue group (split: of both)
        }
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&res_bar[i], 1);
        for (int i = 0; i < 6; ++i) ptx::mbar_init(&chain_bars[i], i < 4 ? 1 : 8);
        ptx::mbar_init(w2_full, 1);
        ptx::mbar_init(&y_full[0], 1);
        ptx::mbar_init(&y_full[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, p.chain == 2 ? 256u : (p.chain ? 4u * BN : kTmemCols));
        ptx::tmem_relinquish();
    }
    // whole bias vector -> smem once per CTA
    {
        const float bscale = p.act == kActSiluTanh ? 0.5f : 1.0f;   // tanh form of SiLU works on (acc + bias) / 2
        for (int i = threadIdx.x; i < p.cout; i += kThreads) sBias[i] = bscale * __ldg(p.bias + i);
    }
    if (p.dot_w)   // dot weights + bias behind the conv bias (host checks 2 * cout + 1 <= kMaxCout)
        for (int i = threadIdx.x; i <= p.cout; i += kThreads) sBias[p.cout + i] = __ldg(p.dot_w + i);
    if (p.chain) {   // bias of the chained 1x1 conv behind the conv bias
        const float bscale2 = p.act2 == kActSiluTanh ? 0.5f : 1.0f;
        const int n2 = p.chain == 2 ? kCatOut : p.cout;
        for (int i = threadIdx.x; i < n2; i += kThreads) sBias[p.cout + i] = bscale2 * __ldg(p.bias2 + i);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tr0.ev(30);
    // Programmatic dependent launch: everything above (barriers, TMEM, bias = constants) overlapped the tail of the
    // previous kernel; the next kernel may start its own prologue now.  Activations are only touched after
    // grid_dependency_wait() (producer and epilogue warps; the MMA warps never touch global memory).
    ptx::grid_launch_dependents();

    if (warp == 2 && p.chain) {
        // second UMMA chain of the chained 1x1 conv
        if constexpr ((BN == 64 || BN == 128) && NT == 1 && IL == 1) {
            if (p.chain == 1 && ptx::elect_one()) conv_chain_issuer<BN>(p, sStage, sW2, chain_bars, w2_full, tmem_base);
        }
        if constexpr (kCatCapable) {
            if (p.chain == 2 && ptx::elect_one())
                conv_cat_issuer(p, sStage, sY, sW2a, sW2b, chain_bars, y_full, w2_full, tmem_base);
        }
        __syncwarp();
    }

    if (warp == 0) {
        // TMA producer: warp-uniform loop, one elected lane issues
        if (p.chain == 1 && ptx::elect_one()) {   // W2 is a constant: loaded before the grid dependency resolves
            ptx::prefetch_tmap(&p.tmB2);
            ptx::mbar_expect_tx(w2_full, BN * BN * 2);
            for (int kb = 0; kb < BN / 64; ++kb)
                ptx::tma_load_2d(const_cast<uint8_t*>(sW2) + kb * (BN * 128), &p.tmB2, w2_full, kb * 64, 0);
        }
        if (p.chain == 2 && ptx::elect_one()) {
            ptx::prefetch_tmap(&p.tmB2);
            ptx::prefetch_tmap(&p.tmB2b);
            ptx::mbar_expect_tx(w2_full, kCatOut * (kCatC + 32) * 2);
            ptx::tma_load_2d(const_cast<uint8_t*>(sW2a), &p.tmB2, w2_full, 0, 0);
            ptx::tma_load_2d(const_cast<uint8_t*>(sW2b), &p.tmB2b, w2_full, kCatC, 0);
        }
        __syncwarp();
        Tracer& tr = tr0;
        ptx::grid_dependency_wait();
        tr.ev(1);
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile = first; tile < p.num_tiles; tile += step) {
            const TileCoord tc = decode_tile(p, tile);
            const int nblk = tc.nblk;
            for (int cb = 0; cb < p.cin_blocks; ++cb) {
                ptx::mbar_wait(&aempty[sa], pa ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&afull[sa], NT * L::kATxBytes);
#pragma unroll
                    for (int sub = 0; sub < NT; ++sub) {   // the halo tiles of the group: same patch, images n0 .. n0 + NT - 1
                        uint8_t* dst = sA + sa * kAStageBytes + sub * kHaloABytes;
                        if (S2)   // pair view: x in pairs (pair ox0 - 1 first), y in input rows (row 2*oy0 - 1 first)
                            ptx::tma_load_4d(dst, &p.tmA[0], &afull[sa], 0, tc.x0 - 1, 2 * tc.y0 - 1, tc.n0 + sub);
                        else if (IL == 2)   // (channel, x, image, row): two images per tile
                            ptx::tma_load_4d(dst, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, tc.x0 - 1,
                                             tc.n0 + 2 * sub, tc.y0 - 1);
                        else
                            ptx::tma_load_4d(dst, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, tc.x0 - 1, tc.y0 - 1,
                                             tc.n0 + sub);
                    }
                }
                __syncwarp();
                tr.ev(2, cb);
                if (++sa == kAStages) { sa = 0; pa ^= 1; }
                if (p.resident && tile != first) continue;   // weights already in shared memory
                for (int tap = 0; tap < 9; ++tap) {
                    ptx::mbar_wait(&bempty[sb], pb ^ 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&bfull[sb], L::kBBytes);
                        ptx::tma_load_3d(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * BK, tap, nblk * BN);
                    }
                    __syncwarp();
                    if (++sb == kBStages) { sb = 0; pb ^= 1; }
                }
                tr.ev(3, cb);
            }
        }
    } else if (warp < kFirstEpiWarp) {
        // MMA issuer (warp 1): one elected lane runs the whole loop;
        // descriptor low words advance by 32-bit adds (halo stage, weight stage, tap offset and K slice are all
        // additive in the start-address field).  Ring slots: halo tile g -> g % a_stages, weight tile g -> g % b_stages.
        const int w = warp - 1;
        if (w == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kARowBytes>(
                ptx::smem_u32(sA), S2 ? 2 * kS2HaloW * 128 : kHaloW * kARowBytes);
            const uint64_t b_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            int it = 0;
            Tracer tr(p, 1, true);
            if (kBStages == 9) {
                // Weight ring of exactly nine slots (resident weights, or the host chose 9 stages): slot == tap, so
                // every weight descriptor and barrier address in the unrolled tap loop is base + immediate and the
                // ring bookkeeping disappears from the issue loop (the issuing thread, not the tensor pipe, bounds
                // narrow-N layers).  Slot parity of channel block g of the CTA: g & 1.
                for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                    const int ab = it & 1;
                    const uint32_t d_tmem = tmem_base + ab * (NT * BN);
                    const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);
                    int sa = int(ga % uint32_t(kAStages));
                    uint32_t pa = (ga / uint32_t(kAStages)) & 1u;
                    uint32_t a_lo = a_lo0 + sa * (kAStageBytes >> 4);
                    const bool wait_b = !p.resident || it == 0;
                    tr.ev(10, it);
                    ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);
                    tr.ev(11, it);
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
                        const uint32_t pb = (ga + cb) & 1u;
                        ptx::mbar_wait(&afull[sa], pa);
                        ptx::tc_fence_after();
                        tr.ev(12, cb);
                        long long wb = 0;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (wait_b) {
                                const long long t0 = tr.now();
                                ptx::mbar_wait(&bfull[tap], pb);
                                wb += tr.now() - t0;
                                ptx::tc_fence_after();
                            }
                            const int kh = tap / 3, kw = tap - kh * 3;
                            const uint32_t a_tap =
                                a_lo + (S2 ? (((kh * kS2HaloW + (kw != 0 ? 1 : 0)) * 128 + (kw != 1 ? 64 : 0)) >> 4)
                                           : (((kh * kTapRowPitch + kw) * kARowBytes) >> 4));
                            const uint32_t b_tap = b_lo0 + tap * (L::kBBytes >> 4);
#pragma unroll
                            for (int sub = 0; sub < NT; ++sub) {   // the same weight tile feeds every pixel tile of the group
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) {
                                    ptx::umma_bf16_lohi(d_tmem + sub * BN, a_tap + sub * (kHaloABytes >> 4) + 2 * kk, a_hi,
                                                            b_tap + 2 * kk, b_hi, idesc, (cb | tap | kk) != 0);
                                }
                            }
                            if (!p.resident) {
                                ptx::umma_commit(&bempty[tap]);
                            }
                        }
                        ptx::umma_commit(&aempty[sa]);
                        tr.ev(13, uint32_t(wb >> 4));
                        a_lo += kAStageBytes >> 4;
                        if (++sa == kAStages) { sa = 0; pa ^= 1; a_lo = a_lo0; }
                    }
                    ptx::umma_commit(&tfull_bar[ab]);
                    tr.ev(14, it);
                }
            } else
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * (NT * BN);
                const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);
                int sa = int(ga % uint32_t(kAStages));
                uint32_t pa = (ga / uint32_t(kAStages)) & 1u;
                int sb = 0;
                uint32_t pb = 0;
                if (!p.resident) {
                    const uint32_t gb = ga * 9u;
                    sb = int(gb % uint32_t(kBStages));
                    pb = (gb / uint32_t(kBStages)) & 1u;
                }
                uint32_t a_lo = a_lo0 + sa * (kAStageBytes >> 4), b_lo = b_lo0 + sb * (L::kBBytes >> 4);
                // resident weights: waited for once, all nine taps (phase 0 of each slot)
                const bool wait_b = !p.resident || it == 0;   // (resident: the first tile of each issuer)
                tr.ev(10, it);
                ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);
                tr.ev(11, it);
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    ptx::mbar_wait(&afull[sa], pa);
                    ptx::tc_fence_after();
                    tr.ev(12, cb);
                    long long wb = 0;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (wait_b) {
                            const long long t0 = tr.now();
                            ptx::mbar_wait(&bfull[sb], pb);
                            wb += tr.now() - t0;
                            ptx::tc_fence_after();
                        }
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t a_tap =
                            a_lo + (S2 ? (((kh * kS2HaloW + (kw != 0 ? 1 : 0)) * 128 + (kw != 1 ? 64 : 0)) >> 4)
                                       : (((kh * kTapRowPitch + kw) * kARowBytes) >> 4));
#pragma unroll
                        for (int sub = 0; sub < NT; ++sub) {
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) {
                                ptx::umma_bf16_lohi(d_tmem + sub * BN, a_tap + sub * (kHaloABytes >> 4) + 2 * kk, a_hi,
                                                        b_lo + 2 * kk, b_hi, idesc, (cb | tap | kk) != 0);
                            }
                        }
                        if (!p.resident) {
                            ptx::umma_commit(&bempty[sb]);
                        }
                        b_lo += L::kBBytes >> 4;
                        if (++sb == kBStages) { sb = 0; pb ^= 1; b_lo = b_lo0; }
                    }
                    ptx::umma_commit(&aempty[sa]);
                    tr.ev(13, uint32_t(wb >> 4));
                    a_lo += kAStageBytes >> 4;
                    if (++sa == kAStages) { sa = 0; pa ^= 1; a_lo = a_lo0; }
                }
                ptx::umma_commit(&tfull_bar[ab]);
                tr.ev(14, it);
            }
        }
        __syncwarp();
    } else {
        if constexpr (kCatCapable) {
            if (p.chain == 2)
                conv_epilogue_cat(p, sStage, sY, sBias, tfull_bar, tempty_bar, chain_bars, y_full, tmem_base, warp, lane);
        }
        if (p.chain != 2)
            conv_epilogue<BN, NT>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane, nullptr,
                                      nullptr, nullptr, chain_bars);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, p.chain == 2 ? 256u : (p.chain ? 4u * BN : kTmemCols));
    }
}

// ------------------------------------------------------------------------------------------------
// First layer (u8 grey -> 32 channels, 3x3 stride 2) on the tensor cores.  K = 9 taps is one UMMA K step once padded
// to 16, so the whole main loop of a 128-pixel tile is TWO instructions: the weights are split into bf16 hi + lo parts
// (w = hi + lo to 2^-17) and the K block is [9 taps | 0 x 7 | the same 9 taps | 0 x 7] against [hi | 0 | lo | 0] —
// the u8 pixels are exact in bf16, the products exact in the fp32 accumulator.  What the CUDA-core form spends its
// time on (288 FMAs per 32 outputs) disappears; what remains is bias + SiLU + bf16 + store.  A 128 x 32 tile is far
// too small a unit of synchronisation (its barriers and index arithmetic cost more than its work), so the unit here
// is a SUPER-TILE of 64 x 8 output pixels = four UMMA tiles in four 32-column slices of one accumulator buffer:
//   warp 0      : TMA — the weight matrix once, then per super-tile the raw 17 x 160-byte input patch (out-of-image
//                 bytes are zero-filled: the conv padding; the inner coordinate is kept 16-byte aligned) through a ring
//   warp 1      : MMA issuer (eight UMMAs per super-tile)
//   warps 3..18 : two epilogue groups, one per accumulator buffer: four tcgen05.ld + bias / SiLU / bf16 passes into one
//                 32 KB staging tile, ONE barrier pair and ONE TMA store (32 ch x 64 x 8 px) per super-tile
//   warps 19..26: im2col builders: warps 0-3 build sub-tiles 0 and 2, warps 4-7 sub-tiles 1 and 3; thread = output
//                 pixel, reads its 9 bytes from the raw patch, converts (PRMT + FADD, no XU pipe) and writes its
//                 64-byte im2col row with the 64-byte swizzle.
constexpr int kC0Threads = kThreads + 256;
constexpr int kC0RawW = 160, kC0RawH = 17, kC0RawBytes = 2816, kC0RawX = 16;   // 17 x 160 B = 2720 B per slot (128-byte aligned)
constexpr int kC0RawStages = 8, kC0AStages = 8;
constexpr int kC0ABytes = kTileM * 64;                           // im2col tile: 128 rows x 32 bf16
constexpr int kC0TileW = 64, kC0TileH = 8, kC0Sub = 4;           // super-tile, UMMA tiles per super-tile

__global__ void __launch_bounds__(kC0Threads, 1) conv0_tc_kernel(const __grid_constant__ ConvTcParams p) {
    constexpr int BN = 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sA = smem;                                            // [kC0AStages][128][64 B]
    uint8_t* sStage = sA + kC0AStages * kC0ABytes;                 // 2 groups x 2 x 32 KB output staging (double-buffered)
    uint8_t* sB = sStage + kEpiGroups * 4 * kStageBufBytes;        // [32][64 B] weights (hi | lo)
    uint8_t* sRaw = sB + 2048;                                     // [kC0RawStages][2816 B]
    float* sBias = reinterpret_cast<float*>(sRaw + kC0RawStages * kC0RawBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 64);
    uint64_t* raw_full = bars;                                     // [8] TMA -> builders
    uint64_t* raw_empty = raw_full + kC0RawStages;                 // [8] builders -> TMA
    uint64_t* a_full = raw_empty + kC0RawStages;                   // [8] builders -> MMA
    uint64_t* a_empty = a_full + kC0AStages;                       // [8] MMA -> builders
    uint64_t* tfull_bar = a_empty + kC0AStages;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                          // [2]
    uint64_t* w_full = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int first = blockIdx.x, step = gridDim.x;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.tmIn);
        ptx::prefetch_tmap(&p.tmB);
        ptx::prefetch_tmap(&p.tmD);
        for (int s = 0; s < kC0RawStages; ++s) {
            ptx::mbar_init(&raw_full[s], 1);
            ptx::mbar_init(&raw_empty[s], 8);     // the eight builder warps
        }
        for (int s = 0; s < kC0AStages; ++s) {
            ptx::mbar_init(&a_full[s], 4);        // the four builder warps of a sub-tile
            ptx::mbar_init(&a_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 8);
        }
        ptx::mbar_init(w_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 2 * kC0Sub * BN);
        ptx::tmem_relinquish();
    }
    {
        const float bscale = p.act == kActSiluTanh ? 0.5f : 1.0f;
        for (int i = threadIdx.x; i < BN; i += kC0Threads) sBias[i] = bscale * __ldg(p.bias + i);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_launch_dependents();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA: weights once, raw patches per super-tile
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(w_full, 2048);
            ptx::tma_load_2d(sB, &p.tmB, w_full, 0, 0);
        }
        __syncwarp();
        ptx::grid_dependency_wait();      // the input image is written by the crop / letterbox kernel
        int it = 0;
        for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
            const TileCoord tc = decode_tile(p, tile);
            const int s = it % kC0RawStages;
            ptx::mbar_wait(&raw_empty[s], ((it / kC0RawStages) & 1) ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(&raw_full[s], kC0RawW * kC0RawH);
                ptx::tma_load_3d(sRaw + s * kC0RawBytes, &p.tmIn, &raw_full[s], 2 * tc.x0 - kC0RawX, 2 * tc.y0 - 1, tc.n0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: 4 x 2 UMMAs per super-tile
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t b_desc = ptx::make_kmajor_desc<64>(ptx::smem_u32(sB));
            const uint64_t a_desc0 = ptx::make_kmajor_desc<64>(ptx::smem_u32(sA));
            ptx::mbar_wait(w_full, 0);
            int it = 0;
            for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
                const int ab = it & 1;
                ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < kC0Sub; ++j) {
                    const int s = (it * kC0Sub + j) % kC0AStages;
                    ptx::mbar_wait(&a_full[s], ((it * kC0Sub + j) / kC0AStages) & 1);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = uint32_t(a_desc0) + s * (kC0ABytes >> 4);
                    const uint32_t d_tmem = tmem_base + (ab * kC0Sub + j) * BN;
                    ptx::umma_bf16_lohi(d_tmem, a_lo, uint32_t(a_desc0 >> 32), uint32_t(b_desc), uint32_t(b_desc >> 32), idesc,
                                        false);
                    ptx::umma_bf16_lohi(d_tmem, a_lo + 2, uint32_t(a_desc0 >> 32), uint32_t(b_desc) + 2,
                                        uint32_t(b_desc >> 32), idesc, true);
                    ptx::umma_commit(&a_empty[s]);
                }
                ptx::umma_commit(&tfull_bar[ab]);
            }
        }
        __syncwarp();
    } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + 16) {
        // ------------------------------------------------------------------ epilogue: one accumulator buffer per group
        const int ew = warp - kFirstEpiWarp;
        const int g = ew >> 3, h = (ew >> 2) & 1;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool store_thread = (threadIdx.x - kFirstEpiWarp * 32 - g * kEpiThreads) == 0;
        uint8_t* stage0 = sStage + g * 4 * kStageBufBytes;
        const int bar_id = kEpiBarrier + g;
        const int xr = (row >> 1) & 3;
        float bias[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) bias[j] = sBias[h * 16 + j];
        ptx::grid_dependency_wait();
        int it = g;
        for (int tile = first + g * step; tile < p.num_tiles; tile += 2 * step, it += 2) {
            const TileCoord tc = decode_tile(p, tile);
            // two staging tiles per group: the TMA store of the previous super-tile may still be reading the other one
            // (ncu: with one tile a quarter of the epilogue's samples sat at the barrier behind tma_store_wait_read<0>)
            uint8_t* stage = stage0 + ((it >> 1) & 1) * 2 * kStageBufBytes;
            ptx::mbar_wait(&tfull_bar[g], (it >> 1) & 1);
            if (store_thread) ptx::tma_store_wait_read<1>();
            ptx::tc_fence_after();
            ptx::bar_sync(bar_id, kEpiThreads);
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * kC0Sub * BN + h * 16;
#pragma unroll 2
            for (int j = 0; j < kC0Sub; ++j) {
                uint32_t acc[16];
                ptx::tmem_ld_32x16(t_lane + j * BN, acc);
                ptx::tmem_ld_wait();
                if (j == kC0Sub - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[g]);
                }
                float v[16];
                if (p.act == kActSiluTanh) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float hh = fmaf(__uint_as_float(acc[k]), 0.5f, bias[k]);
                        v[k] = fmaf(hh, tanh_fast(hh), hh);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        v[k] = __uint_as_float(acc[k]) + bias[k];
                        if (p.act == WT_ACT_SILU) v[k] = __fdividef(v[k], 1.0f + __expf(-v[k]));
                    }
                }
                uint8_t* rowp = stage + j * kC0ABytes + row * 64;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint4 o;
                    o.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
                    o.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
                    o.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
                    o.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
                    *reinterpret_cast<uint4*>(rowp + (((2 * h + c) ^ xr) << 4)) = o;
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::bar_sync(bar_id, kEpiThreads);
            if (store_thread) {
                ptx::tma_store_4d(&p.tmD, stage, p.dst_coff, tc.x0, tc.y0, tc.n0);
                ptx::tma_store_commit();
            }
        }
        if (store_thread) ptx::tma_store_wait<0>();
    } else if (warp >= kFirstEpiWarp + 16) {
        // ------------------------------------------------------------------ im2col builders
        const int bw = warp - (kFirstEpiWarp + 16);          // 0..7
        const int r = (bw & 3) * 32 + lane;                  // pixel of a UMMA tile == accumulator row
        const int lx = r % kC0TileW;
        const int xr = (r >> 1) & 3;
        const uint32_t zero = 0u;
        int it = 0;
        for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
            const int rs = it % kC0RawStages;
            ptx::mbar_wait(&raw_full[rs], (it / kC0RawStages) & 1);
            uint32_t v[2][9];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = (bw >> 2) + 2 * k;                             // sub-tile of this warp in round k
                const int ly = 2 * j + r / kC0TileW;                         // output row inside the super-tile
                const uint8_t* raw = sRaw + rs * kC0RawBytes + (2 * ly) * kC0RawW + 2 * lx + (kC0RawX - 1);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) v[k][kh * 3 + kw] = raw[kh * kC0RawW + kw];
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&raw_empty[rs]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = (bw >> 2) + 2 * k;
                const int as = (it * kC0Sub + j) % kC0AStages;
                // u8 -> f32 exactly (byte in the mantissa of 2^23, minus 2^23), bf16 = the upper half of that f32
                uint32_t f[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) f[t] = __float_as_uint(__uint_as_float(0x4B000000u | v[k][t]) - 8388608.0f);
                uint4 c0, c1;
                c0.x = __byte_perm(f[0], f[1], 0x7632);
                c0.y = __byte_perm(f[2], f[3], 0x7632);
                c0.z = __byte_perm(f[4], f[5], 0x7632);
                c0.w = __byte_perm(f[6], f[7], 0x7632);
                c1 = make_uint4(__byte_perm(f[8], zero, 0x7632), 0u, 0u, 0u);
                ptx::mbar_wait(&a_empty[as], (((it * kC0Sub + j) / kC0AStages) & 1) ^ 1);
                uint8_t* rowp = sA + as * kC0ABytes + r * 64;
                *reinterpret_cast<uint4*>(rowp + ((0 ^ xr) << 4)) = c0;
                *reinterpret_cast<uint4*>(rowp + ((1 ^ xr) << 4)) = c1;
                *reinterpret_cast<uint4*>(rowp + ((2 ^ xr) << 4)) = c0;
                *reinterpret_cast<uint4*>(rowp + ((3 ^ xr) << 4)) = c1;
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&a_full[as]);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * kC0Sub * BN);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host

struct ConvTcPlan {
    ConvTcParams prm;
    int smem_bytes;
    bool halo;
    int bn, bk;
    int s2;                    // halo kernel in its stride-2 pixel-pair form
    int nt = 1;                // halo kernel: pixel tiles per weight pass (ConvTcParams.nt)
    int il = 1;                // halo kernel: images interleaved per tile (ConvTcParams.il)
    bool conv0 = false;        // the first layer on the tensor cores (conv0_tc_kernel)
    int pix_per_image_tiles;   // tiles_x * tiles_y (work items per image and N block)
};

static void choose_patch(int w, int h, int batch, int* tw, int* th, int* tn) {
    // (tw, th, tn) powers of two with product 128 minimising padded work; prefer wide rows
    double best = 1e30;
    for (int a = 128; a >= 1; a >>= 1) {
        for (int b = 128 / a; b >= 1; b >>= 1) {
            const int c = 128 / (a * b);
            const double tiles = double(ceil_div(w, a)) * ceil_div(h, b) * ceil_div(batch, c);
            const double cost = tiles * 128.0 / (double(w) * h * batch) + (c > 1 ? 1e-3 * c : 0.0) - 1e-5 * a;
            if (cost < best) {
                best = cost;
                *tw = a;
                *th = b;
                *tn = c;
            }
        }
    }
}

// N tile: the widest one dividing cout, unless a narrower one needs fewer tile waves on this GPU (a 20x20
// map at batch 64 is 200 pixel tiles: one 256-wide wave and a 52-tile tail cost more than three 128-wide waves).
static int pick_bn(int cout, long long m_tiles, int sm_count, bool allow_192) {
    int best = 0;
    long long best_cost = 0;
    static const int cand[5] = {256, 192, 128, 64, 32};
    for (int ci = 0; ci < 5; ++ci) {
        const int bn = cand[ci];
        if (cout % bn != 0 || (bn == 192 && !allow_192)) continue;
        const long long tiles = m_tiles * (cout / bn);
        const long long waves = (tiles + sm_count - 1) / sm_count;
        static const int fixed_env = knob("WT_BN_FIXED", 32);   // A/B knob
        const long long cost = waves * (bn + fixed_env);   // per-tile time ~ N plus a fixed part
        if (best == 0 || cost < best_cost) {
            best = bn;
            best_cost = cost;
        }
    }
    return best;
}

int conv_tc_plan_create(const ConvDesc& d, ConvTcPlan** out) {
    WT_REQUIRE(d.k == 1 || d.k == 3, "conv kernel size must be 1 or 3");
    WT_REQUIRE(d.stride == 1 || (d.stride == 2 && d.k == 3), "stride 2 needs a 3x3 kernel");
    WT_REQUIRE(d.src.dtype == WT_DT_BF16, "conv input must be bf16");
    WT_REQUIRE(d.dst.dtype == WT_DT_BF16 || d.dst.dtype == WT_DT_F32, "conv output must be bf16 or f32");
    const int ho = d.dst.h, wo = d.dst.w;
    int sm_count = 148;
    wt_device_info(&sm_count, nullptr, nullptr);
    static const int s2_env = knob("WT_CONV_S2HALO", 1);
    const bool tall_enough = ceil_div(ho, 16) * 16 * 4 <= ho * 5;
    // stride-2 pair form: 32 input channels that fill their buffer (a pixel pair is one contiguous 128-byte row)
    const bool halo_s2 = s2_env && d.k == 3 && d.stride == 2 && d.cin == 32 && d.src.ctot == 32 && d.src.coff == 0 &&
                         wo % 8 == 0 && tall_enough && d.cout % 32 == 0 && d.cout <= 64 && !d.res.base;
    const bool halo_shape = halo_s2 || (d.k == 3 && d.stride == 1 && (d.cin % 64 == 0 || d.cin == 32) && wo % 8 == 0 &&
                                        tall_enough);
    long long m_tiles;
    if (halo_shape) {
        m_tiles = (long long)ceil_div(wo, 8) * ceil_div(ho, 16) * d.batch;
    } else {
        int tw, th, tn;
        choose_patch(wo, ho, d.batch, &tw, &th, &tn);
        m_tiles = (long long)ceil_div(wo, tw) * ceil_div(ho, th) * ceil_div(d.batch, tn);
    }
    int bn = d.dot_w ? (d.cout <= 256 && d.cout % 32 == 0 ? d.cout : 0)
                     : pick_bn(d.cout, m_tiles, sm_count, d.cin % 64 == 0 && !halo_s2 && !d.add.base);
    if (d.add.base && bn > 128) bn = 128;   // the addend patch buffers (N x 128 B per epilogue group) stay small
    const bool cat_chain = d.chain_w && d.cat.base;
    if (cat_chain) {
        // concat chain (C2f exit): 3x3 / stride 1, 32 -> 32 channels with the residual y1 = upper half of the 64-channel
        // cat slice, then a 1x1 conv over concat(cat slice, output) -> 64 channels
        WT_REQUIRE(halo_shape && !halo_s2 && d.cin == 32 && d.cout == 32 && d.cat_c == kCatC && d.chain_cout == kCatOut,
                   "a concat chain is a 32 -> 32 3x3 conv followed by a (64 + 32) -> 64 1x1 conv");
        WT_REQUIRE(d.res.base == d.cat.base && d.res.coff == d.cat.coff + 32 && d.res.ctot == d.cat.ctot &&
                       d.cat.h == ho && d.cat.w == wo && d.cat.dtype == WT_DT_BF16,
                   "the residual of a concat chain is the upper half of its cat slice");
        WT_REQUIRE(!d.add.base && !d.dot_w && d.dst.dtype == WT_DT_BF16 && (d.cat.ctot * 2) % 16 == 0 &&
                       (d.cat.coff * 2) % 16 == 0,
                   "a concat chain has no addend / dot head and writes bf16");
        bn = 32;
    } else if (d.chain_w) {
        WT_REQUIRE(d.chain_cout == 0 || d.chain_cout == d.cout, "a chained 1x1 conv keeps the channel count");
        WT_REQUIRE(d.cout == 64 || d.cout == 128, "a chained 1x1 conv needs 64 or 128 channels (one N tile, 4 accumulators)");
        WT_REQUIRE(!d.res.base && !d.add.base && !d.dot_w && d.dst.dtype == WT_DT_BF16,
                   "a chained conv has no residual / addend / dot head and writes bf16");
        bn = d.cout;
    }
    WT_REQUIRE(bn != 0, "cout must be a multiple of 32");
    WT_REQUIRE(d.cout <= kMaxCout, "cout exceeds the shared-memory bias vector");
    int bk = (d.cin % 64 == 0) ? 64 : 32;
    WT_REQUIRE(d.cin % bk == 0, "cin must be a multiple of 32");
    if (d.stride == 1) {
        WT_REQUIRE(d.src.h == ho && d.src.w == wo, "stride-1 conv keeps the spatial size");
    } else {
        WT_REQUIRE(d.src.h == 2 * ho && d.src.w == 2 * wo, "stride-2 conv halves an even spatial size");
    }
    const bool out_f32 = d.dst.dtype == WT_DT_F32;
    WT_REQUIRE(!(out_f32 && d.res.base), "residual only with bf16 output");
    WT_REQUIRE(!(d.add.base && d.dot_w), "addend and dot head are not combined");
    WT_REQUIRE(!d.add.base || (d.k == 1 && d.stride == 1), "the upsampled addend is implemented for 1x1 convs");
    if (d.dot_w) {
        WT_REQUIRE(bn == d.cout && 2 * d.cout + 1 <= kMaxCout, "a dot-head conv needs all channels in one N tile");
        WT_REQUIRE(out_f32 && d.dst.ctot == 1 && !d.res.base, "a dot-head conv writes a 1-channel f32 buffer");
    }
    // TMA needs 16-byte aligned global strides and base addresses
    WT_REQUIRE((d.src.ctot * 2) % 16 == 0 && (d.src.coff * 2) % 16 == 0, "source channel alignment");
    WT_REQUIRE(d.dot_w || (d.dst.ctot * (out_f32 ? 4 : 2)) % 16 == 0, "destination channel alignment");

    ConvTcPlan* pl = new ConvTcPlan();
    ConvTcParams& p = pl->prm;
    pl->bn = bn;
    pl->bk = bk;
    // (CTA pairs — cta_group::2, M = 256 — were implemented in round 1, measured 15-30 % slower than single-CTA MMAs on
    // these layer shapes and removed in round 2: git history has them.)
    pl->s2 = 0;
    // 3x3 / stride-1 layers use the halo-reuse kernel (tile = 16 rows x 8 columns of one image) unless the
    // map height wastes more than a quarter of the 16-row tiles (20x20 maps stay on the generic kernel)
    static const int halo_env = knob("WT_CONV_HALO", 1);
    pl->halo = halo_env != 0 && halo_shape;
    if (pl->halo) {
        bk = d.cin % 64 == 0 ? 64 : 32;
        pl->bk = bk;
        pl->s2 = halo_s2 ? 1 : 0;
        // 8 x 16 tiles of one image, or — where 16-row tiles would hang over the map (40 x 40) — 8 x 8 tiles of two
        // images interleaved row by row (WT_CONV_IL=0 switches the second form off for A/B runs)
        static const int il_env = knob("WT_CONV_IL", 1);
        pl->il = (il_env && !halo_s2 && !d.chain_w && bk == 64 && (bn == 64 || bn == 128 || bn == 192) &&
                  ho % 16 != 0 && ho % 8 == 0) ? 2 : 1;
        p.tw = 8;
        p.th = 16 / pl->il;
        p.tn = pl->il;
    } else {
        choose_patch(wo, ho, d.batch, &p.tw, &p.th, &p.tn);
    }
    p.tiles_x = ceil_div(wo, p.tw);
    p.tiles_y = ceil_div(ho, p.th);
    p.tiles_n = ceil_div(d.batch, p.tn);
    p.n_blocks = d.cout / bn;
    p.ksize = d.k;
    p.stride = d.stride;
    p.cin = d.cin;
    p.cout = d.cout;
    p.cin_blocks = ceil_div(d.cin, bk);
    p.src_coff = d.src.coff;
    p.dst_coff = d.dst.coff;
    p.res_coff = d.res.base ? d.res.coff : 0;
    static const int silu_exact = knob("WT_SILU_EXACT", 0);
    p.act = (d.act == WT_ACT_SILU && !silu_exact) ? kActSiluTanh : d.act;
    p.has_res = (d.res.base && !cat_chain) ? 1 : 0;   // concat chain: the residual is read from the cat tile in smem
    p.out_f32 = out_f32 ? 1 : 0;
    p.bias = d.bias;
    p.has_add = d.add.base ? 1 : 0;
    p.add_coff = d.add.coff;
    p.dot_w = d.dot_w;
    p.chain = cat_chain ? 2 : (d.chain_w ? 1 : 0);
    p.cat_coff = cat_chain ? d.cat.coff : 0;
    p.bias2 = d.chain_bias;
    p.act2 = (d.chain_act == WT_ACT_SILU && !silu_exact) ? kActSiluTanh : d.chain_act;
    p.dot_out = d.dot_w ? static_cast<float*>(d.dst.base) : nullptr;
    p.out_w = wo;
    p.out_h = ho;
    p.n_images = 0;
    p.num_tiles = 0;
    pl->pix_per_image_tiles = p.tiles_x * p.tiles_y;
    // shared-memory plan
    // One staging buffer per epilogue group: the bytes buy more pipeline stages, which these latency-bound loops need more
    // than a second staging buffer (round 2, same-GPU A/B at steady state: +0.5 % on the whole forward; the 1x1 layers
    // at 40 x 40 / 20 x 20 gain 5-10 %).  WT_EPI_BUFS=2 (tuning builds) restores "two for 1x1 and narrow layers".
    p.epi_bufs = 1;
    if (knob("WT_EPI_BUFS", 1) == 2) p.epi_bufs = (d.k == 1 || bn <= 64) ? 2 : 1;
    if (d.chain_w) p.epi_bufs = bn / 64;   // the staging buffers of a group hold the whole bf16 tile (A of the second GEMM)
    if (cat_chain) p.epi_bufs = 2;         // b tile + output tile
    const int fixed = fixed_smem_bytes(p.epi_bufs) + (d.add.base ? add_smem_bytes(bn) : 0) +
                      (cat_chain ? kCatSmemBytes : (d.chain_w ? bn * bn * 2 : 0));
    pl->nt = 1;
    if (pl->halo) {
        // Shared-memory plan of the halo kernel for `nt` pixel tiles per weight pass (one A stage = the halo tiles of a
        // whole group).  Returns false when fewer than two A stages (no load / MMA overlap) or two weight stages fit.
        static const int nine_env = knob("WT_CONV_NINE", 1);
        static const int resident_env = knob("WT_CONV_RESIDENT", 1);
        const int b_bytes = bn * bk * 2;
        auto plan = [&](int nt, int epi_bufs) -> bool {
            const int fx = fixed_smem_bytes(epi_bufs) + (cat_chain ? kCatSmemBytes : (d.chain_w ? bn * bn * 2 : 0));
            const int a_bytes = nt * halo_a_bytes(bk, pl->s2, pl->il);
            int a_stages = nt > 1 ? 2 : (bn == 256 ? 2 : 3);
            int stages = (kSmemBudget - fx - a_stages * a_bytes) / b_bytes;
            if (stages > 12) stages = 12;
            // exactly nine weight slots (slot == tap: the fast issue loop) if they fit beside two halo stages
            if (nine_env && stages != 9 && 9 * b_bytes + 2 * a_bytes + fx <= kSmemBudget) {
                stages = 9;
                a_stages = (kSmemBudget - fx - 9 * b_bytes) / a_bytes;
                if (a_stages > kMaxAStages) a_stages = kMaxAStages;
            }
            const bool resident = resident_env && p.cin_blocks == 1 && p.n_blocks == 1 && stages >= 9;
            if (resident) {
                stages = 9;   // the ring wraps once per tile: stage index == tap
                a_stages = (kSmemBudget - fx - 9 * b_bytes) / a_bytes;
                if (a_stages > kMaxAStages) a_stages = kMaxAStages;
            }
            if (stages < 2 || a_stages < (nt > 1 ? 2 : 1)) return false;
            p.stages = stages;
            p.a_stages = a_stages;
            p.resident = resident ? 1 : 0;
            p.epi_bufs = epi_bufs;
            pl->nt = nt;
            pl->smem_bytes = a_stages * a_bytes + stages * b_bytes + fx;
            return true;
        };
        // tile groups: as many pixel tiles per weight pass as TMEM (2 x nt x bn columns <= 512) and shared memory allow;
        // the chained forms keep one tile per pass (their second GEMM uses the other accumulators).  WT_CONV_NT caps it.
        static const int nt_env = knob("WT_CONV_NT", 4);
        int nt_max = !d.chain_w ? nt_env : 1;
        if (nt_max != 1 && nt_max != 2 && nt_max != 4) nt_max = 1;
        // Among the feasible group sizes take the one with the fewest tile-times on this GPU: groups are dealt to the
        // SMs in waves, so a bigger group can cost a partly empty last wave (960 tiles on 148 SMs: 7 waves of single
        // tiles, but 4 waves of pairs = 8 tile-times); ties go to the bigger group (less weight traffic, fewer barriers).
        bool ok = false;
        double best_cost = 0.0;
        int best_nt = 0, best_epi = 0;
        const int epi0 = p.epi_bufs;
        for (int nt = nt_max; nt >= 1; nt >>= 1) {
            if (nt > 1 && (2 * nt * bn > 512 || (nt == 4 && (bn > 64 || bk != 32)))) continue;
            int epi = epi0;
            bool fits = plan(nt, epi);
            if (!fits && nt > 1 && epi0 == 2 && !d.chain_w) {   // (a second staging buffer matters less than a second A stage)
                epi = 1;
                fits = plan(nt, epi);
            }
            if (!fits) continue;
            const long long groups = (long long)p.tiles_x * p.tiles_y * ceil_div(d.batch, p.tn * nt) * p.n_blocks;
            const double cost = double((groups + sm_count - 1) / sm_count) * nt * (nt == 4 ? 0.94 : (nt == 2 ? 0.97 : 1.0));
            if (!ok || cost < best_cost) {
                ok = true;
                best_cost = cost;
                best_nt = nt;
                best_epi = epi;
            }
        }
        if (ok) plan(best_nt, best_epi);
        if (!ok) {
            delete pl;
            set_error("not enough shared memory for the halo weight pipeline");
            return 1;
        }
        p.nt = pl->nt;
        p.il = pl->il;
    } else {
        p.nt = 1;
        p.il = 1;
        p.a_stages = 0;
        p.resident = 0;
        const int stage_bytes = (kTileM + bn) * bk * 2;
        p.stages = (kSmemBudget - fixed) / stage_bytes;
        if (p.stages > kMaxStages) p.stages = kMaxStages;
        pl->smem_bytes = p.stages * stage_bytes + fixed;
    }

    {
        // split epilogue: both groups drain every work item when it has at least two staging units (units of 64 channels
        // for bf16 output with BN >= 64, of 32 channels for f32 output; a bf16 BN = 32 tile is one unit)
        static const int split_env = knob("WT_EPI_SPLIT", 1);
        const int unit_ch = (bn >= 64 && !out_f32) ? 64 : 32;
        const int units = p.nt * (bn / unit_ch);
        // (measured per layer: an odd unit count — BN = 192 — and the addend layers lose with the split, the rest gain)
        p.split = (split_env && p.chain == 0 && !d.dot_w && units >= 2 && (units % 2 == 0 || split_env == 2) &&
                   (!d.add.base || split_env == 2)) ? 1 : 0;
    }
    const int sw_in = bk * 2;   // swizzle span == K-block row bytes
    int rc = 0;
    const uint32_t box_a[4] = {uint32_t(bk), uint32_t(pl->halo ? kHaloW : p.tw), uint32_t(pl->halo ? kHaloH : p.th),
                               uint32_t(p.tn)};
    if (pl->halo && pl->s2) {
        // pixel-pair view of the whole input: [64 = 2 px x 32 ch][w / 2 pairs][h][n]
        const uint64_t dims[4] = {64, uint64_t(d.src.w / 2), uint64_t(d.src.h), uint64_t(d.batch)};
        const uint64_t str[3] = {128, uint64_t(d.src.w) * 64, uint64_t(d.src.w) * 64 * d.src.h};
        const uint32_t box[4] = {64, kS2HaloW, kS2HaloH, 1};
        rc |= encode_tmap(&p.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.src.base, dims, str, box, 128);
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    } else if (pl->halo && pl->il == 2) {
        // (channel, x, image, row): a box is the (8 + 2) x (8 + 2) halo of TWO images, rows of the two images alternating
        const uint64_t img = uint64_t(d.src.ctot) * 2 * d.src.w * d.src.h, row = uint64_t(d.src.ctot) * 2 * d.src.w;
        const uint64_t dims[4] = {uint64_t(d.src.coff + d.cin), uint64_t(d.src.w), uint64_t(d.batch), uint64_t(d.src.h)};
        const uint64_t str[3] = {uint64_t(d.src.ctot) * 2, img, row};
        const uint32_t box[4] = {uint32_t(bk), kHaloW, 2, 10};
        rc |= encode_tmap(&p.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.src.base, dims, str, box, sw_in);
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    } else if (d.stride == 1) {
        // the channel extent ends with the slice, so a K block wider than the slice is zero-filled
        const uint64_t dims[4] = {uint64_t(d.src.coff + d.cin), uint64_t(d.src.w), uint64_t(d.src.h), uint64_t(d.batch)};
        const uint64_t str[3] = {uint64_t(d.src.ctot) * 2, uint64_t(d.src.ctot) * 2 * d.src.w,
                                 uint64_t(d.src.ctot) * 2 * d.src.w * d.src.h};
        rc |= encode_tmap(&p.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.src.base, dims, str, box_a, sw_in);
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    } else {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                uint8_t* base = static_cast<uint8_t*>(d.src.base) + (size_t(py) * d.src.w + px) * d.src.ctot * 2;
                const uint64_t dims[4] = {uint64_t(d.src.ctot), uint64_t(d.src.w / 2), uint64_t(d.src.h / 2),
                                          uint64_t(d.batch)};
                const uint64_t str[3] = {uint64_t(d.src.ctot) * 2 * 2, uint64_t(d.src.ctot) * 2 * d.src.w * 2,
                                         uint64_t(d.src.ctot) * 2 * d.src.w * d.src.h};
                rc |= encode_tmap(&p.tmA[py * 2 + px], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box_a,
                                  sw_in);
            }
    }
    if (pl->halo) {
        const uint64_t dims[3] = {uint64_t(d.cin), 9, uint64_t(d.cout)};
        const uint64_t str[2] = {uint64_t(d.cin) * 2, uint64_t(d.cin) * 2 * 9};
        const uint32_t box[3] = {uint32_t(bk), 1, uint32_t(bn)};
        rc |= encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), dims, str, box,
                          sw_in);
    } else {
        const uint64_t ktot = uint64_t(d.k) * d.k * d.cin;
        const uint64_t dims[2] = {ktot, uint64_t(d.cout)};
        const uint64_t str[1] = {ktot * 2};
        const uint32_t box[2] = {uint32_t(bk), uint32_t(bn)};
        rc |= encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.w), dims, str, box,
                          sw_in);
    }
    p.tmB2 = p.tmB;
    p.tmB2b = p.tmB;
    p.tmY = p.tmA[0];
    if (cat_chain) {
        // W2 [64][96] bf16: K block 0 = 64 channels (128-byte rows), K block 1 = 32 channels (64-byte rows)
        const uint64_t dims[2] = {uint64_t(kCatC + 32), uint64_t(kCatOut)};
        const uint64_t str[1] = {uint64_t(kCatC + 32) * 2};
        const uint32_t box_a[2] = {uint32_t(kCatC), uint32_t(kCatOut)}, box_b[2] = {32, uint32_t(kCatOut)};
        rc |= encode_tmap(&p.tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.chain_w), dims, str,
                          box_a, 128);
        rc |= encode_tmap(&p.tmB2b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.chain_w), dims, str,
                          box_b, 64);
        // the cat tile of a 16 x 8 pixel patch: 64 channels = one 128-byte row per pixel
        const uint64_t ydims[4] = {uint64_t(d.cat.coff + kCatC), uint64_t(wo), uint64_t(ho), uint64_t(d.batch)};
        const uint64_t ystr[3] = {uint64_t(d.cat.ctot) * 2, uint64_t(d.cat.ctot) * 2 * wo, uint64_t(d.cat.ctot) * 2 * wo * ho};
        const uint32_t ybox[4] = {uint32_t(kCatC), 8, 16, 1};
        rc |= encode_tmap(&p.tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.cat.base, ydims, ystr, ybox, 128);
    } else if (d.chain_w) {   // W2 [cout][cout] bf16, K-major rows; one box = 64 input channels of every output channel
        const uint64_t dims[2] = {uint64_t(d.cout), uint64_t(d.cout)};
        const uint64_t str[1] = {uint64_t(d.cout) * 2};
        const uint32_t box[2] = {64, uint32_t(d.cout)};
        rc |= encode_tmap(&p.tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.chain_w), dims, str,
                          box, 128);
    }
    p.tmP = p.tmA[0];
    if (d.add.base) {
        if (pl->halo || p.tw < 2 || p.th < 2 || d.add.h * 2 != ho || d.add.w * 2 != wo) {
            set_error("the upsampled addend needs the generic kernel and an even pixel patch");
            rc = 1;
        } else {
            // f32 [n][h/2][w/2][ctot]; one box = 32 channels (128 B) of the (tw/2 x th/2 x tn) low-resolution patch
            const uint64_t dims[4] = {uint64_t(d.add.ctot), uint64_t(d.add.w), uint64_t(d.add.h), uint64_t(d.batch)};
            const uint64_t str[3] = {uint64_t(d.add.ctot) * 4, uint64_t(d.add.ctot) * 4 * d.add.w,
                                     uint64_t(d.add.ctot) * 4 * d.add.w * d.add.h};
            const uint32_t box[4] = {32, uint32_t(p.tw / 2), uint32_t(p.th / 2), uint32_t(p.tn)};
            rc |= encode_tmap(&p.tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.add.base, dims, str, box, 128);
        }
    }
    if (d.dot_w) {
        p.tmD = p.tmA[0];   // never used: the dot head stores with plain st.global
        p.tmR = p.tmA[0];
    } else {
        const int es = out_f32 ? 4 : 2;
        const int unit_ch = out_f32 ? 32 : ((bn == 32 && !cat_chain) ? 32 : 64);   // (concat chain: 64 output channels)
        const int sw = unit_ch * es;   // 128 or 64
        const bool il2 = pl->halo && pl->il == 2;   // (channel, x, image, row) instead of (channel, x, row, image)
        const uint64_t drow = uint64_t(d.dst.ctot) * es * wo, dimg = drow * ho;
        const uint64_t dims[4] = {uint64_t(d.dst.ctot), uint64_t(wo), uint64_t(il2 ? d.batch : ho), uint64_t(il2 ? ho : d.batch)};
        const uint64_t str[3] = {uint64_t(d.dst.ctot) * es, il2 ? dimg : drow, il2 ? drow : dimg};
        const uint32_t box[4] = {uint32_t(unit_ch), uint32_t(p.tw), uint32_t(il2 ? p.tn : p.th), uint32_t(il2 ? p.th : p.tn)};
        rc |= encode_tmap(&p.tmD, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                          d.dst.base, dims, str, box, sw);
        if (d.res.base) {
            if (d.res.h != ho || d.res.w != wo || d.res.dtype != WT_DT_BF16) {
                set_error("residual must be a bf16 buffer of the output's spatial size");
                rc = 1;
            } else {
                const uint64_t rrow = uint64_t(d.res.ctot) * 2 * wo, rimg = rrow * ho;
                const uint64_t rdims[4] = {uint64_t(d.res.ctot), uint64_t(wo), uint64_t(il2 ? d.batch : ho),
                                           uint64_t(il2 ? ho : d.batch)};
                const uint64_t rstr[3] = {uint64_t(d.res.ctot) * 2, il2 ? rimg : rrow, il2 ? rrow : rimg};
                rc |= encode_tmap(&p.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.res.base, rdims, rstr, box, sw);
            }
        } else {
            p.tmR = p.tmD;
        }
    }
    if (rc) {
        delete pl;
        return 1;
    }
    *out = pl;
    return 0;
}

void conv_tc_plan_destroy(ConvTcPlan* p) { delete p; }

// First layer on the tensor cores: `wmat` = bf16 [cout = 32][32] rows [hi taps 0..8 | 0 x 7 | lo taps 0..8 | 0 x 7].
int conv0_tc_plan_create(const uint8_t* src, int h, int w, const __nv_bfloat16* wmat, const float* bias, int act,
                         const TensorView& dst, int batch, ConvTcPlan** out) {
    WT_REQUIRE(h % 2 == 0 && w % 16 == 0, "conv0 (tcgen05) needs an even height and a width that is a multiple of 16");
    WT_REQUIRE(dst.dtype == WT_DT_BF16 && dst.h == h / 2 && dst.w == w / 2 && (dst.ctot * 2) % 16 == 0 &&
                   (dst.coff * 2) % 16 == 0,
               "conv0 destination shape");
    ConvTcPlan* pl = new ConvTcPlan();
    ConvTcParams& p = pl->prm;
    memset(&p, 0, sizeof(p));
    pl->bn = 32; pl->bk = 32; pl->s2 = 0; pl->halo = false;
    pl->conv0 = true;
    const int ho = h / 2, wo = w / 2;
    p.tw = kC0TileW; p.th = kC0TileH; p.tn = 1;
    p.tiles_x = ceil_div(wo, p.tw); p.tiles_y = ceil_div(ho, p.th); p.tiles_n = batch;
    p.n_blocks = 1; p.ksize = 3; p.stride = 2; p.cin = 1; p.cin_blocks = 1; p.cout = 32;
    p.dst_coff = dst.coff;
    static const int silu_exact = knob("WT_SILU_EXACT", 0);
    p.act = (act == WT_ACT_SILU && !silu_exact) ? kActSiluTanh : act;
    p.bias = bias;
    p.out_w = wo; p.out_h = ho;
    p.epi_bufs = 2;
    p.nt = 1;
    p.il = 1;
    pl->pix_per_image_tiles = p.tiles_x * p.tiles_y;
    pl->smem_bytes = kC0AStages * kC0ABytes + kEpiGroups * 4 * kStageBufBytes + 2048 + kC0RawStages * kC0RawBytes + 256 +
                     kBarrierBytes;
    int rc = 0;
    {
        const uint64_t dims[3] = {uint64_t(w), uint64_t(h), uint64_t(batch)};
        const uint64_t str[2] = {uint64_t(w), uint64_t(w) * h};
        const uint32_t box[3] = {kC0RawW, kC0RawH, 1};
        rc |= encode_tmap(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(src), dims, str, box, 0);
    }
    {
        const uint64_t dims[2] = {32, 32};
        const uint64_t str[1] = {64};
        const uint32_t box[2] = {32, 32};
        rc |= encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wmat), dims, str, box, 64);
    }
    {
        const uint64_t dims[4] = {uint64_t(dst.ctot), uint64_t(wo), uint64_t(ho), uint64_t(batch)};
        const uint64_t str[3] = {uint64_t(dst.ctot) * 2, uint64_t(dst.ctot) * 2 * wo, uint64_t(dst.ctot) * 2 * wo * ho};
        const uint32_t box[4] = {32, uint32_t(p.tw), uint32_t(p.th), 1};
        rc |= encode_tmap(&p.tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dst.base, dims, str, box, 64);
    }
    p.tmR = p.tmD; p.tmP = p.tmD; p.tmB2 = p.tmB; p.tmB2b = p.tmB; p.tmY = p.tmD;
    for (int i = 0; i < 4; ++i) p.tmA[i] = p.tmD;
    if (rc) {
        delete pl;
        return 1;
    }
    *out = pl;
    return 0;
}

template <typename Kernel>
static int launch_kernel(Kernel kernel, SmemOptIn* opt_in, const ConvTcParams& prm, int smem, int grid,
                         cudaStream_t stream) {
    WT_CHECK_CUDA(opt_in_smem(kernel, *opt_in, kSmemBudget));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    static const int pdl_env = knob("WT_CONV_PDL", 1);
    if (pdl_env) {   // start this kernel's prologue while the previous kernel of the stream drains
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    WT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, prm));
    WT_LAUNCHED();
    return 0;
}

template <int BN, int BK>
static int launch_inst(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static SmemOptIn opt_in;
    return launch_kernel(conv_tc_kernel<BN, BK>, &opt_in, prm, smem, grid, stream);
}

template <int BN, int BK, int S2 = 0, int NT = 1, int IL = 1>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static SmemOptIn opt_in;
    return launch_kernel(conv_halo_kernel<BN, BK, S2, NT, IL>, &opt_in, prm, smem, grid, stream);
}

#ifdef WT_TUNING_KNOBS
// Event trace of the tcgen05 conv launches (tuning builds): launch i since the call writes region i of `buf`
// ([max_launches][148 CTAs][4 roles][cap] u64, zero-filled by the caller); nullptr switches tracing off.
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0, g_trace_max = 0, g_trace_next = 0;
extern "C" int wt_debug_conv_trace(unsigned long long* buf, int cap, int max_launches) {
    g_trace_buf = buf;
    g_trace_cap = cap;
    g_trace_max = max_launches;
    g_trace_next = 0;
    return 0;
}
#endif

int conv_tc_launch(const ConvTcPlan* pl, int n_images, int sm_count, cudaStream_t stream) {
    ConvTcParams prm = pl->prm;
#ifdef WT_TUNING_KNOBS
    prm.trace = nullptr;
    prm.trace_cap = g_trace_cap;
    if (g_trace_buf && g_trace_next < g_trace_max) {
        prm.trace = g_trace_buf + size_t(g_trace_next) * 148 * 4 * g_trace_cap;
        ++g_trace_next;
    }
#endif
    const int tiles_n = ceil_div(n_images, prm.tn * prm.nt);   // (halo kernel: groups of nt images)
    prm.num_tiles = pl->pix_per_image_tiles * tiles_n * prm.n_blocks;   // work items
    prm.n_images = n_images;
    if (prm.num_tiles == 0) return 0;
    {
        auto magic = [&](int d) -> uint32_t {
            if (d <= 1) return 0u;                                           // (d == 1: the plain path, a no-op division)
            if ((unsigned long long)prm.num_tiles * (unsigned long long)d >= (1ull << 32)) return 0u;
            return uint32_t((1ull << 32) / (unsigned long long)d) + 1u;
        };
        prm.mg_nb = magic(prm.n_blocks);
        prm.mg_tx = magic(prm.tiles_x);
        prm.mg_ty = magic(prm.tiles_y);
    }
    const int grid = prm.num_tiles < sm_count ? prm.num_tiles : sm_count;   // persistent: one CTA per SM
    const int smem = pl->smem_bytes;
    if (pl->conv0) {
        static SmemOptIn opt_in;
        WT_CHECK_CUDA(opt_in_smem(conv0_tc_kernel, opt_in, kSmemBudget));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kC0Threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        WT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv0_tc_kernel, prm));
        WT_LAUNCHED();
        return 0;
    }
    if (pl->halo) {
        const int nt = pl->nt;
        if (pl->s2) {         // 32 -> 32/64 channels, stride 2 (layer 1)
            switch (pl->bn * 10 + nt) {
                case 641: return launch_halo<64, 32, 1>(prm, smem, grid, stream);
                case 321: return launch_halo<32, 32, 1>(prm, smem, grid, stream);
                case 642: return launch_halo<64, 32, 1, 2>(prm, smem, grid, stream);
                case 322: return launch_halo<32, 32, 1, 2>(prm, smem, grid, stream);
            }
            set_error("no stride-2 halo instantiation for this (BN, NT)");
            return 1;
        }
        if (pl->bk == 32) {   // 32 input channels (the 160x160 C2f bottlenecks)
            switch (pl->bn * 10 + nt) {
                case 641: return launch_halo<64, 32>(prm, smem, grid, stream);
                case 321: return launch_halo<32, 32>(prm, smem, grid, stream);
                case 642: return launch_halo<64, 32, 0, 2>(prm, smem, grid, stream);
                case 322: return launch_halo<32, 32, 0, 2>(prm, smem, grid, stream);
                case 644: return launch_halo<64, 32, 0, 4>(prm, smem, grid, stream);
                case 324: return launch_halo<32, 32, 0, 4>(prm, smem, grid, stream);
            }
            set_error("no halo instantiation for this (BN, 32, NT)");
            return 1;
        }
        if (pl->il == 1) switch (pl->bn * 10 + nt) {
            case 2561: return launch_halo<256, 64>(prm, smem, grid, stream);
            case 1921: return launch_halo<192, 64>(prm, smem, grid, stream);
            case 1281: return launch_halo<128, 64>(prm, smem, grid, stream);
            case 641:  return launch_halo<64, 64>(prm, smem, grid, stream);
            case 321:  return launch_halo<32, 64>(prm, smem, grid, stream);
            case 1282: return launch_halo<128, 64, 0, 2>(prm, smem, grid, stream);
            case 642:  return launch_halo<64, 64, 0, 2>(prm, smem, grid, stream);
            case 322:  return launch_halo<32, 64, 0, 2>(prm, smem, grid, stream);
        }
        if (pl->il == 2) {   // two images per tile (40 x 40 maps)
            switch (pl->bn * 10 + nt) {
                case 1921: return launch_halo<192, 64, 0, 1, 2>(prm, smem, grid, stream);
                case 1281: return launch_halo<128, 64, 0, 1, 2>(prm, smem, grid, stream);
                case 641:  return launch_halo<64, 64, 0, 1, 2>(prm, smem, grid, stream);
                case 1282: return launch_halo<128, 64, 0, 2, 2>(prm, smem, grid, stream);
                case 642:  return launch_halo<64, 64, 0, 2, 2>(prm, smem, grid, stream);
            }
        }
        set_error("no halo instantiation for this (BN, NT, IL)");
        return 1;
    }
    switch (pl->bn * 100 + pl->bk) {
        case 25664: return launch_inst<256, 64>(prm, smem, grid, stream);
        case 19264: return launch_inst<192, 64>(prm, smem, grid, stream);
        case 12864: return launch_inst<128, 64>(prm, smem, grid, stream);
        case 6464:  return launch_inst<64, 64>(prm, smem, grid, stream);
        case 3264:  return launch_inst<32, 64>(prm, smem, grid, stream);
        case 25632: return launch_inst<256, 32>(prm, smem, grid, stream);
        case 12832: return launch_inst<128, 32>(prm, smem, grid, stream);
        case 6432:  return launch_inst<64, 32>(prm, smem, grid, stream);
        case 3232:  return launch_inst<32, 32>(prm, smem, grid, stream);
        default:
            set_error("no conv_tc instantiation for this (BN, BK)");
            return 1;
    }
}

}  // namespace wt
