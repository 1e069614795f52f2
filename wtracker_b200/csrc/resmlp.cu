// K9: batched ResMLP position predictor (BatchNorm folded into the linear layers on the host).
//   x = relu(W_in x + b_in);  for each block: x = x + block(x), block = block_len x (linear + relu);
//   y = W_out x + b_out                                  (wtracker/neural/mlp.py:144-188)
// One sample per thread; the whole weight set (5.6 k / 20 k floats) sits in shared memory and is
// read as warp-wide broadcasts; per-thread activation vectors live in shared memory with a
// conflict-free [feature][thread] layout.  fp32 FMA throughout.
#include "../../include/wtracker_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace wt {
namespace {

constexpr int kMlpThreads = 128;
constexpr int kLd = kMlpThreads + 1;   // padded leading dimension of the activation tiles

struct MlpParams {
    wt_resmlp_desc d;
    const float* x;
    float* y;
    long long n;
    int maxw;   // widest layer
};

__device__ __forceinline__ void dense(const float* __restrict__ w, const float* __restrict__ b, const float* in,
                                      float* out, int nin, int nout, bool relu, int t) {
    int o = 0;
    for (; o + 4 <= nout; o += 4) {
        float a0 = b[o], a1 = b[o + 1], a2 = b[o + 2], a3 = b[o + 3];
        const float* w0 = w + size_t(o) * nin;
        for (int i = 0; i < nin; ++i) {
            const float v = in[i * kLd + t];
            a0 = fmaf(v, w0[i], a0);
            a1 = fmaf(v, w0[nin + i], a1);
            a2 = fmaf(v, w0[2 * nin + i], a2);
            a3 = fmaf(v, w0[3 * nin + i], a3);
        }
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
        out[o * kLd + t] = a0; out[(o + 1) * kLd + t] = a1; out[(o + 2) * kLd + t] = a2; out[(o + 3) * kLd + t] = a3;
    }
    for (; o < nout; ++o) {
        float a = b[o];
        const float* w0 = w + size_t(o) * nin;
        for (int i = 0; i < nin; ++i) a = fmaf(in[i * kLd + t], w0[i], a);
        out[o * kLd + t] = relu ? fmaxf(a, 0.f) : a;
    }
}

__global__ void __launch_bounds__(kMlpThreads) resmlp_kernel(const MlpParams p) {
    extern __shared__ float mlp_smem[];
    float* sw = mlp_smem;                              // weights
    float* xs = sw + ((p.d.n_weights + 31) & ~31);     // residual stream  [maxw][kLd]
    float* t0 = xs + p.maxw * kLd;                     // scratch A
    float* t1 = t0 + p.maxw * kLd;                     // scratch B
    const int t = threadIdx.x;
    for (int i = t; i < p.d.n_weights; i += kMlpThreads) sw[i] = __ldg(p.d.weights + i);

    const long long base = (long long)blockIdx.x * kMlpThreads;
    const int valid = int(min((long long)kMlpThreads, p.n - base));
    const int ind = p.d.in_dim;
    for (int i = t; i < kMlpThreads * ind; i += kMlpThreads) {
        const int s = i / ind, f = i - s * ind;
        t0[f * kLd + s] = s < valid ? __ldg(p.x + base * ind + i) : 0.f;
    }
    __syncthreads();

    const float* w = sw;
    const int H = p.d.hidden;
    dense(w, w + H * ind, t0, xs, ind, H, true, t);
    w += H * ind + H;
    for (int blk = 0; blk < p.d.n_blocks; ++blk) {
        const float* in = xs;
        int nin = H;
        float* bufs[2] = {t0, t1};
        for (int l = 0; l < p.d.block_len; ++l) {
            const int nout = p.d.block_dims[l];
            float* out = bufs[l & 1];
            dense(w, w + nout * nin, in, out, nin, nout, true, t);
            w += nout * nin + nout;
            in = out;
            nin = nout;
        }
        for (int f = 0; f < H; ++f) xs[f * kLd + t] += in[f * kLd + t];
    }
    dense(w, w + p.d.out_dim * H, xs, t0, H, p.d.out_dim, false, t);
    __syncthreads();
    const int od = p.d.out_dim;
    for (int i = t; i < valid * od; i += kMlpThreads) {
        const int s = i / od, f = i - s * od;
        p.y[base * od + i] = t0[f * kLd + s];
    }
}

// Large batches, two samples per thread: the pair shares every weight load and every issue slot — the weight of
// neuron o for input i is ONE broadcast scalar of a packed f32x2 FMA (FFMA2) whose two lanes are the two samples.
// Weights are staged transposed, [in][out padded to 8], so eight neurons' weights are two LDS.128 broadcasts; per
// input that is 1 LDS.64 (the activation pair) + 2 LDS.128 + 8 FFMA2 for 16 MACs, where the one-sample kernel above
// needs 5 loads + 4 FMAs for 4.  Each neuron still accumulates bias first, then its inputs in order, so all three
// kernels give bit-identical results.  The last layer of a block adds straight into the residual stream.
constexpr int kPairThreads = 128;
constexpr int kPairLd = kPairThreads + 1;    // leading dimension of the [feature][thread] float2 tiles

__device__ __forceinline__ int pad8(int v) { return (v + 7) & ~7; }

// out[o] (+)= act(bias[o] + sum_i in[i] * w[i][o]) for the thread's two samples; `into_x`: out is the residual stream
// (`in_off` / `out_off` are offsets into the dynamic shared-memory window, not pointers: with pointers picked at run
// time the compiler falls back to generic loads, LD.E instead of LDS)
__device__ __forceinline__ void dense_pair(const float* __restrict__ wt, const float* __restrict__ b, int in_off,
                                           int out_off, int nin, int nout, bool relu, bool into_x, int t) {
    extern __shared__ __align__(16) float mlp_smem[];
    const uint64_t* in = reinterpret_cast<const uint64_t*>(mlp_smem) + in_off;
    uint64_t* out = reinterpret_cast<uint64_t*>(mlp_smem) + out_off;
    const int np = pad8(nout);
    for (int o = 0; o < np; o += 8) {
        uint64_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = ptx::pack_f32x2(b[o + k], b[o + k]);
        const float* w0 = wt + o;
#pragma unroll 4
        for (int i = 0; i < nin; ++i) {
            const uint64_t v = in[i * kPairLd + t];
            const float4 wa = *reinterpret_cast<const float4*>(w0 + i * np);
            const float4 wb = *reinterpret_cast<const float4*>(w0 + i * np + 4);
            a[0] = ptx::ffma2(v, ptx::pack_f32x2(wa.x, wa.x), a[0]);
            a[1] = ptx::ffma2(v, ptx::pack_f32x2(wa.y, wa.y), a[1]);
            a[2] = ptx::ffma2(v, ptx::pack_f32x2(wa.z, wa.z), a[2]);
            a[3] = ptx::ffma2(v, ptx::pack_f32x2(wa.w, wa.w), a[3]);
            a[4] = ptx::ffma2(v, ptx::pack_f32x2(wb.x, wb.x), a[4]);
            a[5] = ptx::ffma2(v, ptx::pack_f32x2(wb.y, wb.y), a[5]);
            a[6] = ptx::ffma2(v, ptx::pack_f32x2(wb.z, wb.z), a[6]);
            a[7] = ptx::ffma2(v, ptx::pack_f32x2(wb.w, wb.w), a[7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (o + k < nout) {
                float lo, hi;
                ptx::unpack_f32x2(a[k], lo, hi);
                if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                if (into_x) {
                    float xl, xh;
                    ptx::unpack_f32x2(out[(o + k) * kPairLd + t], xl, xh);
                    lo = xl + lo;
                    hi = xh + hi;
                }
                out[(o + k) * kPairLd + t] = ptx::pack_f32x2(lo, hi);
            }
        }
    }
}

__global__ void __launch_bounds__(kPairThreads) resmlp_pair_kernel(const MlpParams p, int tw0, int tw1,
                                                                   int n_weights_padded) {
    extern __shared__ __align__(16) float mlp_smem[];
    float* sw = mlp_smem;                                            // per layer: [in][pad8(out)] then bias[pad8(out)]
    const int xs_off = n_weights_padded / 2;                         // residual stream [hidden][kPairLd] pairs
    const int t0_off = xs_off + p.d.hidden * kPairLd;                // scratch A [tw0][kPairLd]: input, even block layers, output
    const int t1_off = t0_off + tw0 * kPairLd;                       // scratch B [tw1][kPairLd]: odd block layers
    uint64_t* t0 = reinterpret_cast<uint64_t*>(mlp_smem) + t0_off;
    (void)tw1;
    const int t = threadIdx.x;
    const int H = p.d.hidden, ind = p.d.in_dim, od = p.d.out_dim;
    const int n_layers = 2 + p.d.n_blocks * p.d.block_len;
    {   // stage: transpose to [in][out], pad the outputs to a multiple of 8 with zero weights / zero bias
        int src = 0, dst = 0, nin = ind;
        for (int l = 0; l < n_layers; ++l) {
            const int nout = l == 0 ? H : (l == n_layers - 1 ? od : p.d.block_dims[(l - 1) % p.d.block_len]);
            const int np = pad8(nout);
#pragma unroll 4
            for (int idx = t; idx < nin * np; idx += kPairThreads) {
                const int i = idx / np, o = idx - i * np;
                sw[dst + idx] = o < nout ? __ldg(p.d.weights + src + o * nin + i) : 0.f;
            }
            for (int o = t; o < np; o += kPairThreads) sw[dst + nin * np + o] = o < nout ? __ldg(p.d.weights + src + nout * nin + o) : 0.f;
            src += nout * nin + nout;
            dst += nin * np + np;
            nin = nout;
        }
    }
    // persistent over chunks of 256 samples: the weights are staged once per CTA, not once per 256 samples
    float* t0f = reinterpret_cast<float*>(t0);
    const long long n_chunks = (p.n + 2 * kPairThreads - 1) / (2 * kPairThreads);
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const long long base = chunk * (2 * kPairThreads);
    const int valid = int(min((long long)(2 * kPairThreads), p.n - base));
    __syncthreads();                                                     // the previous chunk's outputs have left t0
#pragma unroll 4
    for (int i = t; i < 2 * kPairThreads * ind; i += kPairThreads) {     // coalesced read, [feature][sample] in smem
        const int s = i / ind, f = i - s * ind;
        t0f[(f * kPairLd + (s >> 1)) * 2 + (s & 1)] = s < valid ? __ldg(p.x + base * ind + i) : 0.f;
    }
    __syncthreads();

    const float* w = sw;
    dense_pair(w, w + ind * pad8(H), t0_off, xs_off, ind, H, true, false, t);
    w += ind * pad8(H) + pad8(H);
    for (int blk = 0; blk < p.d.n_blocks; ++blk) {
        int in = xs_off;
        int nin = H;
        for (int l = 0; l < p.d.block_len; ++l) {
            const int nout = p.d.block_dims[l];
            const bool last = l == p.d.block_len - 1;
            const int out = last ? xs_off : ((l & 1) ? t1_off : t0_off);   // (block_len >= 2: the last layer never reads xs)
            dense_pair(w, w + nin * pad8(nout), in, out, nin, nout, true, last, t);
            w += nin * pad8(nout) + pad8(nout);
            in = out;
            nin = nout;
        }
    }
    dense_pair(w, w + H * pad8(od), xs_off, t0_off, H, od, false, false, t);
    __syncthreads();
    for (int i = t; i < valid * od; i += kPairThreads) {
        const int s = i / od, f = i - s * od;
        p.y[base * od + i] = t0f[(f * kPairLd + (s >> 1)) * 2 + (s & 1)];
    }
    }
}

// Small batches (one simulator step, one detector batch): ONE WARP per sample, lane o owns output neuron o
// (and o + 32), weights transposed to [in][out] while they are staged so that the lanes read consecutive
// words.  Same FMA order per neuron as the thread-per-sample kernel (bias first, inputs in order), so both
// give bit-identical results; the latency of one sample drops from ~4.7k dependent FMAs to ~18 layers x
// (fan-in) FMAs.
constexpr int kMlpWarpThreads = 256;

__device__ __forceinline__ void dense_warp(const float* __restrict__ wt, const float* __restrict__ b, const float* in,
                                           float* out, int nin, int nout, bool relu, int lane) {
    for (int o = lane; o < nout; o += 32) {
        float a = b[o];
#pragma unroll 8
        for (int i = 0; i < nin; ++i) a = fmaf(in[i], wt[i * nout + o], a);   // (same order; the loads run ahead)
        out[o] = relu ? fmaxf(a, 0.f) : a;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kMlpWarpThreads) resmlp_warp_kernel(const MlpParams p) {
    extern __shared__ float mlp_smem[];
    float* sw = mlp_smem;                                          // every layer as [in][out] then bias[out]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = sw + ((p.d.n_weights + 31) & ~31) + warp * 3 * p.maxw;   // residual stream of this warp's sample
    float* t0 = xs + p.maxw;
    float* t1 = t0 + p.maxw;
    const int H = p.d.hidden, ind = p.d.in_dim;
    {   // stage + transpose, layer by layer
        int off = 0, nin = ind;
        const int n_layers = 2 + p.d.n_blocks * p.d.block_len;
        for (int l = 0; l < n_layers; ++l) {
            const int nout = l == 0 ? H : (l == n_layers - 1 ? p.d.out_dim : p.d.block_dims[(l - 1) % p.d.block_len]);
#pragma unroll 4
            for (int idx = threadIdx.x; idx < nout * nin; idx += kMlpWarpThreads) {
                const int o = idx / nin, i = idx - o * nin;
                sw[off + i * nout + o] = __ldg(p.d.weights + off + idx);
            }
            for (int o = threadIdx.x; o < nout; o += kMlpWarpThreads)
                sw[off + nout * nin + o] = __ldg(p.d.weights + off + nout * nin + o);
            off += nout * nin + nout;
            nin = nout;
        }
    }
    const long long sample = (long long)blockIdx.x * (kMlpWarpThreads / 32) + warp;
    const bool live = sample < p.n;
    if (live)
        for (int f = lane; f < ind; f += 32) t0[f] = __ldg(p.x + sample * ind + f);
    __syncthreads();
    if (!live) return;

    const float* w = sw;
    dense_warp(w, w + H * ind, t0, xs, ind, H, true, lane);
    w += H * ind + H;
    for (int blk = 0; blk < p.d.n_blocks; ++blk) {
        const float* in = xs;
        int nin = H;
        float* bufs[2] = {t0, t1};
        for (int l = 0; l < p.d.block_len; ++l) {
            const int nout = p.d.block_dims[l];
            float* out = bufs[l & 1];
            dense_warp(w, w + nout * nin, in, out, nin, nout, true, lane);
            w += nout * nin + nout;
            in = out;
            nin = nout;
        }
        for (int f = lane; f < H; f += 32) xs[f] += in[f];
        __syncwarp();
    }
    dense_warp(w, w + p.d.out_dim * H, xs, t0, H, p.d.out_dim, false, lane);
    for (int f = lane; f < p.d.out_dim; f += 32) p.y[sample * p.d.out_dim + f] = t0[f];
}

__global__ void mlp_gather_kernel(const double* __restrict__ table, long long rows, const int32_t* __restrict__ frame,
                                  const int32_t* __restrict__ offsets, int k, float* __restrict__ x,
                                  uint8_t* __restrict__ valid, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = frame[i];
    double x0 = 0.0, y0 = 0.0;
    bool ok = true;
    for (int j = 0; j < k; ++j) {
        const long long r = (long long)f + offsets[j];
        double b[4];
        if (r >= 0 && r < rows) {
            const double2 lo = reinterpret_cast<const double2*>(table)[2 * r], hi = reinterpret_cast<const double2*>(table)[2 * r + 1];
            b[0] = lo.x; b[1] = lo.y; b[2] = hi.x; b[3] = hi.y;
        } else {
            b[0] = b[1] = b[2] = b[3] = __longlong_as_double(0x7ff8000000000000LL);
        }
        ok = ok && isfinite(b[0]) && isfinite(b[1]) && isfinite(b[2]) && isfinite(b[3]);
        if (j == 0) { x0 = b[0]; y0 = b[1]; }
        float* o = x + (i * k + j) * 4;
        o[0] = float(__dsub_rn(b[0], x0));
        o[1] = float(__dsub_rn(b[1], y0));
        o[2] = float(b[2]);
        o[3] = float(b[3]);
    }
    valid[i] = ok ? 1 : 0;
}

}  // namespace
}  // namespace wt

extern "C" int wt_mlp_gather(const double* table, int64_t table_rows, const int32_t* frame, const int32_t* offsets,
                             int k, float* x, uint8_t* valid, int64_t n, void* stream) {
    using namespace wt;
    if (n == 0) return 0;
    WT_REQUIRE(table && frame && offsets && x && valid && k >= 1, "null argument");
    mlp_gather_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(table, table_rows, frame,
                                                                                                 offsets, k, x, valid, n);
    WT_LAUNCHED();
    return 0;
}

extern "C" int wt_resmlp_forward(const wt_resmlp_desc* d, const float* x, float* y, int64_t n, void* stream) {
    using namespace wt;
    WT_REQUIRE(d && x && y && d->weights, "null argument");
    WT_REQUIRE(d->block_len >= 1 && d->block_len <= 8 && d->n_blocks >= 0, "block shape");
    WT_REQUIRE(d->block_dims[d->block_len - 1] == d->hidden, "a block must map hidden -> hidden");
    int maxw = d->in_dim > d->hidden ? d->in_dim : d->hidden;
    long long expect = (long long)d->hidden * d->in_dim + d->hidden;
    int nin = d->hidden;
    long long per_block = 0;
    for (int l = 0; l < d->block_len; ++l) {
        per_block += (long long)d->block_dims[l] * nin + d->block_dims[l];
        nin = d->block_dims[l];
        if (nin > maxw) maxw = nin;
    }
    expect += per_block * d->n_blocks + (long long)d->out_dim * d->hidden + d->out_dim;
    if (d->out_dim > maxw) maxw = d->out_dim;
    WT_REQUIRE(expect == d->n_weights, "weight blob size does not match the layer description");
    if (n == 0) return 0;
    MlpParams p;
    p.d = *d;
    p.x = x;
    p.y = y;
    p.n = n;
    p.maxw = maxw;
    if (n <= 2048) {   // latency-bound regime: one warp per sample
        const size_t wsmem = (size_t((d->n_weights + 31) & ~31) + size_t(kMlpWarpThreads / 32) * 3 * maxw) * sizeof(float);
        WT_REQUIRE(wsmem <= 220 * 1024, "ResMLP too large for the shared-memory kernel");
        static SmemOptIn wopt;
        WT_CHECK_CUDA(opt_in_smem(resmlp_warp_kernel, wopt, wsmem));
        const long long wblocks = (n + kMlpWarpThreads / 32 - 1) / (kMlpWarpThreads / 32);
        resmlp_warp_kernel<<<(unsigned)wblocks, kMlpWarpThreads, wsmem, static_cast<cudaStream_t>(stream)>>>(p);
        WT_LAUNCHED();
        return 0;
    }
    if (d->block_len >= 2) {   // two samples per thread (FFMA2)
        // padded weight floats, and the scratch width: the input vector and every block layer but the last
        long long wp = (long long)d->in_dim * ((d->hidden + 7) & ~7) + ((d->hidden + 7) & ~7);
        int tw0 = d->in_dim > d->out_dim ? d->in_dim : d->out_dim, tw1 = 1, nin2 = d->hidden;
        long long blockp = 0;
        for (int l = 0; l < d->block_len; ++l) {
            const int np = (d->block_dims[l] + 7) & ~7;
            blockp += (long long)nin2 * np + np;
            nin2 = d->block_dims[l];
            if (l + 1 < d->block_len) {
                int& tw = (l & 1) ? tw1 : tw0;
                if (nin2 > tw) tw = nin2;
            }
        }
        wp += blockp * d->n_blocks + (long long)d->hidden * ((d->out_dim + 7) & ~7) + ((d->out_dim + 7) & ~7);
        wp = (wp + 3) & ~3LL;
        const size_t psmem = size_t(wp) * 4 + size_t(d->hidden + tw0 + tw1) * kPairLd * 8;
        if (psmem <= 220 * 1024) {
            static SmemOptIn popt;
            WT_CHECK_CUDA(opt_in_smem(resmlp_pair_kernel, popt, psmem));
            long long pblocks = (n + 2 * kPairThreads - 1) / (2 * kPairThreads);
            int sm_count = 148;
            wt_device_info(&sm_count, nullptr, nullptr);
            const long long resident = (long long)sm_count * (psmem > 110 * 1024 ? 1 : 2);
            if (pblocks > resident) pblocks = resident;          // persistent: a CTA walks chunks blockIdx.x, + grid, ...
            resmlp_pair_kernel<<<(unsigned)pblocks, kPairThreads, psmem, static_cast<cudaStream_t>(stream)>>>(p, tw0, tw1, int(wp));
            WT_LAUNCHED();
            return 0;
        }
    }
    const size_t smem = (size_t((d->n_weights + 31) & ~31) + size_t(3) * maxw * kLd) * sizeof(float);
    WT_REQUIRE(smem <= 220 * 1024, "ResMLP too large for the shared-memory kernel");
    static SmemOptIn opt;
    WT_CHECK_CUDA(opt_in_smem(resmlp_kernel, opt, smem));
    const long long blocks = (n + kMlpThreads - 1) / kMlpThreads;
    resmlp_kernel<<<(unsigned)blocks, kMlpThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    WT_LAUNCHED();
    return 0;
}
