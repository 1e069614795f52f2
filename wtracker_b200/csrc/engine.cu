// Detector engine: executes a program of fused ops (wt_op) over caller-owned NHWC buffers.
// Host-side bookkeeping only — all device memory (weights, workspace) belongs to the caller.
#include "../../include/wtracker_b200.h"
#include "conv.cuh"

#include <stdlib.h>

#include <vector>

namespace wt {

static int64_t dtype_size(int dt) { return dt == WT_DT_F32 ? 4 : (dt == WT_DT_U8 ? 1 : 2); }
static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static int64_t buf_align() {
    static const int64_t a = knob("WT_BUF_ALIGN", 1024);
    return a;
}
static int64_t buf_bytes(const wt_buf& b, int batch) {
    return align_up(int64_t(batch) * b.h * b.w * b.c * dtype_size(b.dtype), buf_align());
}

}  // namespace wt

struct wt_engine {
    std::vector<wt_buf> bufs;
    std::vector<void*> buf_ptr;
    std::vector<wt_op> ops;
    std::vector<wt::ConvDesc> conv_desc;      // per op (valid for CONV)
    std::vector<wt::ConvTcPlan*> conv_plan;   // per op (tcgen05 path)
    int batch = 0;
    int conv_impl = 0;
    int sm_count = 0;
    const uint8_t* weights = nullptr;
    int64_t weight_bytes = 0;
    // lanes (wt_op.lane): side stream for lane-1 ops, fork / join events, one event per op that an op of the other
    // lane depends on, and per op the latest earlier op of the other lane it has a buffer hazard with (-1 = none)
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    std::vector<cudaEvent_t> done_ev;
    std::vector<int> wait_on;
    bool lanes = false;
};

using namespace wt;

static TensorView view_of(const wt_engine* e, int id, int coff) {
    TensorView v;
    const wt_buf& b = e->bufs[id];
    v.base = e->buf_ptr[id];
    v.h = b.h;
    v.w = b.w;
    v.ctot = b.c;
    v.coff = coff;
    v.dtype = b.dtype;
    return v;
}

extern "C" int64_t wt_engine_workspace_bytes(const wt_buf* bufs, int n_bufs, int batch) {
    int64_t total = 0;
    for (int i = 0; i < n_bufs; ++i) total += buf_bytes(bufs[i], batch);
    return total + buf_align();
}

extern "C" void wt_engine_destroy(wt_engine* e) {
    if (!e) return;
    for (ConvTcPlan* p : e->conv_plan)
        if (p) conv_tc_plan_destroy(p);
    for (cudaEvent_t ev : e->done_ev)
        if (ev) cudaEventDestroy(ev);
    if (e->fork_ev) cudaEventDestroy(e->fork_ev);
    if (e->join_ev) cudaEventDestroy(e->join_ev);
    if (e->side) cudaStreamDestroy(e->side);
    delete e;
}

namespace {
struct ChanRange {
    int buf, lo, hi;
};
// channel ranges an op reads / writes (whole spatial extent; images are never split between ops)
void op_ranges(const wt_op& o, std::vector<ChanRange>& rd, std::vector<ChanRange>& wr) {
    rd.clear();
    wr.clear();
    switch (o.kind) {
        case WT_OP_CONV:
            rd.push_back({o.src, o.src_coff, o.src_coff + o.cin});
            if (o.res >= 0) rd.push_back({o.res, o.res_coff, o.res_coff + o.cout});
            if (o.add_buf >= 0) rd.push_back({o.add_buf, o.add_coff, o.add_coff + o.cout});
            if (o.chain_w_off >= 0 && o.cat_buf >= 0) rd.push_back({o.cat_buf, o.cat_coff, o.cat_coff + o.cat_c});
            if (o.dot_off >= 0) wr.push_back({o.dst, 0, 1});
            else wr.push_back({o.dst, o.dst_coff, o.dst_coff + ((o.chain_w_off >= 0 && o.cat_buf >= 0) ? o.chain_cout : o.cout)});
            break;
        case WT_OP_CONV0:
            rd.push_back({o.src, 0, 1});
            wr.push_back({o.dst, o.dst_coff, o.dst_coff + o.cout});
            break;
        default:   // WT_OP_SPPF_POOL
            rd.push_back({o.src, o.src_coff, o.src_coff + o.cin});
            wr.push_back({o.dst, o.dst_coff, o.dst_coff + 3 * o.cin});
    }
}
bool overlap(const std::vector<ChanRange>& a, const std::vector<ChanRange>& b) {
    for (const ChanRange& x : a)
        for (const ChanRange& y : b)
            if (x.buf == y.buf && x.lo < y.hi && y.lo < x.hi) return true;
    return false;
}
}  // namespace

extern "C" int wt_engine_create(const wt_buf* bufs, int n_bufs, const wt_op* ops, int n_ops, int batch,
                                const void* weights, int64_t weight_bytes, void* workspace, int64_t workspace_bytes,
                                int conv_impl, wt_engine** out) {
    WT_REQUIRE(bufs && ops && weights && workspace && out, "null argument");
    WT_REQUIRE(batch >= 1, "batch must be positive");
    WT_REQUIRE(workspace_bytes >= wt_engine_workspace_bytes(bufs, n_bufs, batch), "workspace too small");
    int cc_major = 0, sm = 0;
    if (wt_device_info(&sm, &cc_major, nullptr)) return 1;
    WT_REQUIRE(cc_major == 10, "wtracker_b200 needs an sm_100 (Blackwell B200) device");

    wt_engine* e = new wt_engine();
    e->batch = batch;
    e->conv_impl = conv_impl;
    e->sm_count = sm;
    e->weights = static_cast<const uint8_t*>(weights);
    e->weight_bytes = weight_bytes;
    e->bufs.assign(bufs, bufs + n_bufs);
    e->ops.assign(ops, ops + n_ops);
    uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<int64_t>(workspace), buf_align()));
    for (int i = 0; i < n_bufs; ++i) {
        e->buf_ptr.push_back(ws);
        ws += buf_bytes(bufs[i], batch);
    }
    e->conv_desc.resize(n_ops);
    e->conv_plan.assign(n_ops, nullptr);

    auto fail = [&](const char* msg, int i) {
        set_error(std::string(msg) + " (op " + std::to_string(i) + ")");
        wt_engine_destroy(e);
        return 1;
    };
    for (int i = 0; i < n_ops; ++i) {
        const wt_op& o = ops[i];
        if (o.src < 0 || o.src >= n_bufs || o.dst < 0 || o.dst >= n_bufs) return fail("buffer id out of range", i);
        if (o.res >= n_bufs) return fail("residual buffer id out of range", i);
        if (o.kind == WT_OP_CONV) {
            if (o.w_off < 0 || o.b_off < 0 || o.w_off + int64_t(o.cout) * o.k * o.k * o.cin * 2 > weight_bytes ||
                o.b_off + int64_t(o.cout) * 4 > weight_bytes)
                return fail("weight offsets out of range", i);
            if (o.w_off % 16 != 0 || o.b_off % 4 != 0) return fail("weight offsets must be 16/4-byte aligned", i);
            ConvDesc& d = e->conv_desc[i];
            d.src = view_of(e, o.src, o.src_coff);
            d.dst = view_of(e, o.dst, o.dst_coff);
            if (o.res >= 0) d.res = view_of(e, o.res, o.res_coff);
            else d.res = TensorView{nullptr, 0, 0, 0, 0, 0};
            d.add = TensorView{nullptr, 0, 0, 0, 0, 0};
            if (o.add_buf >= 0) {
                if (o.add_buf >= n_bufs) return fail("addend buffer id out of range", i);
                const wt_buf& ab = e->bufs[o.add_buf];
                if (ab.dtype != WT_DT_F32 || ab.h * 2 != e->bufs[o.dst].h || ab.w * 2 != e->bufs[o.dst].w ||
                    o.add_coff < 0 || o.add_coff + o.cout > ab.c || o.add_coff % 4 != 0 || ab.c % 4 != 0)
                    return fail("the upsampled addend must be an f32 buffer of half the destination size", i);
                d.add = view_of(e, o.add_buf, o.add_coff);
            }
            d.cin = o.cin;
            d.cout = o.cout;
            d.k = o.k;
            d.stride = o.stride;
            d.act = o.act;
            d.w = reinterpret_cast<const __nv_bfloat16*>(e->weights + o.w_off);
            d.bias = reinterpret_cast<const float*>(e->weights + o.b_off);
            d.dot_w = nullptr;
            if (o.dot_off >= 0) {
                if (o.dot_off % 4 != 0 || o.dot_off + int64_t(o.cout + 1) * 4 > weight_bytes)
                    return fail("dot-head weights out of range", i);
                if (e->bufs[o.dst].dtype != WT_DT_F32 || e->bufs[o.dst].c != 1 || o.dst_coff != 0 || o.res >= 0)
                    return fail("a dot-head conv writes a 1-channel f32 buffer and has no residual", i);
                d.dot_w = reinterpret_cast<const float*>(e->weights + o.dot_off);
            }
            d.chain_w = nullptr;
            d.chain_bias = nullptr;
            d.chain_act = 0;
            d.cat = TensorView{nullptr, 0, 0, 0, 0, 0};
            d.cat_c = 0;
            d.chain_cout = 0;
            if (o.chain_w_off >= 0) {
                const bool cat = o.cat_buf >= 0;
                const int64_t c2 = cat ? o.chain_cout : o.cout, k2 = cat ? int64_t(o.cat_c) + o.cout : o.cout;
                if (cat) {
                    if (o.cat_buf >= n_bufs || o.cat_c <= 0 || o.chain_cout <= 0 || o.cat_coff < 0 ||
                        o.cat_coff + o.cat_c > e->bufs[o.cat_buf].c)
                        return fail("concat-chain slice out of range", i);
                    d.cat = view_of(e, o.cat_buf, o.cat_coff);
                    d.cat_c = o.cat_c;
                    d.chain_cout = o.chain_cout;
                }
                if (o.chain_w_off % 16 != 0 || o.chain_b_off < 0 || o.chain_b_off % 4 != 0 ||
                    o.chain_w_off + c2 * k2 * 2 > weight_bytes || o.chain_b_off + c2 * 4 > weight_bytes)
                    return fail("chained conv weights out of range", i);
                if (conv_impl != 0) return fail("chained convs exist on the tcgen05 path only", i);
                d.chain_w = reinterpret_cast<const __nv_bfloat16*>(e->weights + o.chain_w_off);
                d.chain_bias = reinterpret_cast<const float*>(e->weights + o.chain_b_off);
                d.chain_act = o.chain_act;
            }
            d.batch = batch;
            if (o.src_coff + o.cin > e->bufs[o.src].c ||
                (!d.dot_w && o.dst_coff + (d.cat.base ? d.chain_cout : o.cout) > e->bufs[o.dst].c))
                return fail("channel slice exceeds buffer", i);
            if (conv_impl == 0) {
                if (conv_tc_plan_create(d, &e->conv_plan[i])) {
                    std::string m = wt_last_error();
                    return fail(m.c_str(), i);
                }
            }
        } else if (o.kind == WT_OP_CONV0) {
            if (e->bufs[o.src].dtype != WT_DT_U8 || e->bufs[o.src].c != 1) return fail("conv0 source must be u8 grey", i);
            if (o.w_off + int64_t(o.cout) * 9 * 4 > weight_bytes || o.b_off + int64_t(o.cout) * 4 > weight_bytes)
                return fail("weight offsets out of range", i);
            // chain_w_off of a CONV0 op: the bf16 [32][32] hi | lo weight matrix of the tcgen05 form (-1: CUDA cores)
            static const int c0tc_env = knob("WT_CONV0_TC", 1);
            if (o.chain_w_off >= 0 && conv_impl == 0 && c0tc_env && o.cout == 32 && e->bufs[o.src].w % 16 == 0) {
                if (o.chain_w_off % 16 != 0 || o.chain_w_off + 2048 > weight_bytes) return fail("conv0 weight matrix out of range", i);
                if (conv0_tc_plan_create(static_cast<const uint8_t*>(e->buf_ptr[o.src]), e->bufs[o.src].h, e->bufs[o.src].w,
                                         reinterpret_cast<const __nv_bfloat16*>(e->weights + o.chain_w_off),
                                         reinterpret_cast<const float*>(e->weights + o.b_off), o.act,
                                         view_of(e, o.dst, o.dst_coff), batch, &e->conv_plan[i])) {
                    std::string m = wt_last_error();
                    return fail(m.c_str(), i);
                }
            }
        } else if (o.kind != WT_OP_SPPF_POOL) {
            return fail("unknown op kind", i);
        }
    }
    // lanes: cross-lane hazards (read-after-write, write-after-write, write-after-read) -> event waits
    static const int lanes_env = knob("WT_LANES", 1);
    e->wait_on.assign(n_ops, -1);
    e->done_ev.assign(n_ops, nullptr);
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i].lane != 0 && ops[i].lane != 1) return fail("lane must be 0 or 1", i);
        if (!lanes_env) e->ops[i].lane = 0;
        if (e->ops[i].lane == 1) e->lanes = true;
    }
    if (e->lanes) {
        std::vector<ChanRange> ri, wi, rj, wj;
        for (int i = 0; i < n_ops; ++i) {
            op_ranges(e->ops[i], ri, wi);
            for (int j = i - 1; j >= 0; --j) {
                if (e->ops[j].lane == e->ops[i].lane) continue;
                op_ranges(e->ops[j], rj, wj);
                if (overlap(wj, ri) || overlap(wj, wi) || overlap(rj, wi)) {
                    e->wait_on[i] = j;   // the latest one: the other lane's stream orders everything before it
                    break;
                }
            }
        }
        bool ok = cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&e->join_ev, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < n_ops && ok; ++i)
            if (e->wait_on[i] >= 0 && !e->done_ev[e->wait_on[i]])
                ok = cudaEventCreateWithFlags(&e->done_ev[e->wait_on[i]], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) return fail("could not create the side stream / events", 0);
    }
    *out = e;
    return 0;
}

extern "C" void* wt_engine_buffer(wt_engine* e, int id) {
    if (!e || id < 0 || id >= int(e->bufs.size())) return nullptr;
    return e->buf_ptr[id];
}

extern "C" int wt_engine_forward(wt_engine* e, int n, int first_op, int last_op, void* stream_) {
    WT_REQUIRE(e, "null engine");
    WT_REQUIRE(n >= 0 && n <= e->batch, "n exceeds the engine batch");
    WT_REQUIRE(first_op >= 0 && last_op <= int(e->ops.size()) && first_op <= last_op, "op range");
    cudaStream_t main_stream = static_cast<cudaStream_t>(stream_);
    bool used_side = false;
    int waited[2] = {-1, -1};   // per lane: the latest op of the other lane this call already waited for
    for (int i = first_op; i < last_op; ++i) {
        const wt_op& o = e->ops[i];
        cudaStream_t stream = o.lane ? e->side : main_stream;
        if (o.lane && !used_side) {   // fork: the side stream starts behind everything already queued by the caller
            WT_CHECK_CUDA(cudaEventRecord(e->fork_ev, main_stream));
            WT_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->fork_ev, 0));
            used_side = true;
        }
        if (e->wait_on[i] > waited[o.lane]) {   // (an event never recorded, op outside this call's range, is a no-op)
            WT_CHECK_CUDA(cudaStreamWaitEvent(stream, e->done_ev[e->wait_on[i]], 0));
            waited[o.lane] = e->wait_on[i];
        }
        int rc = 0;
        switch (o.kind) {
            case WT_OP_CONV:
                if (e->conv_impl == 0) rc = conv_tc_launch(e->conv_plan[i], n, e->sm_count, stream);
                else rc = conv_simt_launch(e->conv_desc[i], n, stream);
                break;
            case WT_OP_CONV0:
                if (e->conv_plan[i]) {
                    rc = conv_tc_launch(e->conv_plan[i], n, e->sm_count, stream);
                    break;
                }
                rc = conv0_launch(static_cast<const uint8_t*>(e->buf_ptr[o.src]), e->bufs[o.src].h, e->bufs[o.src].w,
                                  reinterpret_cast<const float*>(e->weights + o.w_off),
                                  reinterpret_cast<const float*>(e->weights + o.b_off), o.cout, o.act,
                                  view_of(e, o.dst, o.dst_coff), n, stream);
                break;
            case WT_OP_SPPF_POOL:
                rc = sppf_pool_launch(view_of(e, o.src, o.src_coff), view_of(e, o.dst, o.dst_coff), o.cin, n, stream);
                break;
            default:
                set_error("unknown op kind");
                rc = 1;
        }
        if (rc) return rc;
        if (e->done_ev[i]) WT_CHECK_CUDA(cudaEventRecord(e->done_ev[i], stream));
    }
    if (used_side) {   // join: the caller's stream continues after the side stream's last op
        WT_CHECK_CUDA(cudaEventRecord(e->join_ev, e->side));
        WT_CHECK_CUDA(cudaStreamWaitEvent(main_stream, e->join_ev, 0));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ self-test
namespace {
__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t h = uint32_t(i) * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn((float(h & 0xFFFF) / 32768.f - 1.f) * scale);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    uint32_t h = uint32_t(i) * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = (float(h & 0xFFFF) / 32768.f - 1.f) * scale;
}
__global__ void max_diff_kernel(const void* a, const void* b, size_t n, int f32, float* out_max, float* out_ref_max) {
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float x, y;
    if (f32) { x = static_cast<const float*>(a)[i]; y = static_cast<const float*>(b)[i]; }
    else { x = __bfloat162float(static_cast<const __nv_bfloat16*>(a)[i]); y = __bfloat162float(static_cast<const __nv_bfloat16*>(b)[i]); }
    float d = fabsf(x - y);
    if (!(d == d)) d = 1e30f;
    atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(d));
    atomicMax(reinterpret_cast<int*>(out_ref_max), __float_as_int(fabsf(y)));
}
}  // namespace

extern "C" int wt_selftest_conv(int batch, int h, int w, int cin, int cout, int k, int stride, int act,
                                int with_residual, int out_f32, int verbose, double* max_abs_diff) {
    int sm = 0, cc = 0;
    if (wt_device_info(&sm, &cc, nullptr)) return 1;
    const int ho = h / stride, wo = w / stride;
    // source / destination sit inside wider buffers at a channel offset to exercise slicing
    // (verbose bit 1: the source fills its buffer, which the stride-2 pixel-pair kernel requires)
    const bool tight = (verbose & 2) != 0;
    verbose &= 1;
    const int src_ct = tight ? cin : cin + 64, src_off = tight ? 0 : 32, dst_ct = cout + 64, dst_off = 32;
    const size_t n_src = size_t(batch) * h * w * src_ct, n_dst = size_t(batch) * ho * wo * dst_ct;
    const size_t n_w = size_t(cout) * k * k * cin;
    const int es = out_f32 ? 4 : 2;
    __nv_bfloat16 *d_src = nullptr, *d_w = nullptr, *d_res = nullptr;
    void *d_out_tc = nullptr, *d_out_ref = nullptr;
    float *d_bias = nullptr, *d_stats = nullptr;
    WT_CHECK_CUDA(cudaMalloc(&d_src, n_src * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_w, n_w * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_res, n_dst * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_out_tc, n_dst * es));
    WT_CHECK_CUDA(cudaMalloc(&d_out_ref, n_dst * es));
    WT_CHECK_CUDA(cudaMalloc(&d_bias, cout * 4));
    WT_CHECK_CUDA(cudaMalloc(&d_stats, 8));
    WT_CHECK_CUDA(cudaMemset(d_out_tc, 0, n_dst * es));
    WT_CHECK_CUDA(cudaMemset(d_out_ref, 0, n_dst * es));
    WT_CHECK_CUDA(cudaMemset(d_stats, 0, 8));
    fill_bf16<<<unsigned((n_src + 255) / 256), 256>>>(d_src, n_src, 1u, 1.0f);
    fill_bf16<<<unsigned((n_w + 255) / 256), 256>>>(d_w, n_w, 2u, 1.0f / sqrtf(float(k * k * cin)));
    fill_bf16<<<unsigned((n_dst + 255) / 256), 256>>>(d_res, n_dst, 3u, 1.0f);
    fill_f32<<<unsigned((cout + 255) / 256), 256>>>(d_bias, cout, 4u, 0.5f);
    ConvDesc d;
    d.src = TensorView{d_src, h, w, src_ct, src_off, WT_DT_BF16};
    d.dst = TensorView{d_out_tc, ho, wo, dst_ct, dst_off, out_f32 ? WT_DT_F32 : WT_DT_BF16};
    d.res = with_residual ? TensorView{d_res, ho, wo, dst_ct, dst_off, WT_DT_BF16} : TensorView{nullptr, 0, 0, 0, 0, 0};
    d.add = TensorView{nullptr, 0, 0, 0, 0, 0};
    d.cin = cin; d.cout = cout; d.k = k; d.stride = stride; d.act = act;
    d.w = d_w; d.bias = d_bias; d.dot_w = nullptr; d.batch = batch;
    d.chain_w = nullptr; d.chain_bias = nullptr; d.chain_act = 0;
    d.cat = TensorView{nullptr, 0, 0, 0, 0, 0}; d.cat_c = 0; d.chain_cout = 0;
    ConvTcPlan* plan = nullptr;
    int rc = conv_tc_plan_create(d, &plan);
    if (!rc) rc = conv_tc_launch(plan, batch, sm, 0);
    if (plan) conv_tc_plan_destroy(plan);
    if (!rc) {
        ConvDesc r = d;
        r.dst.base = d_out_ref;
        rc = conv_simt_launch(r, batch, 0);
    }
    float stats[2] = {0, 0};
    if (!rc) {
        max_diff_kernel<<<unsigned((n_dst + 255) / 256), 256>>>(d_out_tc, d_out_ref, n_dst, out_f32, d_stats, d_stats + 1);
        cudaError_t e1 = cudaDeviceSynchronize();
        if (e1 != cudaSuccess) { set_error(std::string("selftest kernel failed: ") + cudaGetErrorString(e1)); rc = 1; }
        else cudaMemcpy(stats, d_stats, 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_src); cudaFree(d_w); cudaFree(d_res); cudaFree(d_out_tc); cudaFree(d_out_ref); cudaFree(d_bias); cudaFree(d_stats);
    if (rc) return rc;
    if (max_abs_diff) *max_abs_diff = stats[0];
    if (verbose)
        printf("selftest_conv b%d %dx%d cin%d cout%d k%d s%d act%d res%d f32%d : max|diff| %.5g (ref max %.4g)\n", batch, h, w,
               cin, cout, k, stride, act, with_residual, out_f32, stats[0], stats[1]);
    return 0;
}

extern "C" int wt_selftest_conv_chain(int batch, int h, int w, int cin, int cout, int k, int stride, int verbose,
                                      double* max_abs_diff) {
    int sm = 0, cc = 0;
    if (wt_device_info(&sm, &cc, nullptr)) return 1;
    const int ho = h / stride, wo = w / stride;
    // the source fills its buffer (the stride-2 pixel-pair kernel requires it); the destination is a channel slice
    const int dst_ct = cout + 64, dst_off = 32;
    const size_t n_src = size_t(batch) * h * w * cin, n_mid = size_t(batch) * ho * wo * cout;
    const size_t n_dst = size_t(batch) * ho * wo * dst_ct, n_w = size_t(cout) * k * k * cin, n_w2 = size_t(cout) * cout;
    __nv_bfloat16 *d_src = nullptr, *d_w = nullptr, *d_w2 = nullptr, *d_mid = nullptr, *d_out = nullptr, *d_ref = nullptr;
    float *d_bias = nullptr, *d_bias2 = nullptr, *d_stats = nullptr;
    WT_CHECK_CUDA(cudaMalloc(&d_src, n_src * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_w, n_w * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_w2, n_w2 * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_mid, n_mid * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_out, n_dst * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_ref, n_dst * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_bias, cout * 4));
    WT_CHECK_CUDA(cudaMalloc(&d_bias2, cout * 4));
    WT_CHECK_CUDA(cudaMalloc(&d_stats, 8));
    WT_CHECK_CUDA(cudaMemset(d_out, 0, n_dst * 2));
    WT_CHECK_CUDA(cudaMemset(d_ref, 0, n_dst * 2));
    WT_CHECK_CUDA(cudaMemset(d_stats, 0, 8));
    fill_bf16<<<unsigned((n_src + 255) / 256), 256>>>(d_src, n_src, 1u, 1.0f);
    fill_bf16<<<unsigned((n_w + 255) / 256), 256>>>(d_w, n_w, 2u, 1.0f / sqrtf(float(k * k * cin)));
    fill_bf16<<<unsigned((n_w2 + 255) / 256), 256>>>(d_w2, n_w2, 5u, 2.0f / sqrtf(float(cout)));
    fill_f32<<<unsigned((cout + 255) / 256), 256>>>(d_bias, cout, 4u, 0.5f);
    fill_f32<<<unsigned((cout + 255) / 256), 256>>>(d_bias2, cout, 6u, 0.5f);
    const TensorView none{nullptr, 0, 0, 0, 0, 0};
    ConvDesc a;   // the conv, written to the intermediate buffer
    a.src = TensorView{d_src, h, w, cin, 0, WT_DT_BF16};
    a.dst = TensorView{d_mid, ho, wo, cout, 0, WT_DT_BF16};
    a.res = none; a.add = none;
    a.cin = cin; a.cout = cout; a.k = k; a.stride = stride; a.act = WT_ACT_SILU;
    a.w = d_w; a.bias = d_bias; a.dot_w = nullptr; a.batch = batch;
    a.chain_w = nullptr; a.chain_bias = nullptr; a.chain_act = 0;
    a.cat = TensorView{nullptr, 0, 0, 0, 0, 0}; a.cat_c = 0; a.chain_cout = 0;
    ConvDesc b = a;   // the 1x1 conv on the intermediate buffer
    b.src = a.dst;
    b.dst = TensorView{d_ref, ho, wo, dst_ct, dst_off, WT_DT_BF16};
    b.cin = cout; b.k = 1; b.stride = 1; b.w = d_w2; b.bias = d_bias2;
    ConvDesc c = a;   // both in one launch
    c.dst = TensorView{d_out, ho, wo, dst_ct, dst_off, WT_DT_BF16};
    c.chain_w = d_w2; c.chain_bias = d_bias2; c.chain_act = WT_ACT_SILU;
    int rc = 0;
    for (const ConvDesc* d : {&a, &b, &c}) {
        ConvTcPlan* plan = nullptr;
        if (!rc) rc = conv_tc_plan_create(*d, &plan);
        if (!rc) rc = conv_tc_launch(plan, batch, sm, 0);
        if (plan) conv_tc_plan_destroy(plan);
    }
    float stats[2] = {0, 0};
    if (!rc) {
        max_diff_kernel<<<unsigned((n_dst + 255) / 256), 256>>>(d_out, d_ref, n_dst, 0, d_stats, d_stats + 1);
        cudaError_t e1 = cudaDeviceSynchronize();
        if (e1 != cudaSuccess) { set_error(std::string("selftest kernel failed: ") + cudaGetErrorString(e1)); rc = 1; }
        else cudaMemcpy(stats, d_stats, 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_src); cudaFree(d_w); cudaFree(d_w2); cudaFree(d_mid); cudaFree(d_out); cudaFree(d_ref);
    cudaFree(d_bias); cudaFree(d_bias2); cudaFree(d_stats);
    if (rc) return rc;
    if (max_abs_diff) *max_abs_diff = stats[0];
    if (verbose)
        printf("selftest_conv_chain b%d %dx%d cin%d cout%d k%d s%d : max|diff| %.5g (ref max %.4g)\n", batch, h, w, cin, cout,
               k, stride, stats[0], stats[1]);
    return 0;
}

extern "C" int wt_selftest_conv_cat(int batch, int h, int w, int verbose, double* max_abs_diff) {
    // C2f exit: b = y1 + SiLU(conv3x3(b1)); out = SiLU(W2 * [y0 | y1 | b] + b2), as ONE concat-chain launch and as two
    // tcgen05 launches through the concat buffer; the results must be identical.
    int sm = 0, cc = 0;
    if (wt_device_info(&sm, &cc, nullptr)) return 1;
    const int c = 32, cat_ct = 3 * c, out_ct = 2 * c + 64, out_off = 32;
    const size_t n_px = size_t(batch) * h * w;
    const size_t n_cat = n_px * cat_ct, n_b1 = n_px * c, n_out = n_px * out_ct, n_w = size_t(c) * 9 * c, n_w2 = size_t(2 * c) * cat_ct;
    __nv_bfloat16 *d_cat = nullptr, *d_b1 = nullptr, *d_w = nullptr, *d_w2 = nullptr, *d_out = nullptr, *d_ref = nullptr;
    float *d_bias = nullptr, *d_bias2 = nullptr, *d_stats = nullptr;
    WT_CHECK_CUDA(cudaMalloc(&d_cat, n_cat * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_b1, n_b1 * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_w, n_w * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_w2, n_w2 * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_out, n_out * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_ref, n_out * 2));
    WT_CHECK_CUDA(cudaMalloc(&d_bias, c * 4));
    WT_CHECK_CUDA(cudaMalloc(&d_bias2, 2 * c * 4));
    WT_CHECK_CUDA(cudaMalloc(&d_stats, 8));
    WT_CHECK_CUDA(cudaMemset(d_out, 0, n_out * 2));
    WT_CHECK_CUDA(cudaMemset(d_ref, 0, n_out * 2));
    WT_CHECK_CUDA(cudaMemset(d_stats, 0, 8));
    fill_bf16<<<unsigned((n_cat + 255) / 256), 256>>>(d_cat, n_cat, 1u, 1.0f);
    fill_bf16<<<unsigned((n_b1 + 255) / 256), 256>>>(d_b1, n_b1, 7u, 1.0f);
    fill_bf16<<<unsigned((n_w + 255) / 256), 256>>>(d_w, n_w, 2u, 1.0f / sqrtf(float(9 * c)));
    fill_bf16<<<unsigned((n_w2 + 255) / 256), 256>>>(d_w2, n_w2, 5u, 2.0f / sqrtf(float(cat_ct)));
    fill_f32<<<1, 256>>>(d_bias, c, 4u, 0.5f);
    fill_f32<<<1, 256>>>(d_bias2, 2 * c, 6u, 0.5f);
    const TensorView none{nullptr, 0, 0, 0, 0, 0};
    ConvDesc a;   // the bottleneck's second conv into the third slice of the concat buffer, residual = second slice
    a.src = TensorView{d_b1, h, w, c, 0, WT_DT_BF16};
    a.dst = TensorView{d_cat, h, w, cat_ct, 2 * c, WT_DT_BF16};
    a.res = TensorView{d_cat, h, w, cat_ct, c, WT_DT_BF16};
    a.add = none; a.cat = none; a.cat_c = 0; a.chain_cout = 0;
    a.cin = c; a.cout = c; a.k = 3; a.stride = 1; a.act = WT_ACT_SILU;
    a.w = d_w; a.bias = d_bias; a.dot_w = nullptr; a.batch = batch;
    a.chain_w = nullptr; a.chain_bias = nullptr; a.chain_act = 0;
    ConvDesc b = a;   // cv2 over the whole concat buffer
    b.src = TensorView{d_cat, h, w, cat_ct, 0, WT_DT_BF16};
    b.dst = TensorView{d_ref, h, w, out_ct, out_off, WT_DT_BF16};
    b.res = none;
    b.cin = cat_ct; b.cout = 2 * c; b.k = 1; b.w = d_w2; b.bias = d_bias2;
    ConvDesc f = a;   // both in one launch; the third slice of the concat buffer is not written
    f.dst = TensorView{d_out, h, w, out_ct, out_off, WT_DT_BF16};
    f.cat = TensorView{d_cat, h, w, cat_ct, 0, WT_DT_BF16};
    f.cat_c = 2 * c; f.chain_cout = 2 * c;
    f.chain_w = d_w2; f.chain_bias = d_bias2; f.chain_act = WT_ACT_SILU;
    int rc = 0;
    for (const ConvDesc* d : {&f, &a, &b}) {   // the fused launch first: it must not depend on slice 3 being written
        ConvTcPlan* plan = nullptr;
        if (!rc) rc = conv_tc_plan_create(*d, &plan);
        if (!rc) rc = conv_tc_launch(plan, batch, sm, 0);
        if (plan) conv_tc_plan_destroy(plan);
    }
    float stats[2] = {0, 0};
    if (!rc) {
        max_diff_kernel<<<unsigned((n_out + 255) / 256), 256>>>(d_out, d_ref, n_out, 0, d_stats, d_stats + 1);
        cudaError_t e1 = cudaDeviceSynchronize();
        if (e1 != cudaSuccess) { set_error(std::string("selftest kernel failed: ") + cudaGetErrorString(e1)); rc = 1; }
        else cudaMemcpy(stats, d_stats, 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_cat); cudaFree(d_b1); cudaFree(d_w); cudaFree(d_w2); cudaFree(d_out); cudaFree(d_ref);
    cudaFree(d_bias); cudaFree(d_bias2); cudaFree(d_stats);
    if (rc) return rc;
    if (max_abs_diff) *max_abs_diff = stats[0];
    if (verbose)
        printf("selftest_conv_cat b%d %dx%d : max|diff| %.5g (ref max %.4g)\n", batch, h, w, stats[0], stats[1]);
    return 0;
}
