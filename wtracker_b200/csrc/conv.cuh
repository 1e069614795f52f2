// Internal description of one convolution over NHWC buffers, shared by the tcgen05 implicit-GEMM
// kernel (conv_tcgen05.cu), the scalar validation kernel and the small data-movement ops (ops_simt.cu).
#pragma once
#include "common.cuh"

namespace wt {

struct TensorView {      // NHWC view of (a channel slice of) an activation buffer
    void* base;          // buffer base address ([batch][h][w][ctot])
    int h, w;            // spatial size
    int ctot;            // channels physically present in the buffer
    int coff;            // first channel of the slice
    int dtype;           // WT_DT_*
};

struct ConvDesc {
    TensorView src, dst, res;   // res.base == nullptr -> no residual
    TensorView add;             // add.base == nullptr -> none; else f32 half-resolution pre-activation addend (wt_op.add_buf)
    int cin, cout, k, stride, act;
    const __nv_bfloat16* w;     // [cout][k][k][cin]
    const float* bias;          // [cout]
    const float* dot_w;         // fused 1-channel 1x1 head (wt_op.dot_off): f32 [cout + 1], or nullptr
    // chained 1x1 conv (wt_op.chain_w_off): dst receives act2(W2 * bf16(act(conv(src))) + b2), cout -> cout channels;
    // the intermediate map only ever exists as bf16 tiles in shared memory.  nullptr = none.
    const __nv_bfloat16* chain_w;   // [cout][cout]   (concat chain: [chain_cout][cat_c + cout])
    const float* chain_bias;        // [cout]         (concat chain: [chain_cout])
    int chain_act;
    // concat chain (wt_op.cat_buf): the chained 1x1 conv runs over concat(cat slice, this conv's output); cat.base ==
    // nullptr -> plain chain.  The conv's residual must be the upper half of the cat slice (C2f: y1 of [y0 | y1]).
    TensorView cat;
    int cat_c, chain_cout;
    int batch;                  // images the buffers were sized for
};

// ---- tcgen05 path -------------------------------------------------------------------------
struct ConvTcPlan;   // opaque: tensor maps + tiling, built once per layer
int conv_tc_plan_create(const ConvDesc& d, ConvTcPlan** out);
void conv_tc_plan_destroy(ConvTcPlan* p);
int conv_tc_launch(const ConvTcPlan* p, int n_images, int sm_count, cudaStream_t stream);
// first layer on the tensor cores: wmat = bf16 [32][32] (hi taps | 0 | lo taps | 0 per output channel); launched with
// conv_tc_launch like every other plan
int conv0_tc_plan_create(const uint8_t* src, int h, int w, const __nv_bfloat16* wmat, const float* bias, int act,
                         const TensorView& dst, int batch, ConvTcPlan** out);

// ---- scalar validation path ---------------------------------------------------------------
int conv_simt_launch(const ConvDesc& d, int n_images, cudaStream_t stream);

// ---- other ops ----------------------------------------------------------------------------
// first layer: u8 grey [n][h][w] -> bf16 NHWC [n][h/2][w/2][cout], 3x3 s2 p1, fp32 weights
int conv0_launch(const uint8_t* src, int h, int w, const float* w9, const float* bias, int cout, int act,
                 const TensorView& dst, int n_images, cudaStream_t stream);
// SPPF: three chained 5x5/s1/p2 max-pools of src slice -> dst slices at dst.coff, +c, +2c
int sppf_pool_launch(const TensorView& src, const TensorView& dst, int c, int n_images, cudaStream_t stream);

}  // namespace wt
