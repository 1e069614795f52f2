import re
def edit(path, pairs):
    s=open(path).read()
    for old,new in pairs:
        assert old in s, (path, old[:70])
        s=s.replace(old,new)
    open(path,'w').write(s)

# ---------------- header
edit('include/wtracker_b200.h', [
('#define WT_ABI_VERSION 1','#define WT_ABI_VERSION 2'),
('''    int64_t w_off, b_off;        /* byte offsets into the weight blob:                         */
                                 /*   CONV : bf16 [cout][k][k][cin], f32 bias[cout]            */
                                 /*   CONV0: f32 [cout][3][3] (grey-folded, /255 folded), f32 bias */
} wt_op;''','''    int64_t w_off, b_off;        /* byte offsets into the weight blob:                         */
                                 /*   CONV : bf16 [cout][k][k][cin], f32 bias[cout]            */
                                 /*   CONV0: f32 [cout][3][3] (grey-folded, /255 folded), f32 bias */
    int64_t dot_off;             /* CONV only, -1 = none.  Otherwise the byte offset of        */
                                 /* f32 [cout + 1] = weights w[cout] then a bias b of a FUSED   */
                                 /* following 1x1 convolution with ONE output channel (the     */
                                 /* class-logit conv of the head, nc = 1): the activated output */
                                 /* is not stored; dst (f32, c = 1) receives                    */
                                 /* sum_c out[c] * w[c] + b per pixel.  Needs cout <= 256.      */
} wt_op;'''),
])

# ---------------- _lib.py
edit('wtracker_b200/_lib.py', [
('''        ("w_off", C.c_int64), ("b_off", C.c_int64),
    ]''','''        ("w_off", C.c_int64), ("b_off", C.c_int64), ("dot_off", C.c_int64),
    ]'''),
('if handle.wt_abi_version() != 1:','if handle.wt_abi_version() != 2:'),
])
edit('__graft_entry__.py', [('assert lib.wt_abi_version() == 1','assert lib.wt_abi_version() == 2')])

# ---------------- conv.cuh
edit('wtracker_b200/csrc/conv.cuh', [
('''    const float* bias;          // [cout]
    int batch;                  // images the buffers were sized for''','''    const float* bias;          // [cout]
    const float* dot_w;         // fused 1-channel 1x1 head (wt_op.dot_off): f32 [cout + 1], or nullptr
    int batch;                  // images the buffers were sized for'''),
])

# ---------------- engine.cu
edit('wtracker_b200/csrc/engine.cu', [
('''            d.bias = reinterpret_cast<const float*>(e->weights + o.b_off);
            d.batch = batch;''','''            d.bias = reinterpret_cast<const float*>(e->weights + o.b_off);
            d.dot_w = nullptr;
            if (o.dot_off >= 0) {
                if (o.dot_off % 4 != 0 || o.dot_off + int64_t(o.cout + 1) * 4 > weight_bytes)
                    return fail("dot-head weights out of range", i);
                if (e->bufs[o.dst].dtype != WT_DT_F32 || e->bufs[o.dst].c != 1 || o.dst_coff != 0 || o.res >= 0)
                    return fail("a dot-head conv writes a 1-channel f32 buffer and has no residual", i);
                d.dot_w = reinterpret_cast<const float*>(e->weights + o.dot_off);
            }
            d.batch = batch;'''),
('''            if (o.src_coff + o.cin > e->bufs[o.src].c || o.dst_coff + o.cout > e->bufs[o.dst].c)
                return fail("channel slice exceeds buffer", i);''','''            if (o.src_coff + o.cin > e->bufs[o.src].c || (!d.dot_w && o.dst_coff + o.cout > e->bufs[o.dst].c))
                return fail("channel slice exceeds buffer", i);'''),
('''    d.w = d_w; d.bias = d_bias; d.batch = batch;''','''    d.w = d_w; d.bias = d_bias; d.dot_w = nullptr; d.batch = batch;'''),
('extern "C" int wt_abi_version','extern "C" int wt_abi_version') if False else ('',''),
])
