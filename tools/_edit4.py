def edit(path, pairs):
    s=open(path).read()
    for old,new in pairs:
        assert old in s, (path, old[:70])
        s=s.replace(old,new)
    open(path,'w').write(s)

edit('wtracker_b200/csrc/conv_tcgen05.cu', [
('''    int act, has_res, out_f32;
    int num_tiles;''','''    int act, has_res, out_f32;
    int num_tiles;
    // fused 1-channel 1x1 head (wt_op.dot_off): out pixel = sum_c act(conv)[c] * dot_w[c] + dot_w[cout]
    const float* dot_w;
    float* dot_out;                  // f32 [n][out_h][out_w]
    int out_w, out_h, n_images;'''),
# load dot weights next to the bias
('''    for (int i = threadIdx.x; i < p.cout; i += kThreads) sBias[i] = __ldg(p.bias + i);
''','''    for (int i = threadIdx.x; i < p.cout; i += kThreads) sBias[i] = __ldg(p.bias + i);
    if (p.dot_w)   // dot weights + bias behind the conv bias (host checks 2 * cout + 1 <= kMaxCout)
        for (int i = threadIdx.x; i <= p.cout; i += kThreads) sBias[p.cout + i] = __ldg(p.dot_w + i);
'''),
# epilogue: dot path
('''        ptx::mbar_wait(&tfull_bar[g], aphase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN;

#pragma unroll 1
        for (int sub = 0; sub < BN / 32; ++sub) {
            const int sub_in_unit''','''        ptx::mbar_wait(&tfull_bar[g], aphase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN;

        if (p.dot_w) {
            // fused class-logit head: this thread owns one pixel and all its output channels (n_blocks == 1)
            const float* dw = sBias + p.cout;
            float dot = 0.f;
#pragma unroll 1
            for (int sub = 0; sub < BN / 32; ++sub) {
                uint32_t acc[32];
                ptx::tmem_ld_32x32(t_row + sub * 32, acc);
                ptx::tmem_ld_wait();
                if (sub == BN / 32 - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[g]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(acc[j]) + bias[sub * 32 + j];
                    if (p.act == WT_ACT_SILU) {
                        v = __fdividef(v, 1.0f + __expf(-v));
                    } else if (p.act == kActSiluTanh) {
                        const float h = 0.5f * v;
                        v = fmaf(h, tanh_fast(h), h);
                    }
                    dot = fmaf(v, dw[sub * 32 + j], dot);
                }
            }
            const int px = x0 + row % p.tw;
            const int py = y0 + (row / p.tw) % p.th;
            const int pn = n0 + row / (p.tw * p.th);
            if (px < p.out_w && py < p.out_h && pn < p.n_images)
                p.dot_out[(size_t(pn) * p.out_h + py) * p.out_w + px] = dot + dw[p.cout];
            continue;
        }

#pragma unroll 1
        for (int sub = 0; sub < BN / 32; ++sub) {
            const int sub_in_unit'''),
# plan checks
('''    const bool out_f32 = d.dst.dtype == WT_DT_F32;
    WT_REQUIRE(!(out_f32 && d.res.base), "residual only with bf16 output");''','''    const bool out_f32 = d.dst.dtype == WT_DT_F32;
    WT_REQUIRE(!(out_f32 && d.res.base), "residual only with bf16 output");
    if (d.dot_w) {
        WT_REQUIRE(bn == d.cout && 2 * d.cout + 1 <= kMaxCout, "a dot-head conv needs all channels in one N tile");
        WT_REQUIRE(out_f32 && d.dst.ctot == 1 && !d.res.base, "a dot-head conv writes a 1-channel f32 buffer");
    }'''),
('''    WT_REQUIRE((d.dst.ctot * (out_f32 ? 4 : 2)) % 16 == 0, "destination channel alignment");''','''    WT_REQUIRE(d.dot_w || (d.dst.ctot * (out_f32 ? 4 : 2)) % 16 == 0, "destination channel alignment");'''),
('''    p.bias = d.bias;
    p.num_tiles = 0;''','''    p.bias = d.bias;
    p.dot_w = d.dot_w;
    p.dot_out = d.dot_w ? static_cast<float*>(d.dst.base) : nullptr;
    p.out_w = wo;
    p.out_h = ho;
    p.n_images = 0;
    p.num_tiles = 0;'''),
('''    {
        const int es = out_f32 ? 4 : 2;
        const int unit_ch = out_f32 ? 32 : (bn == 32 ? 32 : 64);''','''    if (d.dot_w) {
        p.tmD = p.tmA[0];   // never used: the dot head stores with plain st.global
        p.tmR = p.tmA[0];
    } else {
        const int es = out_f32 ? 4 : 2;
        const int unit_ch = out_f32 ? 32 : (bn == 32 ? 32 : 64);'''),
('''    prm.num_tiles = pl->pix_per_image_tiles * tiles_n * prm.n_blocks;''','''    prm.num_tiles = pl->pix_per_image_tiles * tiles_n * prm.n_blocks;
    prm.n_images = n_images;'''),
])

# SIMT validation: dot head
edit('wtracker_b200/csrc/ops_simt.cu', [
('''// ------------------------------------------------------------------ first layer''','''// scalar validation of a conv with a fused 1-channel head: one thread per pixel walks all channels
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

// ------------------------------------------------------------------ first layer'''),
('''    WT_REQUIRE(d.src.dtype == WT_DT_BF16, "conv input must be bf16");
    const long long total = (long long)n_images * d.dst.h * d.dst.w * d.cout;
    if (total == 0) return 0;
    const int threads = 256;''','''    WT_REQUIRE(d.src.dtype == WT_DT_BF16, "conv input must be bf16");
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
    const int threads = 256;'''),
])

# program.py
edit('wtracker_b200/detector/program.py', [
('''    def conv(name: str, src: tuple[int, int], dst: tuple[int, int], res: tuple[int, int] | None = None):
        s = specs[name]
        w_off, b_off = add_weights(s)
        p.ops.append(dict(kind=L.WT_OP_CONV, name=name, src=src[0], src_coff=src[1], dst=dst[0], dst_coff=dst[1],
                          res=-1 if res is None else res[0], res_coff=0 if res is None else res[1], cin=s.cin,
                          cout=s.cout, k=s.k, stride=s.stride, act=L.WT_ACT_SILU if s.bn_act else L.WT_ACT_NONE,
                          w_off=w_off, b_off=b_off))''','''    def conv(name: str, src: tuple[int, int], dst: tuple[int, int], res: tuple[int, int] | None = None,
             dot: tuple[torch.Tensor, float] | None = None):
        s = specs[name]
        w_off, b_off = add_weights(s)
        dot_off = -1
        if dot is not None:     # fused 1-channel 1x1 head: f32 [cout] weights then the bias
            dot_off = _align(p.blob, 16)
            p.blob.extend(torch.cat([dot[0].reshape(-1).float(), torch.tensor([dot[1]])]).numpy().tobytes())
        p.ops.append(dict(kind=L.WT_OP_CONV, name=name, src=src[0], src_coff=src[1], dst=dst[0], dst_coff=dst[1],
                          res=-1 if res is None else res[0], res_coff=0 if res is None else res[1], cin=s.cin,
                          cout=s.cout, k=s.k, stride=s.stride, act=L.WT_ACT_SILU if s.bn_act else L.WT_ACT_NONE,
                          w_off=w_off, b_off=b_off, dot_off=dot_off))'''),
('''        u1 = new_buf(f"head{lvl}.cls1", down, arch.cls_c)
        u2 = new_buf(f"head{lvl}.cls2", down, arch.cls_c)
''','''        u1 = new_buf(f"head{lvl}.cls1", down, arch.cls_c)
        logit = new_buf(f"head{lvl}.cls", down, 1, L.WT_DT_F32)
'''),
('''        conv(f"model.22.cv3.{lvl}.1", (u1, 0), (u2, 0))
        wc, bc = folded_conv(sd, specs[f"model.22.cv3.{lvl}.2"])
        cw_off = _align(p.blob, 16)
        p.blob.extend(wc.reshape(-1).to(torch.bfloat16).view(torch.int16).numpy().tobytes())
        p.head.append(dict(box=box, cls_feat=u2, cls_w_off=cw_off, cls_b=float(bc.reshape(-1)[0]), h=net_h // down,
                           w=net_w // down, stride=STRIDES[lvl]))''','''        # the class branch ends in a 1x1 conv with nc = 1 output: a dot product fused into the epilogue of
        # the conv before it (fp32 weights on the fp32 accumulator, the 128-channel feature map is never stored)
        wc, bc = folded_conv(sd, specs[f"model.22.cv3.{lvl}.2"])
        conv(f"model.22.cv3.{lvl}.1", (u1, 0), (logit, 0), dot=(wc, float(bc.reshape(-1)[0])))
        p.head.append(dict(box=box, cls_logit=logit, h=net_h // down, w=net_w // down, stride=STRIDES[lvl]))'''),
("o[\"cin\"], o[\"cout\"], o[\"k\"], o[\"stride\"], o[\"act\"], o[\"w_off\"], o[\"b_off\"])","o[\"cin\"], o[\"cout\"], o[\"k\"], o[\"stride\"], o[\"act\"], o[\"w_off\"], o[\"b_off\"], o.get(\"dot_off\", -1))"),
('''  * the last 1x1 of the box branch writes fp32 (DFL is sensitive to logit rounding); the last 1x1
    of the class branch (cout = nc = 1) is a dot product fused into the decode kernel.''','''  * the last 1x1 of the box branch writes fp32 (DFL is sensitive to logit rounding); the last 1x1
    of the class branch (cout = nc = 1) is a dot product fused into the epilogue of the 3x3 conv
    before it (wt_op.dot_off), which then writes one fp32 logit per anchor.'''),
])
edit('wtracker_b200/detector/engine.py', [
('''                levels[i] = L.WtHeadLevel(self.buffer_ptr(h["box"]), self.buffer_ptr(h["cls_feat"]), None, h["h"],
                                          h["w"], h["stride"], L.WT_DT_F32, self.arch.cls_c,
                                          self.weights.data_ptr() + h["cls_w_off"], h["cls_b"])''','''                levels[i] = L.WtHeadLevel(self.buffer_ptr(h["box"]), None, self.buffer_ptr(h["cls_logit"]), h["h"],
                                          h["w"], h["stride"], L.WT_DT_F32, 0, None, 0.0)'''),
])
