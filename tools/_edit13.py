p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new,cnt=None):
    global s
    c=s.count(old)
    assert c>=1, old[:90]
    if cnt is not None: assert c==cnt,(c,old[:90])
    s=s.replace(old,new)

rep('''constexpr int kHaloW = 10, kHaloH = 18;
// bytes of one halo stage for a K block of BK channels (rows of 2 * BK bytes), 1024-byte aligned
__host__ __device__ constexpr int halo_a_bytes(int bk) { return ((kHaloW * kHaloH * bk * 2 + 1023) / 1024) * 1024; }
''','''constexpr int kHaloW = 10, kHaloH = 18;
// Stride-2 form (S2 = 1, 32 input channels, e.g. layer 1): the input is viewed as PAIRS of x-adjacent pixels
// (2 x 32 channels = one 128-byte SWIZZLE_128B row).  The tile's 33 x 9 pair-row patch is loaded once; output
// pixel (ty, tx) at tap (kh, kw) reads input row 2*ty + kh of the patch and input x = 2*ox + kw - 1, i.e. the
// second half of pair tx (kw = 0), the first half of pair tx + 1 (kw = 1) or its second half (kw = 2): a
// descriptor start offset of (kh * 9 + (kw != 0)) rows + (kw != 1) * 64 bytes, 8 consecutive rows per tile row,
// and a stride of two patch rows (18 rows) between tile rows.  One fill instead of nine strided ones.
constexpr int kS2HaloW = 9, kS2HaloH = 33;
// bytes of one halo stage (1024-byte aligned): S2 = 0: K block of BK channels (rows of 2 * BK bytes)
__host__ __device__ constexpr int halo_rows(int s2) { return s2 ? kS2HaloW * kS2HaloH : kHaloW * kHaloH; }
__host__ __device__ constexpr int halo_a_bytes(int bk, int s2) {
    return ((halo_rows(s2) * (s2 ? 128 : bk * 2) + 1023) / 1024) * 1024;
}
''')
rep('''template <int BN, int BK, int CG>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kABytes = halo_a_bytes(BK);
    static constexpr int kBBytes = (BN / CG) * kRowBytes;    // a CTA of a pair holds BN / 2 weight rows
};''','''template <int BN, int BK, int CG, int S2>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // weight rows: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kARowBytes = S2 ? 128 : BK * 2;     // halo rows (S2: a pixel pair)
    static constexpr int kABytes = halo_a_bytes(BK, S2);
    static constexpr int kATxBytes = halo_rows(S2) * kARowBytes;
    static constexpr int kBBytes = (BN / CG) * kRowBytes;    // a CTA of a pair holds BN / 2 weight rows
};''')
rep('''template <int BN, int BK, int CG>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    using L = HaloSmem<BN, BK, CG>;''','''template <int BN, int BK, int CG, int S2>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    static_assert(!S2 || (BK == 32 && CG == 1), "the stride-2 pair form is written for 32 input channels");
    using L = HaloSmem<BN, BK, CG, S2>;
    constexpr int kARowBytes = L::kARowBytes;''')
rep('''                        if (rank == 0) ptx::mbar_expect_tx(&afull[sa], 2 * kHaloW * kHaloH * kRowBytes);
                        ptx::tma_load_4d_cg2(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK,
                                             tc.x0 - 1, tc.y0 - 1, tc.n0);
                    } else {
                        ptx::mbar_expect_tx(&afull[sa], kHaloW * kHaloH * kRowBytes);
                        ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, tc.x0 - 1,
                                         tc.y0 - 1, tc.n0);
                    }''','''                        if (rank == 0) ptx::mbar_expect_tx(&afull[sa], 2 * L::kATxBytes);
                        ptx::tma_load_4d_cg2(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK,
                                             tc.x0 - 1, tc.y0 - 1, tc.n0);
                    } else {
                        ptx::mbar_expect_tx(&afull[sa], L::kATxBytes);
                        if (S2)   // pair view: x in pairs (pair ox0 - 1 first), y in input rows (row 2*oy0 - 1 first)
                            ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], 0, tc.x0 - 1, 2 * tc.y0 - 1,
                                             tc.n0);
                        else
                            ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK,
                                             tc.x0 - 1, tc.y0 - 1, tc.n0);
                    }''')
rep('''            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kRowBytes>(ptx::smem_u32(sA), kHaloW * kRowBytes);''','''            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kARowBytes>(
                ptx::smem_u32(sA), S2 ? 2 * kS2HaloW * 128 : kHaloW * kARowBytes);''')
rep('''                        const uint32_t a_tap = a_lo + (((kh * kHaloW + kw) * kRowBytes) >> 4);''','''                        const uint32_t a_tap =
                            a_lo + (S2 ? (((kh * kS2HaloW + (kw != 0 ? 1 : 0)) * 128 + (kw != 1 ? 64 : 0)) >> 4)
                                       : (((kh * kHaloW + kw) * kARowBytes) >> 4));''')
# host
rep('''    const bool halo_shape = d.k == 3 && d.stride == 1 && (d.cin % 64 == 0 || d.cin == 32) && wo % 8 == 0 &&
                            ceil_div(ho, 16) * 16 * 4 <= ho * 5;''','''    static const int s2_env = getenv("WT_CONV_S2HALO") ? atoi(getenv("WT_CONV_S2HALO")) : 1;
    const bool tall_enough = ceil_div(ho, 16) * 16 * 4 <= ho * 5;
    // stride-2 pair form: 32 input channels that fill their buffer (a pixel pair is one contiguous 128-byte row)
    const bool halo_s2 = s2_env && d.k == 3 && d.stride == 2 && d.cin == 32 && d.src.ctot == 32 && d.src.coff == 0 &&
                         wo % 8 == 0 && tall_enough && d.cout % 32 == 0 && d.cout <= 64 && !d.res.base;
    const bool halo_shape = halo_s2 || (d.k == 3 && d.stride == 1 && (d.cin % 64 == 0 || d.cin == 32) && wo % 8 == 0 &&
                                        tall_enough);''')
rep('''    pl->cg = (cg_env == 2 && bn >= 128 && bk == 64) ? 2 : 1;''','''    pl->cg = (cg_env == 2 && bn >= 128 && bk == 64) ? 2 : 1;
    pl->s2 = 0;''')
rep('''    int cg;                    // CTAs per MMA (1, or 2 = cta_group::2 pairs launched as 2-CTA clusters)''','''    int cg;                    // CTAs per MMA (1, or 2 = cta_group::2 pairs launched as 2-CTA clusters)
    int s2;                    // halo kernel in its stride-2 pixel-pair form''')
rep('''    if (pl->halo) {
        bk = d.cin % 64 == 0 ? 64 : 32;
        pl->bk = bk;''','''    if (pl->halo) {
        bk = d.cin % 64 == 0 ? 64 : 32;
        pl->bk = bk;
        pl->s2 = halo_s2 ? 1 : 0;''')
rep('''        const int kHaloABytes = halo_a_bytes(bk);''','''        const int kHaloABytes = halo_a_bytes(bk, pl->s2);''')
rep('''    if (d.stride == 1) {
        // the channel extent ends with the slice, so a K block wider than the slice is zero-filled
        const uint64_t dims[4]''','''    if (pl->halo && pl->s2) {
        // pixel-pair view of the whole input: [64 = 2 px x 32 ch][w / 2 pairs][h][n]
        const uint64_t dims[4] = {64, uint64_t(d.src.w / 2), uint64_t(d.src.h), uint64_t(d.batch)};
        const uint64_t str[3] = {128, uint64_t(d.src.w) * 64, uint64_t(d.src.w) * 64 * d.src.h};
        const uint32_t box[4] = {64, kS2HaloW, kS2HaloH, 1};
        rc |= encode_tmap(&p.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.src.base, dims, str, box, 128);
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    } else if (d.stride == 1) {
        // the channel extent ends with the slice, so a K block wider than the slice is zero-filled
        const uint64_t dims[4]''')
# launch templates
rep('''template <int BN, int BK, int CG>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    return launch_kernel(conv_halo_kernel<BN, BK, CG>, &configured, prm, CG, smem, grid, stream);
}''','''template <int BN, int BK, int CG, int S2 = 0>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    return launch_kernel(conv_halo_kernel<BN, BK, CG, S2>, &configured, prm, CG, smem, grid, stream);
}''')
rep('''        if (pl->bk == 32) {   // 32 input channels (the 160x160 C2f bottlenecks)
            switch (pl->bn) {''','''        if (pl->s2) {         // 32 -> 32/64 channels, stride 2 (layer 1)
            switch (pl->bn) {
                case 64: return launch_halo<64, 32, 1, 1>(prm, smem, grid, stream);
                case 32: return launch_halo<32, 32, 1, 1>(prm, smem, grid, stream);
            }
            set_error("no stride-2 halo instantiation for this BN");
            return 1;
        }
        if (pl->bk == 32) {   // 32 input channels (the 160x160 C2f bottlenecks)
            switch (pl->bn) {''')
open(p,'w').write(s)
