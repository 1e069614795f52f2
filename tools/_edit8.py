p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new):
    global s
    assert old in s, old[:80]
    s=s.replace(old,new)

rep('''constexpr int kHaloABytes = ((kHaloW * kHaloH * 128 + 1023) / 1024) * 1024;   // 23552
''','''// bytes of one halo stage for a K block of BK channels (rows of 2 * BK bytes), 1024-byte aligned
__host__ __device__ constexpr int halo_a_bytes(int bk) { return ((kHaloW * kHaloH * bk * 2 + 1023) / 1024) * 1024; }
''')
rep('''template <int BN>
struct HaloSmem {
    static constexpr int kBBytes = BN * 128;
};''','''template <int BN, int BK>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kABytes = halo_a_bytes(BK);
    static constexpr int kBBytes = BN * kRowBytes;
};''')
rep('''template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    using L = HaloSmem<BN>;''','''template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    using L = HaloSmem<BN, BK>;
    constexpr int kHaloABytes = L::kABytes;
    constexpr int kRowBytes = L::kRowBytes;''')
rep('''                    ptx::mbar_expect_tx(&afull[sa], kHaloW * kHaloH * 128);
                    ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * 64, xb * 8 - 1,''','''                    ptx::mbar_expect_tx(&afull[sa], kHaloW * kHaloH * kRowBytes);
                    ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, xb * 8 - 1,''')
rep('''                        ptx::tma_load_3d(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * 64, tap, nblk * BN);''','''                        ptx::tma_load_3d(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * BK, tap, nblk * BN);''')
rep('''            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo(ptx::smem_u32(sA), kHaloW * 128);
            const uint64_t b_desc0 = ptx::make_kmajor_desc<128>(ptx::smem_u32(sB));''','''            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kRowBytes>(ptx::smem_u32(sA), kHaloW * kRowBytes);
            const uint64_t b_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sB));''')
rep('''                        const uint32_t a_tap = a_lo + (((kh * kHaloW + kw) * 128) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)''','''                        const uint32_t a_tap = a_lo + (((kh * kHaloW + kw) * kRowBytes) >> 4);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)''')
# comment of halo kernel
rep('''// is needed), so TMA (writer) and UMMA (reader) agree on it for any start pixel.  Layers with 32
// input channels use the same 64-wide K block: the tensor maps end at the slice's last channel, so
// TMA zero-fills the upper half of both operands.  Weights stream through their own, deeper pipeline (one BN x 64 tile per tap).''','''// is needed), so TMA (writer) and UMMA (reader) agree on it for any start pixel.  Layers with 32
// input channels use a 32-wide K block: 64-byte rows and SWIZZLE_64B for both operands (half the
// shared-memory traffic of zero-padding K to 64 — these layers are shared-memory-bandwidth bound).
// Weights stream through their own, deeper pipeline (one BN x BK tile per tap) or stay resident.''')
# host
rep('''    if (pl->halo) {
        bk = 64;
        pl->bk = 64;
        p.tw = 8;''','''    if (pl->halo) {
        bk = d.cin % 64 == 0 ? 64 : 32;
        pl->bk = bk;
        p.tw = 8;''')
rep('''        const uint64_t dims[3] = {uint64_t(d.cin), 9, uint64_t(d.cout)};
        const uint64_t str[2] = {uint64_t(d.cin) * 2, uint64_t(d.cin) * 2 * 9};
        const uint32_t box[3] = {64, 1, uint32_t(bn)};
        rc |= encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), dims, str, box,
                          128);''','''        const uint64_t dims[3] = {uint64_t(d.cin), 9, uint64_t(d.cout)};
        const uint64_t str[2] = {uint64_t(d.cin) * 2, uint64_t(d.cin) * 2 * 9};
        const uint32_t box[3] = {uint32_t(bk), 1, uint32_t(bn)};
        rc |= encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), dims, str, box,
                          sw_in);''')
rep('''        p.a_stages = bn == 256 ? 2 : 3;
        const int b_bytes = bn * 128;
        p.stages = (kSmemBudget - fixed - p.a_stages * kHaloABytes) / b_bytes;''','''        p.a_stages = bn == 256 ? 2 : 3;
        const int b_bytes = bn * bk * 2;
        const int kHaloABytes = halo_a_bytes(bk);
        p.stages = (kSmemBudget - fixed - p.a_stages * kHaloABytes) / b_bytes;''')
rep('''            if (p.a_stages % 2 == 0 && issuers_env == 2) p.issuers = 2;''','''            if (p.a_stages % 2 == 0 && issuers_env == 2) p.issuers = 2;   // measured slower than one issuer: off by default''')
rep('''getenv("WT_CONV_ISSUERS") ? atoi(getenv("WT_CONV_ISSUERS")) : 2;''','''getenv("WT_CONV_ISSUERS") ? atoi(getenv("WT_CONV_ISSUERS")) : 1;''')
rep('''template <int BN>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBudget));
        configured = true;
    }
    conv_halo_kernel<BN><<<grid, kThreads, smem, stream>>>(prm);''','''template <int BN, int BK>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBudget));
        configured = true;
    }
    conv_halo_kernel<BN, BK><<<grid, kThreads, smem, stream>>>(prm);''')
rep('''    if (pl->halo) {
        switch (pl->bn) {
            case 256: return launch_halo<256>(prm, pl->smem_bytes, grid, stream);
            case 128: return launch_halo<128>(prm, pl->smem_bytes, grid, stream);
            case 64:  return launch_halo<64>(prm, pl->smem_bytes, grid, stream);
            case 32:  return launch_halo<32>(prm, pl->smem_bytes, grid, stream);
        }
    }''','''    if (pl->halo) {
        if (pl->bk == 32) {   // 32 input channels (the 160x160 C2f bottlenecks)
            switch (pl->bn) {
                case 64: return launch_halo<64, 32>(prm, pl->smem_bytes, grid, stream);
                case 32: return launch_halo<32, 32>(prm, pl->smem_bytes, grid, stream);
            }
            set_error("no halo instantiation for this (BN, 32)");
            return 1;
        }
        switch (pl->bn) {
            case 256: return launch_halo<256, 64>(prm, pl->smem_bytes, grid, stream);
            case 128: return launch_halo<128, 64>(prm, pl->smem_bytes, grid, stream);
            case 64:  return launch_halo<64, 64>(prm, pl->smem_bytes, grid, stream);
            case 32:  return launch_halo<32, 64>(prm, pl->smem_bytes, grid, stream);
        }
    }''')
open(p,'w').write(s)

p='wtracker_b200/csrc/ptx.cuh'
s=open(p).read()
rep('''// Same, 128-byte rows / SWIZZLE_128B, with an explicit stride between 8-row groups (halo tiles).
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}''','''// Same with an explicit stride between 8-row groups (halo tiles).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "swizzle span");
    constexpr uint64_t layout = (ROW_BYTES == 128) ? 2ull : 4ull;
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= layout << 61;
    return d;
}''')
open(p,'w').write(s)
