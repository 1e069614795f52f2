p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new,cnt=None):
    global s
    c=s.count(old)
    assert c>=1, old[:90]
    if cnt is not None: assert c==cnt,(c,old[:90])
    s=s.replace(old,new)

# ---------------- params
rep('''    int tiles_x, tiles_y, tiles_n;   // pixel-tile grid''','''    int tiles_x, tiles_y, tiles_n;   // pixel-tile grid (CTA pairs: tiles_x counts PAIRS of x-adjacent tiles)''')

# ---------------- tile decode helper + epilogue
rep('''// Epilogue: two groups of 4 warps;''','''// Work item t -> (N block, pixel patch origin).  CG == 2: a work item is a PAIR of x-adjacent patches, CTA
// `rank` of the pair takes patch 2 * xb + rank (a patch beyond the map is all TMA zero-fill / clipped stores).
struct TileCoord {
    int nblk, x0, y0, n0;
};
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, int t, int rank) {
    TileCoord c;
    c.nblk = t % p.n_blocks;
    int m = t / p.n_blocks;
    int xb = m % p.tiles_x;
    m /= p.tiles_x;
    const int yb = m % p.tiles_y;
    const int nb = m / p.tiles_y;
    if (CG == 2) xb = 2 * xb + rank;
    c.x0 = xb * p.tw;
    c.y0 = yb * p.th;
    c.n0 = nb * p.tn;
    return c;
}

// Epilogue: two groups of 4 warps;''')
rep('''template <int BN>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_bar_all,
                                              uint32_t tmem_base, int warp, int lane) {''','''template <int BN, int CG>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, uint8_t* sStageAll, const float* sBias,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_bar_all,
                                              uint32_t tmem_base, int warp, int lane, int rank) {''')
rep('''    for (int tile = blockIdx.x + g * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
        const int nblk = tile % p.n_blocks;
        int m = tile / p.n_blocks;
        const int xb = m % p.tiles_x;
        m /= p.tiles_x;
        const int yb = m % p.tiles_y;
        const int nb = m / p.tiles_y;
        const int x0 = xb * p.tw, y0 = yb * p.th, n0 = nb * p.tn;
        const uint32_t aphase = (it >> 1) & 1;''','''    const int first = blockIdx.x / CG, step = gridDim.x / CG;
    for (int tile = first + g * step; tile < p.num_tiles; tile += 2 * step, it += 2) {
        const TileCoord tc = decode_tile<CG>(p, tile, rank);
        const int nblk = tc.nblk, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
        const uint32_t aphase = (it >> 1) & 1;''')
# tempty arrive (two places: dot path and normal path)
rep('''                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[g]);''','''                    if (lane == 0) {
                        if (CG == 2) ptx::mbar_arrive_leader(&tempty_bar[g]);   // the leader's MMA thread waits for both CTAs
                        else ptx::mbar_arrive(&tempty_bar[g]);
                    }''',1)
rep('''                if (lane == 0) ptx::mbar_arrive(&tempty_bar[g]);''','''                if (lane == 0) {
                    if (CG == 2) ptx::mbar_arrive_leader(&tempty_bar[g]);
                    else ptx::mbar_arrive(&tempty_bar[g]);
                }''',1)

# ---------------- generic kernel
rep('''template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    using L = SmemLayout<BN, BK>;''','''template <int BN, int BK, int CG>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    using L = SmemLayout<BN / CG, BK>;   // a CTA of a pair holds half of the B tile (BN / 2 weight rows)
    const int rank = CG == 2 ? int(ptx::cluster_ctarank()) : 0;
    const int first = blockIdx.x / CG, step = gridDim.x / CG;''')
s=s.replace('''        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 4);
        }''','''        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 4 * CG);   // 4 epilogue warps per CTA of the pair
        }''')
assert s.count('4 * CG);   // 4 epilogue warps')==2
s=s.replace('''    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }''','''    if (warp == 1) {
        if (CG == 2) {
            ptx::tmem_alloc_cg2(tmem_slot, kTmemCols);
            ptx::tmem_relinquish_cg2();
        } else {
            ptx::tmem_alloc(tmem_slot, kTmemCols);
            ptx::tmem_relinquish();
        }
    }''')
s=s.replace('''    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;''','''    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();   // the peer's barriers are initialised before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;''')
assert s.count('if (CG == 2) ptx::cluster_sync();   // the peer')==2
s=s.replace('''    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}''','''    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();   // nobody signals a peer barrier or reads peer smem/TMEM after this
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
        else ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}''')
assert s.count('ptx::tmem_dealloc_cg2(tmem_base, kTmemCols)')==2

# generic producer
rep('''        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nblk = tile % p.n_blocks;
            int m = tile / p.n_blocks;
            const int xb = m % p.tiles_x;
            m /= p.tiles_x;
            const int yb = m % p.tiles_y;
            const int nb = m / p.tiles_y;
            const int x0 = xb * p.tw, y0 = yb * p.th, n0 = nb * p.tn;
            for (int tap = 0; tap < taps; ++tap) {''','''        for (int tile = first; tile < p.num_tiles; tile += step) {
            const TileCoord tc = decode_tile<CG>(p, tile, rank);
            const int nblk = tc.nblk, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
            for (int tap = 0; tap < taps; ++tap) {''')
rep('''                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                        ptx::tma_load_4d(sA + stage * L::kABytes, &p.tmA[mapi], &full_bar[stage],
                                         p.src_coff + cb * BK, ax, ay, n0);
                        ptx::tma_load_2d(sB + stage * L::kBBytes, &p.tmB, &full_bar[stage], tap * p.cin + cb * BK,
                                         nblk * BN);
                    }''','''                    if (ptx::elect_one()) {
                        if (CG == 2) {
                            // both CTAs' bytes are counted on the leader's barrier; each CTA loads its own pixel
                            // patch and its half of the weight rows
                            if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                            ptx::tma_load_4d_cg2(sA + stage * L::kABytes, &p.tmA[mapi], &full_bar[stage],
                                                 p.src_coff + cb * BK, ax, ay, n0);
                            ptx::tma_load_2d_cg2(sB + stage * L::kBBytes, &p.tmB, &full_bar[stage],
                                                 tap * p.cin + cb * BK, nblk * BN + rank * (BN / 2));
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                            ptx::tma_load_4d(sA + stage * L::kABytes, &p.tmA[mapi], &full_bar[stage],
                                             p.src_coff + cb * BK, ax, ay, n0);
                            ptx::tma_load_2d(sB + stage * L::kBBytes, &p.tmB, &full_bar[stage], tap * p.cin + cb * BK,
                                             nblk * BN);
                        }
                    }''')
# generic MMA
rep('''        const int w = warp - 1;
        if (w < p.issuers && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));''','''        const int w = warp - 1;
        if (w < p.issuers && rank == 0 && ptx::elect_one()) {   // CTA pairs: only the leader issues
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM * CG, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));''')
rep('''            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles;
                 tile += p.issuers * gridDim.x, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);''','''            for (int tile = first + w * step; tile < p.num_tiles; tile += p.issuers * step, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);''')
rep('''                        ptx::umma_bf16_lohi(d_tmem, a_lo + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc, (kb | kk) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs finish''','''                        if (CG == 2)
                            ptx::umma_bf16_lohi_cg2(d_tmem, a_lo + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc,
                                                    (kb | kk) != 0);
                        else
                            ptx::umma_bf16_lohi(d_tmem, a_lo + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc, (kb | kk) != 0);
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs finish
                    if (CG == 2) ptx::umma_commit_cg2(&empty_bar[stage]);
                    else ptx::umma_commit(&empty_bar[stage]);''')
rep('''                ptx::umma_commit(&tfull_bar[ab]);   // accumulator complete''','''                if (CG == 2) ptx::umma_commit_cg2(&tfull_bar[ab]);   // accumulator complete (both CTAs' epilogues)
                else ptx::umma_commit(&tfull_bar[ab]);''')
rep('''        conv_epilogue<BN>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();   // nobody signals a peer barrier or reads peer smem/TMEM after this
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
        else ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 / stride-1 variant''','''        conv_epilogue<BN, CG>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane, rank);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();   // nobody signals a peer barrier or reads peer smem/TMEM after this
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
        else ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 / stride-1 variant''')

# ---------------- halo kernel
rep('''template <int BN, int BK>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kABytes = halo_a_bytes(BK);
    static constexpr int kBBytes = BN * kRowBytes;
};''','''template <int BN, int BK, int CG>
struct HaloSmem {
    static constexpr int kRowBytes = BK * 2;                 // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    static constexpr int kABytes = halo_a_bytes(BK);
    static constexpr int kBBytes = (BN / CG) * kRowBytes;    // a CTA of a pair holds BN / 2 weight rows
};''')
rep('''template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    using L = HaloSmem<BN, BK>;''','''template <int BN, int BK, int CG>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
    using L = HaloSmem<BN, BK, CG>;
    const int rank = CG == 2 ? int(ptx::cluster_ctarank()) : 0;
    const int first = blockIdx.x / CG, step = gridDim.x / CG;''')
rep('''        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nblk = tile % p.n_blocks;
            int m = tile / p.n_blocks;
            const int xb = m % p.tiles_x;
            m /= p.tiles_x;
            const int yb = m % p.tiles_y;
            const int nb = m / p.tiles_y;
            for (int cb = 0; cb < p.cin_blocks; ++cb) {
                ptx::mbar_wait(&aempty[sa], pa ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&afull[sa], kHaloW * kHaloH * kRowBytes);
                    ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, xb * 8 - 1,
                                     yb * 16 - 1, nb);
                }''','''        for (int tile = first; tile < p.num_tiles; tile += step) {
            const TileCoord tc = decode_tile<CG>(p, tile, rank);
            const int nblk = tc.nblk;
            for (int cb = 0; cb < p.cin_blocks; ++cb) {
                ptx::mbar_wait(&aempty[sa], pa ^ 1);
                if (ptx::elect_one()) {
                    if (CG == 2) {
                        if (rank == 0) ptx::mbar_expect_tx(&afull[sa], 2 * kHaloW * kHaloH * kRowBytes);
                        ptx::tma_load_4d_cg2(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK,
                                             tc.x0 - 1, tc.y0 - 1, tc.n0);
                    } else {
                        ptx::mbar_expect_tx(&afull[sa], kHaloW * kHaloH * kRowBytes);
                        ptx::tma_load_4d(sA + sa * kHaloABytes, &p.tmA[0], &afull[sa], p.src_coff + cb * BK, tc.x0 - 1,
                                         tc.y0 - 1, tc.n0);
                    }
                }''')
rep('''                if (p.resident && tile != blockIdx.x) continue;   // weights already in shared memory''','''                if (p.resident && tile != first) continue;   // weights already in shared memory''')
rep('''                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&bfull[sb], L::kBBytes);
                        ptx::tma_load_3d(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * BK, tap, nblk * BN);
                    }''','''                    if (ptx::elect_one()) {
                        if (CG == 2) {
                            if (rank == 0) ptx::mbar_expect_tx(&bfull[sb], 2 * L::kBBytes);
                            ptx::tma_load_3d_cg2(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * BK, tap,
                                                 nblk * BN + rank * (BN / 2));
                        } else {
                            ptx::mbar_expect_tx(&bfull[sb], L::kBBytes);
                            ptx::tma_load_3d(sB + sb * L::kBBytes, &p.tmB, &bfull[sb], cb * BK, tap, nblk * BN);
                        }
                    }''')
rep('''        const int w = warp - 1;
        if (w < p.issuers && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kRowBytes>''','''        const int w = warp - 1;
        if (w < p.issuers && rank == 0 && ptx::elect_one()) {   // CTA pairs: only the leader issues
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM * CG, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo<kRowBytes>''')
rep('''            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles;
                 tile += p.issuers * gridDim.x, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);''','''            for (int tile = first + w * step; tile < p.num_tiles; tile += p.issuers * step, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);''')
rep('''                        for (int kk = 0; kk < BK / 16; ++kk)
                            ptx::umma_bf16_lohi(d_tmem, a_tap + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc,
                                                (cb | tap | kk) != 0);
                        if (!p.resident) ptx::umma_commit(&bempty[sb]);''','''                        for (int kk = 0; kk < BK / 16; ++kk) {
                            if (CG == 2)
                                ptx::umma_bf16_lohi_cg2(d_tmem, a_tap + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc,
                                                        (cb | tap | kk) != 0);
                            else
                                ptx::umma_bf16_lohi(d_tmem, a_tap + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc,
                                                    (cb | tap | kk) != 0);
                        }
                        if (!p.resident) {
                            if (CG == 2) ptx::umma_commit_cg2(&bempty[sb]);
                            else ptx::umma_commit(&bempty[sb]);
                        }''')
rep('''                    ptx::umma_commit(&aempty[sa]);
                    a_lo += kHaloABytes >> 4;''','''                    if (CG == 2) ptx::umma_commit_cg2(&aempty[sa]);
                    else ptx::umma_commit(&aempty[sa]);
                    a_lo += kHaloABytes >> 4;''')
rep('''                ptx::umma_commit(&tfull_bar[ab]);
            }
        }
        __syncwarp();''','''                if (CG == 2) ptx::umma_commit_cg2(&tfull_bar[ab]);
                else ptx::umma_commit(&tfull_bar[ab]);
            }
        }
        __syncwarp();''')
rep('''        conv_epilogue<BN>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);''','''        conv_epilogue<BN, CG>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane, rank);''')
rep('''                const bool wait_b = !p.resident || it < p.issuers;''','''                const bool wait_b = !p.resident || it < p.issuers;   // (resident: the first tile of each issuer)''')
open(p,'w').write(s)
