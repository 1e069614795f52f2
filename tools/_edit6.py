p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new,cnt=None):
    global s
    assert old in s, old[:80]
    if cnt is not None: assert s.count(old)==cnt, (s.count(old), old[:60])
    s=s.replace(old,new)

rep('''constexpr int kThreads = 64 + kEpiGroups * kEpiThreads;''','''constexpr int kMmaWarps = 2;            // MMA issuer warps; issuer w owns accumulator w (tiles it % 2 == w)
constexpr int kFirstEpiWarp = 1 + kMmaWarps;
constexpr int kThreads = kFirstEpiWarp * 32 + kEpiGroups * kEpiThreads;''')
rep('''    const int g = (warp - 2) >> 2;          // epilogue group == accumulator buffer
    const int et = threadIdx.x - 64 - g * kEpiThreads;   // 0..127 inside the group''','''    const int g = (warp - kFirstEpiWarp) >> 2;          // epilogue group == accumulator buffer
    const int et = threadIdx.x - kFirstEpiWarp * 32 - g * kEpiThreads;   // 0..127 inside the group''')
rep('''// Warp roles (320 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer and
// TMEM owner, warps 2..5 / 6..9 = two epilogue groups (TMEM -> registers -> swizzled smem -> TMA
// store).  The TMEM accumulator is double-buffered and each epilogue group owns one buffer, so the
// epilogues of tiles i and i+1 overlap each other and the main loop of tile i+2.''','''// Warp roles (352 threads, persistent over tiles): warp 0 = TMA producer, warps 1..2 = MMA issuers
// (warp 1 also owns the TMEM allocation), warps 3..6 / 7..10 = two epilogue groups (TMEM -> registers
// -> swizzled smem -> TMA store).  The TMEM accumulator is double-buffered; MMA issuer w and epilogue
// group w own buffer w, i.e. every second tile of the CTA.  Two issuers because ONE thread cannot issue
// narrow MMAs fast enough (measured: ~9 uniform-datapath instructions at ~8-10 clk each per UMMA,
// while a 128x64x16 UMMA occupies the tensor pipe for only 32 clk).''')

# ---- generic kernel MMA role
i0=s.index('    } else if (warp == 1) {\n        // ------------------------------------------------------------------ MMA issuer')
i1=s.index('    } else {\n        // ------------------------------------------------------------------ epilogue (warps 2..5)')
new='''    } else if (warp < kFirstEpiWarp) {
        // ------------------------------------------------------------------ MMA issuers (warps 1, 2)
        // ONE elected lane per issuer warp runs the whole loop (waits included).  elect.sync tells the
        // compiler that a single thread is active, so descriptors and barrier addresses stay on the
        // uniform datapath; the descriptor low words advance by plain 32-bit adds (stage, K slice).
        // Issuer w handles tiles it = w, w + 2, ...; both walk the same smem ring, whose slot for the
        // g-th K block of the CTA is g % stages.
        const int w = warp - 1;
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));
            const uint64_t b_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            const uint32_t d_tmem = tmem_base + w * BN;
            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);
                int stage = int(g0 % uint32_t(kStages));
                uint32_t phase = (g0 / uint32_t(kStages)) & 1u;
                uint32_t a_lo = a_lo0 + stage * (L::kABytes >> 4), b_lo = b_lo0 + stage * (L::kBBytes >> 4);
                ptx::mbar_wait(&tempty_bar[w], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        // advance 16 bf16 = 32 B along K inside the swizzle span: start address field += 2
                        ptx::umma_bf16_lohi(d_tmem, a_lo + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc, (kb | kk) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs finish
                    a_lo += L::kABytes >> 4;
                    b_lo += L::kBBytes >> 4;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                        a_lo = a_lo0;
                        b_lo = b_lo0;
                    }
                }
                ptx::umma_commit(&tfull_bar[w]);   // accumulator complete
            }
        }
        __syncwarp();
'''
s=s[:i0]+new+s[i1:]

# ---- halo kernel MMA role
i0=s.index('    } else if (warp == 1) {\n        // MMA issuer: one elected lane runs the whole loop')
i1=s.index('    } else {\n        conv_epilogue<BN>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);\n    }\n\n    ptx::tc_fence_before();\n    __syncthreads();\n    if (warp == 1) {\n        ptx::tc_fence_after();\n        ptx::tmem_dealloc(tmem_base, kTmemCols);\n    }\n}\n\n}  // namespace')
new='''    } else if (warp < kFirstEpiWarp) {
        // MMA issuers (warps 1, 2): one elected lane each runs the whole loop for tiles it = w, w + 2, ...;
        // descriptor low words advance by 32-bit adds (halo stage, weight stage, tap offset and K slice are all
        // additive in the start-address field).  Ring slots: halo tile g -> g % a_stages, weight tile g -> g % b_stages.
        const int w = warp - 1;
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo(ptx::smem_u32(sA), kHaloW * 128);
            const uint64_t b_desc0 = ptx::make_kmajor_desc<128>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            const uint32_t d_tmem = tmem_base + w * BN;
            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
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
                uint32_t a_lo = a_lo0 + sa * (kHaloABytes >> 4), b_lo = b_lo0 + sb * (L::kBBytes >> 4);
                // resident weights: both issuers wait once for all nine taps (phase 0 of each slot)
                const bool wait_b = !p.resident || it < 2;
                ptx::mbar_wait(&tempty_bar[w], ((it >> 1) & 1) ^ 1);
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    ptx::mbar_wait(&afull[sa], pa);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (wait_b) {
                            ptx::mbar_wait(&bfull[sb], pb);
                            ptx::tc_fence_after();
                        }
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t a_tap = a_lo + (((kh * kHaloW + kw) * 128) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            ptx::umma_bf16_lohi(d_tmem, a_tap + 2 * kk, a_hi, b_lo + 2 * kk, b_hi, idesc,
                                                (cb | tap | kk) != 0);
                        if (!p.resident) ptx::umma_commit(&bempty[sb]);
                        b_lo += L::kBBytes >> 4;
                        if (++sb == kBStages) { sb = 0; pb ^= 1; b_lo = b_lo0; }
                    }
                    ptx::umma_commit(&aempty[sa]);
                    a_lo += kHaloABytes >> 4;
                    if (++sa == kAStages) { sa = 0; pa ^= 1; a_lo = a_lo0; }
                }
                ptx::umma_commit(&tfull_bar[w]);
            }
        }
        __syncwarp();
'''
s=s[:i0]+new+s[i1:]
open(p,'w').write(s)
