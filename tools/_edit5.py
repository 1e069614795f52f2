def edit(path, pairs):
    s=open(path).read()
    for old,new in pairs:
        assert old in s, (path, old[:80])
        s=s.replace(old,new)
    open(path,'w').write(s)

# ptx: umma with (lo, hi) descriptor halves
edit('wtracker_b200/csrc/ptx.cuh', [
('''// arrive on an mbarrier when all previously issued MMAs of this thread have completed''','''// Same with the descriptors given as (low word, high word): the high word (strides, version, swizzle)
// is a per-kernel constant and the low word is "start address >> 4 | LBO", so stepping through taps /
// K slices / pipeline stages is ONE 32-bit add per operand instead of rebuilding a 64-bit descriptor.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\\n\\t"
        ".reg .pred p;\\n\\t"
        ".reg .b64 da, db;\\n\\t"
        "mov.b64 da, {%1, %2};\\n\\t"
        "mov.b64 db, {%3, %4};\\n\\t"
        "setp.ne.b32 p, %6, 0;\\n\\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\\n\\t"
        "}\\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed'''),
])

p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
# ---- generic kernel MMA
i0=s.index('        // ------------------------------------------------------------------ MMA issuer\n        // Warp-uniform loop; one elected lane issues the UMMAs and the commits.')
i1=s.index('    } else {\n        // ------------------------------------------------------------------ epilogue (warps 2..5)')
new='''        // ------------------------------------------------------------------ MMA issuer
        // ONE elected lane runs the whole loop (waits included).  elect.sync tells the compiler that a
        // single thread is active, so descriptors and barrier addresses stay on the uniform datapath; the
        // descriptor low words advance by plain 32-bit adds (stage, K slice).
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));
            const uint64_t b_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t a_lo = a_lo0, b_lo = b_lo0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[ab], aphase ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + ab * BN;
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
                ptx::umma_commit(&tfull_bar[ab]);   // accumulator complete
            }
        }
        __syncwarp();
'''
s=s[:i0]+new+s[i1:]

# ---- halo kernel MMA
i0=s.index('        // MMA issuer: warp-uniform loop, one elected lane issues\n')
i1=s.index('    } else {\n        conv_epilogue<BN>(p, sStage, sBias, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);\n    }\n\n    ptx::tc_fence_before();\n    __syncthreads();\n    if (warp == 1) {\n        ptx::tc_fence_after();\n        ptx::tmem_dealloc(tmem_base, kTmemCols);\n    }\n}\n\n}  // namespace')
new='''        // MMA issuer: one elected lane runs the whole loop; descriptor low words advance by 32-bit adds
        // (halo stage, weight stage, tap offset and K slice are all additive in the start-address field)
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo(ptx::smem_u32(sA), kHaloW * 128);
            const uint64_t b_desc0 = ptx::make_kmajor_desc<128>(ptx::smem_u32(sB));
            const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
            const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            uint32_t a_lo = a_lo0, b_lo = b_lo0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[ab], aphase ^ 1);
                const uint32_t d_tmem = tmem_base + ab * BN;
                const bool wait_b = !p.resident || tile == blockIdx.x;
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
                ptx::umma_commit(&tfull_bar[ab]);
            }
        }
        __syncwarp();
'''
s=s[:i0]+new+s[i1:]
open(p,'w').write(s)
