p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new,cnt=None):
    global s
    assert old in s, old[:80]
    s=s.replace(old,new)

rep('''    int resident;''','''    int resident;
    // MMA issuer warps in use (1 | 2).  Two issuers take alternate tiles; that is only safe when every
    // ring slot has ONE consumer (an mbarrier parity wait cannot tell phase k from phase k + 2), i.e. in
    // the resident-weight halo kernel with one halo tile per output tile and an even number of halo stages.
    int issuers;''')
# generic kernel
rep('''        const int w = warp - 1;
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));''','''        const int w = warp - 1;
        if (w < p.issuers && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc<kRowBytes>(ptx::smem_u32(sA));''')
rep('''            const uint32_t d_tmem = tmem_base + w * BN;
            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);''','''            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles;
                 tile += p.issuers * gridDim.x, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t g0 = uint32_t(it) * uint32_t(num_kb);''')
rep('''                ptx::mbar_wait(&tempty_bar[w], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {''','''                ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {''')
rep('''                ptx::umma_commit(&tfull_bar[w]);   // accumulator complete''','''                ptx::umma_commit(&tfull_bar[ab]);   // accumulator complete''')
# halo kernel
rep('''        const int w = warp - 1;
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo(ptx::smem_u32(sA), kHaloW * 128);''','''        const int w = warp - 1;
        if (w < p.issuers && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN);
            const uint64_t a_desc0 = ptx::make_kmajor_desc_sbo(ptx::smem_u32(sA), kHaloW * 128);''')
rep('''            const uint32_t d_tmem = tmem_base + w * BN;
            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
                const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);''','''            int it = w;
            for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles;
                 tile += p.issuers * gridDim.x, it += p.issuers) {
                const int ab = it & 1;
                const uint32_t d_tmem = tmem_base + ab * BN;
                const uint32_t ga = uint32_t(it) * uint32_t(p.cin_blocks);''')
rep('''                const bool wait_b = !p.resident || it < 2;
                ptx::mbar_wait(&tempty_bar[w], ((it >> 1) & 1) ^ 1);''','''                const bool wait_b = !p.resident || it < p.issuers;
                ptx::mbar_wait(&tempty_bar[ab], ((it >> 1) & 1) ^ 1);''')
rep('''                ptx::umma_commit(&tfull_bar[w]);
            }''','''                ptx::umma_commit(&tfull_bar[ab]);
            }''')
# host
rep('''        if (p.resident) {
            p.stages = 9;   // the ring wraps once per tile: stage index == tap
            const int spare = (kSmemBudget - fixed - 9 * b_bytes) / kHaloABytes;
            p.a_stages = spare > kMaxAStages ? kMaxAStages : spare;
        }''','''        p.issuers = 1;
        if (p.resident) {
            p.stages = 9;   // the ring wraps once per tile: stage index == tap
            const int spare = (kSmemBudget - fixed - 9 * b_bytes) / kHaloABytes;
            p.a_stages = spare >= 4 ? 4 : (spare >= 2 ? 2 : spare);
            static const int issuers_env = getenv("WT_CONV_ISSUERS") ? atoi(getenv("WT_CONV_ISSUERS")) : 2;
            if (p.a_stages % 2 == 0 && issuers_env == 2) p.issuers = 2;   // slot s is only ever read by issuer s % 2
        }''')
rep('''        p.a_stages = 0;
        p.resident = 0;''','''        p.a_stages = 0;
        p.resident = 0;
        p.issuers = 1;''')
open(p,'w').write(s)
