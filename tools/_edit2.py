p='wtracker_b200/csrc/conv_tcgen05.cu'
s=open(p).read()
def rep(old,new):
    global s
    assert old in s, old[:70]
    s=s.replace(old,new)

rep('''constexpr int kEpiSmemBytes = kEpiGroups * 2 * kStageBufBytes;   // two staging buffers per group
''','''constexpr int kSmemBudget = 232448 - 1024;   // 227 KB minus alignment slack
constexpr int kBarrierBytes = 1024;
''')
rep('''    int act, has_res, out_f32;
    int num_tiles;''','''    int act, has_res, out_f32;
    int num_tiles;
    // shared-memory plan (host-chosen): pipeline depth and epilogue staging buffers per group (1 | 2).
    // HBM-bound layers (1x1, narrow N) want two staging buffers per epilogue group, MMA-bound layers
    // want the bytes as pipeline stages instead.
    int stages;        // generic kernel: A+B stages; halo kernel: weight (B) stages
    int a_stages;      // halo kernel: halo-tile stages
    int epi_bufs;''')
# SmemLayout: keep byte sizes only
rep('''    static constexpr int kFixedBytes = kEpiSmemBytes + kMaxCout * 4 + 1024;  // staging + bias + barriers
    static constexpr int kBudget = 232448 - 1024;                          // 227 KB minus alignment slack
    static constexpr int kStagesRaw = (kBudget - kFixedBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
    static constexpr int kTotalBytes = kStages * kStageBytes + kFixedBytes + 1024;
    static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};''','''};

// bytes after the pipeline stages: epilogue staging + bias vector + barriers
__host__ __device__ constexpr int fixed_smem_bytes(int epi_bufs) {
    return kEpiGroups * epi_bufs * kStageBufBytes + kMaxCout * 4 + kBarrierBytes;
}''')
# epilogue signature: add epi_bufs handling
rep('''    uint8_t* sStage = sStageAll + g * 2 * kStageBufBytes;
    uint64_t* res_bar = res_bar_all + 2 * g;''','''    const bool two_bufs = p.epi_bufs == 2;
    uint8_t* sStage = sStageAll + g * p.epi_bufs * kStageBufBytes;
    uint64_t* res_bar = res_bar_all + 2 * g;''')
rep('''            const int sb = unit_counter & 1;
            uint8_t* stage_buf = sStage + sb * kStageBufBytes;''','''            const int sb = two_bufs ? (unit_counter & 1) : 0;
            uint8_t* stage_buf = sStage + sb * kStageBufBytes;''')
rep('''                if (store_thread) ptx::tma_store_wait_read<1>();
                ptx::bar_sync(bar_id, kEpiThreads);''','''                if (store_thread) {
                    if (two_bufs) ptx::tma_store_wait_read<1>();
                    else ptx::tma_store_wait_read<0>();
                }
                ptx::bar_sync(bar_id, kEpiThreads);''')
rep('''            if (p.has_res && sub_in_unit == 0) ptx::mbar_wait(&res_bar[sb], (unit_counter >> 1) & 1);''','''            if (p.has_res && sub_in_unit == 0)
                ptx::mbar_wait(&res_bar[sb], (two_bufs ? (unit_counter >> 1) : unit_counter) & 1);''')
# generic kernel
rep('''    constexpr int kStages = L::kStages;
''','''    const int kStages = p.stages;
''')
rep('''    uint8_t* sStage = smem + kStages * L::kStageBytes;   // 2 groups x 2 x 16 KB epilogue staging
    float* sBias = reinterpret_cast<float*>(sStage + kEpiSmemBytes);   // [kMaxCout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);
    uint64_t* full_bar = bars;                       // [stages]  TMA -> MMA
    uint64_t* empty_bar = bars + kStages;            // [stages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kStages;        // [2]       MMA -> epilogue group
    uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]       epilogue group -> MMA
    uint64_t* res_bar = bars + 2 * kStages + 4;      // [2][2]    residual TMA -> epilogue group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 8);
''','''    uint8_t* sStage = smem + kStages * L::kStageBytes;   // 2 groups x epi_bufs x 16 KB epilogue staging
    float* sBias = reinterpret_cast<float*>(sStage + kEpiGroups * p.epi_bufs * kStageBufBytes);   // [kMaxCout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);
    uint64_t* full_bar = bars;                          // [stages]  TMA -> MMA
    uint64_t* empty_bar = bars + kMaxStages;            // [stages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [2]       MMA -> epilogue group
    uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;   // [2]       epilogue group -> MMA
    uint64_t* res_bar = bars + 2 * kMaxStages + 4;      // [2][2]    residual TMA -> epilogue group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 8);
''')
# halo
rep('''    static constexpr int kAStages = BN == 256 ? 2 : 3;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kFixedBytes = kEpiSmemBytes + kMaxCout * 4 + 1024;
    static constexpr int kBudget = 232448 - 1024;
    static constexpr int kBStagesRaw = (kBudget - kFixedBytes - kAStages * kHaloABytes) / kBBytes;
    static constexpr int kBStages = kBStagesRaw > 12 ? 12 : kBStagesRaw;
    static constexpr int kTotalBytes = kAStages * kHaloABytes + kBStages * kBBytes + kFixedBytes + 1024;
    static_assert(kBStages >= 3, "not enough shared memory for the weight pipeline");
};''','''    static constexpr int kBBytes = BN * 128;
};
constexpr int kMaxAStages = 4;''')
rep('''    constexpr int kAStages = L::kAStages, kBStages = L::kBStages;
''','''    const int kAStages = p.a_stages, kBStages = p.stages;
''')
rep('''    float* sBias = reinterpret_cast<float*>(sStage + kEpiSmemBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);''','''    float* sBias = reinterpret_cast<float*>(sStage + kEpiGroups * p.epi_bufs * kStageBufBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxCout);''')
rep('''    uint64_t* aempty = afull + kAStages;
    uint64_t* bfull = aempty + kAStages;
    uint64_t* bempty = bfull + kBStages;
    uint64_t* tfull_bar = bempty + kBStages;''','''    uint64_t* aempty = afull + kMaxAStages;
    uint64_t* bfull = aempty + kMaxAStages;
    uint64_t* bempty = bfull + kMaxStages;
    uint64_t* tfull_bar = bempty + kMaxStages;''')
# host: plan
rep('''struct ConvTcPlan {
    ConvTcParams prm;
    bool halo;''','''struct ConvTcPlan {
    ConvTcParams prm;
    int smem_bytes;
    bool halo;''')
rep('''    p.num_tiles = 0;
    pl->pix_per_image_tiles = p.tiles_x * p.tiles_y;
''','''    p.num_tiles = 0;
    pl->pix_per_image_tiles = p.tiles_x * p.tiles_y;
    // shared-memory plan
    p.epi_bufs = (d.k == 1 || bn <= 64) ? 2 : 1;
    const int fixed = fixed_smem_bytes(p.epi_bufs);
    if (pl->halo) {
        p.a_stages = bn == 256 ? 2 : 3;
        const int b_bytes = bn * 128;
        p.stages = (kSmemBudget - fixed - p.a_stages * kHaloABytes) / b_bytes;
        if (p.stages > 12) p.stages = 12;
        if (p.stages < 2) {
            delete pl;
            set_error("not enough shared memory for the halo weight pipeline");
            return 1;
        }
        pl->smem_bytes = p.a_stages * kHaloABytes + p.stages * b_bytes + fixed + 1024;
    } else {
        p.a_stages = 0;
        const int stage_bytes = (kTileM + bn) * bk * 2;
        p.stages = (kSmemBudget - fixed) / stage_bytes;
        if (p.stages > kMaxStages) p.stages = kMaxStages;
        pl->smem_bytes = p.stages * stage_bytes + fixed + 1024;
    }
''')
rep('''template <int BN, int BK>
static int launch_inst(const ConvTcParams& prm, int grid, cudaStream_t stream) {
    using L = SmemLayout<BN, BK>;
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           L::kTotalBytes));
        configured = true;
    }
    conv_tc_kernel<BN, BK><<<grid, kThreads, L::kTotalBytes, stream>>>(prm);
    WT_LAUNCHED();
    return 0;
}

template <int BN>
static int launch_halo(const ConvTcParams& prm, int grid, cudaStream_t stream) {
    using L = HaloSmem<BN>;
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           L::kTotalBytes));
        configured = true;
    }
    conv_halo_kernel<BN><<<grid, kThreads, L::kTotalBytes, stream>>>(prm);
    WT_LAUNCHED();
    return 0;
}
''','''template <int BN, int BK>
static int launch_inst(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBudget + 1024));
        configured = true;
    }
    conv_tc_kernel<BN, BK><<<grid, kThreads, smem, stream>>>(prm);
    WT_LAUNCHED();
    return 0;
}

template <int BN>
static int launch_halo(const ConvTcParams& prm, int smem, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        WT_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBudget + 1024));
        configured = true;
    }
    conv_halo_kernel<BN><<<grid, kThreads, smem, stream>>>(prm);
    WT_LAUNCHED();
    return 0;
}
''')
for bn in (256,128,64,32):
    rep(f'return launch_halo<{bn}>(prm, grid, stream);', f'return launch_halo<{bn}>(prm, pl->smem_bytes, grid, stream);')
    for bk in (64,32):
        rep(f'return launch_inst<{bn}, {bk}>(prm, grid, stream);', f'return launch_inst<{bn}, {bk}>(prm, pl->smem_bytes, grid, stream);')
open(p,'w').write(s)
