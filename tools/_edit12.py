def edit(path, pairs):
    s=open(path).read()
    for old,new,*cnt in pairs:
        c = s.count(old)
        assert c >= 1, (path, old[:90])
        if cnt: assert c == cnt[0], (c, old[:90])
        s=s.replace(old,new)
    open(path,'w').write(s)

edit('wtracker_b200/csrc/ptx.cuh', [
('''// ---------------------------------------------------------------- CTA pairs (cta_group::2)''','''// ---------------------------------------------------------------- programmatic dependent launch
// launch_dependents: the next kernel of the stream (if launched with the programmatic-serialization
// attribute) may start running its prologue once every CTA of this grid has executed this or exited.
// grid_dependency_wait: blocks until the previous grid has completed and its memory is visible; every
// global read of activations and every global write must come after it.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- CTA pairs (cta_group::2)'''),
])

p='wtracker_b200/csrc/conv_tcgen05.cu'
edit(p, [
# default cg = 1
('''getenv("WT_CONV_CG") ? atoi(getenv("WT_CONV_CG")) : 2;''','''getenv("WT_CONV_CG") ? atoi(getenv("WT_CONV_CG")) : 1;'''),
('''    // CTA pairs (cta_group::2, M = 256): each CTA supplies half of the weight tile, which cuts the shared-memory
    // bytes per FLOP by a third for N >= 128 — the binding resource of these kernels (DESIGN.md section 5)''','''    // CTA pairs (cta_group::2, M = 256, each CTA supplies half of the weight tile): implemented and parity-tested,
    // but measured 15-30 % SLOWER than single-CTA MMAs on these layer shapes (B200, round 1), so off by default;
    // WT_CONV_CG=2 selects it for N >= 128.'''),
# kernel: after the setup sync, allow dependents + (roles wait individually)
('''    if (CG == 2) ptx::cluster_sync();   // the peer's barriers are initialised before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;''','''    if (CG == 2) ptx::cluster_sync();   // the peer's barriers are initialised before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barriers, TMEM, bias = constants) overlapped the tail of the
    // previous kernel; the next kernel may start its own prologue now.  Activations are only touched after
    // grid_dependency_wait() (producer and epilogue warps; the MMA warps never touch global memory).
    ptx::grid_launch_dependents();''', 2),
# generic producer
('''        // The whole warp walks the loop (uniform control flow); one elected lane issues the copies.
        int stage = 0;
        uint32_t phase = 0;''','''        // The whole warp walks the loop (uniform control flow); one elected lane issues the copies.
        ptx::grid_dependency_wait();
        int stage = 0;
        uint32_t phase = 0;'''),
# halo producer
('''        // TMA producer: warp-uniform loop, one elected lane issues
        int sa = 0, sb = 0;''','''        // TMA producer: warp-uniform loop, one elected lane issues
        ptx::grid_dependency_wait();
        int sa = 0, sb = 0;'''),
# epilogue
('''    uint32_t unit_counter = 0;
    int it = g;
    const int first = blockIdx.x / CG, step = gridDim.x / CG;''','''    uint32_t unit_counter = 0;
    int it = g;
    const int first = blockIdx.x / CG, step = gridDim.x / CG;
    ptx::grid_dependency_wait();   // residual loads and output stores come after the previous kernel'''),
# launch attr
('''    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cg > 1 ? 1 : 0;''','''    cudaLaunchAttribute attr[2];
    int na = 0;
    static const int pdl_env = getenv("WT_CONV_PDL") ? atoi(getenv("WT_CONV_PDL")) : 1;
    if (pdl_env) {   // start this kernel's prologue while the previous kernel of the stream drains
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cg > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cg;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;'''),
])
