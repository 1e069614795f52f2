def edit(path, pairs):
    s=open(path).read()
    for old,new,*cnt in pairs:
        c = s.count(old)
        assert c >= 1, (path, old[:90])
        if cnt: assert c == cnt[0], (c, old[:90])
        s=s.replace(old,new)
    open(path,'w').write(s)

# ------------------------------------------------------------------ ptx.cuh
edit('wtracker_b200/csrc/ptx.cuh', [
('''// ---------------------------------------------------------------- named barrier''','''// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// A shared::cta address of the odd CTA of a pair with bit 24 cleared is the same offset in the even
// (leader) CTA's shared memory, seen through the shared::cluster window.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\\n\\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of this barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are
// counted on the LEADER CTA's barrier.
__device__ __forceinline__ void tma_load_2d_cg2(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
        "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}

// ---------------------------------------------------------------- named barrier'''),
('''// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {''','''// CTA-pair form (issued by the leader CTA only): M = 256 (128 rows from each CTA's A tile and TMEM), the B
// operand is N/2 rows from each CTA's shared memory; descriptors are offsets valid in both CTAs.
__device__ __forceinline__ void umma_bf16_lohi_cg2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                   uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\\n\\t"
        ".reg .pred p;\\n\\t"
        ".reg .b64 da, db;\\n\\t"
        "mov.b64 da, {%1, %2};\\n\\t"
        "mov.b64 db, {%3, %4};\\n\\t"
        "setp.ne.b32 p, %6, 0;\\n\\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\\n\\t"
        "}\\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when all prior MMAs of this thread completed
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
    asm volatile(
        "{\\n\\t"
        ".reg .b16 m;\\n\\t"
        "mov.b16 m, 3;\\n\\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\\n\\t"
        "}\\n" ::"r"(smem_u32(bar))
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {'''),
('''__device__ __forceinline__ void tc_fence_before() {''','''// CTA-pair forms: executed by the same warp index in BOTH CTAs
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {'''),
])
