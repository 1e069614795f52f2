// Error reporting, launch counter, device info and tensor-map encoding (driver entry point fetched
// at run time so the library links against the CUDA runtime only).
#include "../../include/wtracker_b200.h"
#include "common.cuh"

#include <stdlib.h>

#include <cudaTypedefs.h>

#include <mutex>

namespace wt {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launch_count{0};

void set_error(const std::string& msg) { t_last_error = msg; }

#ifdef WT_TUNING_KNOBS
int knob(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available (no CUDA driver / no GPU)");
        return 1;
    }
    cuuint64_t gdims[5];
    cuuint64_t gstr[4];
    cuuint32_t gbox[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    // L2 promotion of TMA loads: 128 B by default; WT_TMAP_L2=256 | 64 | 0 for A/B runs
    static const int l2_env = knob("WT_TMAP_L2", 128);
    const CUtensorMapL2promotion l2 = l2_env == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                      : l2_env == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : l2_env == 0  ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                     : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = fn(out, dtype, cuuint32_t(rank), base, gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[512];
        snprintf(buf, sizeof buf,
                 "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu] box [%u %u %u %u] "
                 "stride0 %llu swizzle %d",
                 int(r), rank, base, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                 rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                 (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), swizzle_bytes);
        set_error(buf);
        return 1;
    }
    return 0;
}

}  // namespace wt

extern "C" {

const char* wt_last_error(void) { return wt::t_last_error.c_str(); }
int wt_abi_version(void) { return WT_ABI_VERSION; }
uint64_t wt_launch_count(void) { return wt::g_launch_count.load(); }

int wt_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    WT_CHECK_CUDA(cudaGetDevice(&dev));
    // attribute queries (microseconds), not cudaGetDeviceProperties (milliseconds): this is called on launch paths
    int v = 0;
    if (sm_count) {
        WT_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        WT_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        WT_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return 0;
}

}  // extern "C"
