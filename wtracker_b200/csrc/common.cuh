// Shared host-side helpers: error reporting, launch counting, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

namespace wt {

void set_error(const std::string& msg);             // thread-local last error
extern std::atomic<uint64_t> g_launch_count;        // kernels launched by this library

#define WT_CHECK_CUDA(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::wt::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                            __FILE__ + ":" + std::to_string(__LINE__));                          \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define WT_REQUIRE(cond, msg)                                                              \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::wt::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +      \
                            std::to_string(__LINE__));                                     \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

// count + check a kernel launch
#define WT_LAUNCHED()                                  \
    do {                                               \
        ::wt::g_launch_count.fetch_add(1);             \
        WT_CHECK_CUDA(cudaGetLastError());             \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Tuning knobs: the product library never reads the environment — every knob IS its default.  A tuning build
// (`python -m wtracker_b200.build --tuning`, -DWT_TUNING_KNOBS: tools/gpu_r2c.sh A/B runs only) reads WT_* variables.
#ifdef WT_TUNING_KNOBS
int knob(const char* name, int dflt);
#else
constexpr int knob(const char*, int dflt) { return dflt; }
#endif

// Opt-in for > 48 KB of dynamic shared memory.  Function attributes belong to the (device, context) pair, so the
// "already raised to N bytes" state is kept PER DEVICE (one engine per GPU in one process is legal: DetectorEngine,
// ResMLPEngine and HotPath all take a `device`), behind a mutex (launch paths are re-entrant per engine + stream).
struct SmemOptIn {
    static constexpr int kMaxDevices = 64;
    std::mutex m;
    size_t bytes[kMaxDevices] = {};
};
template <typename Kernel>
inline cudaError_t opt_in_smem(Kernel kernel, SmemOptIn& st, size_t want) {
    // (no "dynamic <= 48 KB needs nothing" shortcut: the limit applies to static + dynamic shared memory together)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= SmemOptIn::kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(want));
    std::lock_guard<std::mutex> lock(st.m);
    if (want <= st.bytes[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(want));
    if (e == cudaSuccess) st.bytes[dev] = want;
    return e;
}

// Launch with the programmatic-stream-serialization attribute: the kernel may start (up to its
// griddepcontrol.wait) while the previous kernel of the stream drains.  The kernel MUST execute
// griddepcontrol.wait before touching memory written by earlier kernels.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Encodes a tiled tensor map (rank <= 5). dims/box in elements (innermost first), strides in bytes
// for dims 1..rank-1. swizzle_bytes in {0, 32, 64, 128}. Returns 0 on success.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

}  // namespace wt
