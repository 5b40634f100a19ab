// api.cu -- library-level entry points of libhopk.so (error string, version) and the little per-device state the
// library keeps: the dropout epoch (one device-side counter, shared by every attention kernel) and the record of which
// kernels already had their dynamic shared-memory limit raised on which device.
#include <mutex>
#include <set>
#include <utility>
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {
char* err_buf()
{
    static thread_local char buf[ERR_BUF] = {0};
    return buf;
}
long long& launch_counter()
{
    static long long n = 0;
    return n;
}

// Added to every attention call's seed; advanced by a (graph-capturable) kernel so that replays of a captured training
// step do not repeat the dropout mask that was baked into the launch arguments at capture time.  0 until advanced.
// One instance per device (a __device__ variable is instantiated on every device the module is loaded on).
__device__ unsigned long long g_drop_epoch = 0ull;
__global__ void drop_epoch_kernel(int reset)
{
    if (reset) g_drop_epoch = 0ull; else g_drop_epoch += 0x9E3779B97F4A7C15ull;
}

const unsigned long long* drop_epoch_ptr()
{
    constexpr int MAXDEV = 64;
    static const unsigned long long* table[MAXDEV] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    if (!table[dev]) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_drop_epoch) != cudaSuccess) return nullptr;
        table[dev] = (const unsigned long long*)p;
    }
    return table[dev];
}

cudaError_t configure_smem_once(const void* func, size_t bytes)
{
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;          // (device, kernel): the attribute is per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, func})) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done.insert({dev, func});
    return e;
}
}  // namespace hopk

extern "C" const char* hopk_last_error(void) { return hopk::err_buf(); }
extern "C" int hopk_version(void) { return 200; }
extern "C" long long hopk_launch_count(void) { return hopk::launch_counter(); }

extern "C" int hopk_dropout_epoch_advance(int reset, void* stream)
{
    hopk::drop_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reset);
    HOPK_LAUNCH_CHECK("dropout_epoch");
    return 0;
}
