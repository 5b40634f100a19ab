// api.cu -- library-level entry points of libhopk.so (error string, version).
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
}  // namespace hopk

extern "C" const char* hopk_last_error(void) { return hopk::err_buf(); }
extern "C" int hopk_version(void) { return 100; }
extern "C" long long hopk_launch_count(void) { return hopk::launch_counter(); }
