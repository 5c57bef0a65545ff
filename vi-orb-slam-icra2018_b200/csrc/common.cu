#include "common.cuh"

#include <cstdarg>

namespace orbb {

std::string& last_error() {
    static thread_local std::string e;
    return e;
}

int fail(orb_status st, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return (int)st;
}

}  // namespace orbb

extern "C" {
const char* orb_last_error(void) { return orbb::last_error().c_str(); }
int orb_version(void) { return ORBB200_VERSION; }
int orb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}
}
