// Thread-local error string and version of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace gnb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gnb

extern "C" int gnb_version(void) { return GNB_VERSION; }
extern "C" const char* gnb_last_error(void) { return gnb::g_err; }
extern "C" int gnb_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(GnbLiftParams);
        case 1: return (int)sizeof(GnbSampleParams);
        case 2: return (int)sizeof(GnbDecoderWeights);
        case 3: return (int)sizeof(GnbFusionParams);
        default: return -1;
    }
}
