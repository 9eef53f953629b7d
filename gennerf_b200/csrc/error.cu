// Thread-local error string, version and tuning options of the C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace gnb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Tuning / debugging options.  Each one takes its default from the environment variable of the same name ONCE, when the
// first option is read (not on every hot-path call); gnb_set_option changes it afterwards.
static const char* const kOptNames[OPT_COUNT] = {
    "GNB_TC_TWO_CTA", "GNB_TC_NO_EARLY", "GNB_DEBUG_MAX_CLUSTERS", "GNB_DEBUG_PRINT", "GNB_LIFT_NVW", "GNB_SCATTER_SCALAR",
    "GNB_FPS_SINGLE_CTA", "GNB_FPS_CLUSTER", "GNB_SAMPLE_GENERIC", "GNB_BIN_UNIT", "GNB_BIN_ROWCOPY", "GNB_SCATTER_TILED",
    "GNB_BIN_PRESORTED", "GNB_TC_NO_STG", "GNB_TC_PAIR", "GNB_DEBUG_NO_WCOPY", "GNB_QUERY_FUSED", "GNB_FPS_GRID"};
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static void opt_init() {
    for (int i = 0; i < OPT_COUNT; ++i) {
        const char* e = getenv(kOptNames[i]);
        int v = 0;
        if (e) { v = atoi(e); if (v == 0 && e[0] != '0') v = 1; }      // "GNB_X=anything" switches a flag on
        g_opt[i].store(v, std::memory_order_relaxed);
    }
}
int opt(int which) {
    std::call_once(g_opt_once, opt_init);
    return g_opt[which].load(std::memory_order_relaxed);
}
static int opt_index(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, kOptNames[i]) == 0) return i;
    return -1;
}
}  // namespace gnb

extern "C" int gnb_set_option(const char* name, int value) {
    const int i = gnb::opt_index(name);
    if (i < 0) { gnb::set_error("gnb_set_option: unknown option %s", name ? name : "(null)"); return GNB_E_INVALID; }
    std::call_once(gnb::g_opt_once, gnb::opt_init);
    gnb::g_opt[i].store(value, std::memory_order_relaxed);
    return 0;
}
extern "C" int gnb_get_option(const char* name, int* value) {
    const int i = gnb::opt_index(name);
    if (i < 0 || !value) { gnb::set_error("gnb_get_option: unknown option %s", name ? name : "(null)"); return GNB_E_INVALID; }
    *value = gnb::opt(i);
    return 0;
}

extern "C" int gnb_version(void) { return GNB_VERSION; }
extern "C" const char* gnb_last_error(void) { return gnb::g_err; }
extern "C" int gnb_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(GnbLiftParams);
        case 1: return (int)sizeof(GnbSampleParams);
        case 2: return (int)sizeof(GnbDecoderWeights);
        case 3: return (int)sizeof(GnbFusionParams);
        case 4: return (int)sizeof(GnbDecoderGrads);
        default: return -1;
    }
}
