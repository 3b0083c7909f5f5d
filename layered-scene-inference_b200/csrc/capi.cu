// C-ABI plumbing shared by every entry point: version, thread-local error text, launch counter.
#include <atomic>
#include <string.h>

#include "capi_common.h"

namespace lsi {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lsi

extern "C" int lsi_b200_version(void) { return 100; }
extern "C" const char* lsi_b200_last_error(void) { return lsi::g_err; }
extern "C" unsigned long long lsi_b200_launch_count(void) { return lsi::g_launches.load(std::memory_order_relaxed); }
