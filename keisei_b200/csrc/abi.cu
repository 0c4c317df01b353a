// abi.cu — C-ABI plumbing shared by every entry point: version, thread-local error string.
#include "kb_common.cuh"
#include <stdarg.h>

static thread_local char g_err[1024] = {0};

void kb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void kb_count_launch(void) { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long kb_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" const char* kb_last_error(void) { return g_err; }

extern "C" int kb_abi_version(void) { return 1; }

// Device the library was compiled for; callers use it to fail loudly on anything else.
extern "C" int kb_compiled_sm(void) { return 100; }

extern "C" int kb_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}
