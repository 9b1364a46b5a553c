// Library-wide plumbing: thread-local error string, device attribute cache, version.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = {0};

void b2l_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void b2l_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t b2l_launch_count(void) { return (int64_t)g_launches.load(); }

// Programmatic dependent launch switch: env B2L_PDL=0 disables it at load, b2l_set_pdl() at run time
// (tests compare both modes; a captured CUDA graph keeps the mode it was captured with).
#include <stdlib.h>
static int g_pdl = -1;
int b2l_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("B2L_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl;
}
extern "C" int b2l_set_pdl(int enabled) {
  const int old = b2l_pdl_enabled();
  g_pdl = enabled ? 1 : 0;
  return old;
}

extern "C" const char* b2l_last_error(void) { return g_err; }
extern "C" int b2l_version(void) { return 100; }

static int g_sms[64] = {0};
static int g_smem[64] = {0};

int b2l_num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    g_sms[dev] = v;
  }
  return g_sms[dev];
}

int b2l_max_smem_optin() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
  if (g_smem[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0)
      v = 227 * 1024;
    g_smem[dev] = v;
  }
  return g_smem[dev];
}

extern "C" int b2l_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin) {
  int dev = 0;
  B2L_CUDA_CHECK(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) { B2L_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
  if (cc_major) { B2L_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
  if (cc_minor) { B2L_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
  if (max_smem_optin) {
    B2L_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    *max_smem_optin = v;
  }
  return B2L_OK;
}

// ---- device-side hang watchdog: host half (common.cuh, "device-side hang watchdog") ----
static unsigned int* g_hang_rec = nullptr;
unsigned int* b2l_hang_record_ptr() {
  if (!g_hang_rec) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, 64 * sizeof(unsigned int), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    memset(p, 0, 64 * sizeof(unsigned int));
    g_hang_rec = (unsigned int*)p;
  }
  return g_hang_rec;
}
unsigned long long b2l_hang_limit_ns() {
  const char* e = getenv("B2L_HANG_MS");
  long long ms = (e && e[0]) ? atoll(e) : 4000;
  if (ms < 0) ms = 0;
  return (unsigned long long)ms * 1000000ull;
}
// Copies the 16-word hang record (all zero = no kernel of this library has trapped in a barrier wait):
// out[0] = 0xB2D00DED when valid, [1] = FNV-1a hash of the source file's basename, [2] = line of the wait,
// [3..5] = blockIdx, [6] = threadIdx.x, [7] = shared-memory address of the barrier, [8] = parity waited for,
// [9] = milliseconds waited, [10..12] = gridDim, [13] = blockDim.x.  Readable after the context has died.
extern "C" int b2l_hang_report(int* out16) {
  B2L_REQUIRE(out16, "hang_report: null pointer");
  for (int i = 0; i < 16; ++i) out16[i] = g_hang_rec ? (int)g_hang_rec[i] : 0;
  return B2L_OK;
}
