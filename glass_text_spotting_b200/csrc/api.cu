// api.cu -- process-wide state of libglass_b200.so: error string, launch counter, driver entry points.
#include <atomic>
#include <mutex>
#include <string>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(const std::string& msg) {
  g_last_error = msg;
  return -1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
    }
  });
  return fn;
}

int num_sms() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    sms = v;
  }
  return sms;
}

}  // namespace glass

extern "C" const char* glass_last_error(void) { return glass::g_last_error.c_str(); }
extern "C" int glass_abi_version(void) { return 2; }
extern "C" int64_t glass_launch_count(void) { return glass::g_launches.load(); }
