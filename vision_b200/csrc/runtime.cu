// runtime.cu — error plumbing, device attribute cache, ABI version.
#include "common.cuh"

#include <cstring>

namespace vb200 {

std::atomic<uint64_t> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
}

namespace {
struct DevAttr { int sms = 0; int smem_optin = 0; bool init = false; };
DevAttr g_attr[64];
DevAttr& attr() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  DevAttr& a = g_attr[dev];
  if (!a.init) {
    cudaDeviceGetAttribute(&a.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&a.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (a.sms <= 0) a.sms = kNumSMsB200;
    if (a.smem_optin <= 0) a.smem_optin = 232448;
    a.init = true;
  }
  return a;
}
}  // namespace

int sm_count() { return attr().sms; }
int max_smem_optin() { return attr().smem_optin; }

}  // namespace vb200

extern "C" int vb200_abi_version(void) { return VB200_ABI_VERSION; }
extern "C" const char* vb200_last_error(void) { return vb200::last_error_buf(); }
extern "C" uint64_t vb200_launch_count(void) { return vb200::g_launch_count.load(); }
