// Library-level entry points of libpn2_b200.so (see include/pn2_b200.h).
#include <atomic>
#include <cmath>
#include <cstdlib>

#include "pn2_common.cuh"

namespace pn2 {

static std::atomic<unsigned long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n)); }

// Programmatic dependent launch of the training step's kernels (pn2_common.cuh): opt-in, PN2_PDL=1.
// Measured in one call on the replayed step graph (round 2): 3.583 / 3.583 ms with it, 3.532 ms without -- inside
// a CUDA graph the kernel-to-kernel edges already cost ~0.3 us, and CTAs parked in griddepcontrol.wait hold
// SM slots the co-running geometry stream would use.  Results are identical either way (270 GPU tests).
bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("PN2_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

int ref_block_threads(int work) {
  // same expression, same libm, as _ext_src/include/cuda_utils.h:15-19
  const int pow_2 = static_cast<int>(std::log(static_cast<double>(work)) / std::log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

}  // namespace pn2

extern "C" int pn2_abi_version(void) { return PN2_ABI_VERSION; }

extern "C" unsigned long long pn2_launch_count(void) { return pn2::g_launches.load(); }

extern "C" void pn2_reset_launch_count(void) { pn2::g_launches.store(0); }

extern "C" const char *pn2_status_string(int status) {
  switch (status) {
    case PN2_OK: return "ok";
    case PN2_ERR_INVALID_ARG: return "invalid argument (null pointer or negative size)";
    case PN2_ERR_UNSUPPORTED: return "size outside the supported range";
    case PN2_ERR_ALIGNMENT: return "pointer is not 4-byte aligned";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown status";
}
