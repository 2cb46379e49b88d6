// placeholder until the tcgen05 3xTF32 kernel lands: reports "not supported" so the dispatcher
// uses the FP32 kernel.
#include "common.cuh"
namespace rtat_b200 {
bool sgemm_tcgen05_supported(const rtat_b200_context *, int, int, int, int, int, const float *, int,
                             const float *, int, const float *, int) { return false; }
int sgemm_tcgen05(rtat_b200_context *, int, int, int, int, int, float, const float *, int,
                  const float *, int, float, float *, int) { return RTAT_B200_STATUS_NOT_SUPPORTED; }
}
