// placeholder until the tcgen05 kernel lands
#include <cstdio>
#include "va_common.cuh"
namespace va {
struct FusedPlan { int dummy; };
FusedPlan* fused_plan_create(const Dims&, int, char* err, size_t errlen) { snprintf(err, errlen, "not built"); return nullptr; }
void fused_plan_destroy(FusedPlan* p) { delete p; }
cudaError_t launch_fused(FusedPlan*, const Dims&, const float*, const float*, const float*, const int*, int, uint8_t*,
                         float*, InstStats*, unsigned int*, cudaStream_t, char*, size_t) { return cudaErrorNotSupported; }
}
