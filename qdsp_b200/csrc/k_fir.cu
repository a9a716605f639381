// placeholder until the register-blocked dense FIR lands
#include <new>
#include "internal.cuh"
#include "kernels.cuh"
namespace qdsp {
struct FirPlan { int T; };
FirPlan* fir_plan_create(const float*, int) { return nullptr; }
void fir_plan_destroy(FirPlan* p) { delete p; }
int launch_fir_dense(FirPlan*, const float2*, int, const float2*, long long, int, float2*, cudaStream_t) {
    set_last_error("dense FIR kernel not built");
    return -1;
}
}  // namespace qdsp
