// ld_mma.cu — tensor-core --LD path (placeholder until the tcgen05 kernel lands).
#include "engine.h"
namespace ibdgem {
bool ld_tensor_eligible(ibdgem_engine *, int32_t, int32_t, const uint8_t *) { return false; }
int ld_tensor_score(ibdgem_engine *, int32_t, const int32_t *, int32_t, const int32_t *, int32_t, double *, int32_t) {
    set_error("[::] ERROR: tensor --LD path not built.");
    return 1;
}
void ld_tensor_release(ibdgem_engine *) {}
}  // namespace ibdgem
