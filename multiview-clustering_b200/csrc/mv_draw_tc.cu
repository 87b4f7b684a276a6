// mv_draw_tc.cu — likelihood + draw on the tcgen05 tensor cores (engine MVG_ENGINE_TCGEN05).
// Placeholder until the TMA/TMEM kernel lands: reports "unsupported" so AUTO selects the SIMT engine.
#include "mv_ctx.h"

namespace mv {
bool draw_tc_supported(const Ctx&) { return false; }
cudaError_t launch_draw_tc(const Ctx&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace mv
