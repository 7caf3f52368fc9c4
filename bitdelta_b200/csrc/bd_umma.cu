// tcgen05 / TMEM / TMA forward kernel (placeholder until the kernel lands; everything routes to the SIMT kernel).
#include "bd_common.cuh"
namespace bd {
bool umma_supports(const FwdProblem&, const char** why) { *why = "tcgen05 kernel not built yet"; return false; }
int launch_fwd_umma(const FwdProblem&) { return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel not built yet"); }
size_t umma_workspace_bytes(int64_t, int64_t) { return 0; }
}  // namespace bd
