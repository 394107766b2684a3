// screen.cu -- placeholder until the tcgen05 screen lands (next commit)
#include "common.cuh"
#include "kernels.h"
namespace b2k {
struct ScreenPlan {};
bool screen_supported(const b2k_ctx*, int, int, int64_t) { return false; }
int screen_plan_create(b2k_ctx*, int64_t, int, int, ScreenPlan**) { return set_error(B2K_ERR_INVALID_ARG, "screen not built"); }
void screen_plan_destroy(ScreenPlan*) {}
int screen_prepare_frames(ScreenPlan*, const float*, int64_t) { return set_error(B2K_ERR_INVALID_ARG, "screen not built"); }
int screen_assign(ScreenPlan*, const float*, int64_t, const float*, int32_t*, float*, int) { return set_error(B2K_ERR_INVALID_ARG, "screen not built"); }
}
