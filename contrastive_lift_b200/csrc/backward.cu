// placeholder until the backward family lands (kept so every symbol of include/clift_b200.h exists)
#include "launchers.h"
using namespace clift;
extern "C" int32_t clift_render_backward(const clift_render_cfg*, const clift_field*, const float*, const float*, int64_t, int32_t,
                                         void*, int64_t, int64_t, const clift_render_out*, const float*, const float*, const float*,
                                         const float*, const clift_field_grad*, void*) {
    set_error("clift_render_backward: not implemented in this build");
    return CLIFT_ERR_UNSUPPORTED;
}
