// linked instead of the fp32 kernel TU when CHL_NO_F32 is requested at build time
#include "chl_internal.h"
namespace chl { ModeOps const *mode_ops_f32_fast() { return nullptr; } }
