// passive_stub.cu -- linked instead of the -DSCALAR / -DDE kernel TUs by tuning builds (make PASSIVE=0): chl_create then
// reports those contexts as not built.
#include "chl_internal.h"
namespace chl
{
ModeOps const *mode_ops_f64_strict_s1() { return nullptr; }
ModeOps const *mode_ops_f64_strict_de() { return nullptr; }
ModeOps const *mode_ops_f64_strict_de_s1() { return nullptr; }
ModeOps const *mode_ops_f64_fast_s1() { return nullptr; }
ModeOps const *mode_ops_f64_fast_de() { return nullptr; }
ModeOps const *mode_ops_f64_fast_de_s1() { return nullptr; }
}  // namespace chl
