// Placeholder translation unit for the fused conv->conv kernel (filled in below).
#define DLWP_SW_TU_FLAGS g_tc_flags_fused
#include "conv_sw.cuh"
namespace dlwp {
int sw_flags_fused() { return sw_tu_flags_read_clear(); }
}  // namespace dlwp
