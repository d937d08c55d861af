// Folded sliding-window kernel instances for "Net B", the skip U-Net of examples/train_functional.py:248-275 on the
// 12-channel 1-degree grid (BASELINE.json configs[2-3]).  See conv_sw.cuh; the (NCOLS, KS, ...) values are what
// tc_plan_layer picks for each layer (tests/test_tc_host_cpu.py checks that every row of this table is reachable).
#define DLWP_SW_TU_FLAGS g_tc_flags_net_b
#include "conv_sw.cuh"

namespace dlwp {
static const SwFolded kTable[] = {
    //              KH KWE NC NCOLS KS D CBLK ACT              OUT FULL
    SW_FOLDED_ENTRY(3, 1, 8, 32, 3, 2, 4, DLWP_ACT_TANH, 1, 1, "Net B conv_2d_1: 12->32 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY(3, 1, 8, 64, 3, 1, 8, DLWP_ACT_TANH, 1, 1, "Net B conv_2d_2: 16->64 3x3 tanh"),
    SW_FOLDED_ENTRY(3, 1, 8, 128, 6, 1, 16, DLWP_ACT_TANH, 1, 1, "Net B conv_2d_3: 32->128 3x3 tanh"),
    SW_FOLDED_ENTRY(3, 3, 8, 96, 8, 1, 4, DLWP_ACT_TANH, 1, 1, "Net B conv_2d_4: 128->32 3x3 tanh"),
    SW_FOLDED_ENTRY(3, 3, 8, 48, 4, 2, 2, DLWP_ACT_TANH, 1, 1, "Net B conv_2d_5: 64->16 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY(5, 5, 8, 80, 2, 1, 2, DLWP_ACT_LINEAR, 3, 0, "Net B conv_2d_6: 32->12 5x5 linear, fp32 series + P feedback"),
    SW_FOLDED_ENTRY(5, 5, 8, 80, 2, 1, 2, DLWP_ACT_LINEAR, 2, 0, "Net B conv_2d_6: 32->12 5x5 linear, fp32 only"),
    SW_FOLDED_ENTRY_X(5, 5, 8, 80, 2, 1, 2, DLWP_ACT_LINEAR, 3, 0,
                      "Net B conv_2d_6: 32->12 5x5 linear, fp32 series + P feedback + latitude-band neighbours' halo rows", 0, 1),
};
const SwFolded* sw_folded_net_b(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
int sw_flags_net_b() { return sw_tu_flags_read_clear(); }
void sw_counters_net_b(unsigned long long* acc8) { sw_tu_counters_read_clear(acc8); }
}  // namespace dlwp
