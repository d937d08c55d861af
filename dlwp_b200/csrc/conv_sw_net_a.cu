// Folded sliding-window kernel instances for "Net A", the 2-layer benchmark net of BASELINE.json configs[0-1]
// (first and last conv blocks of examples/train.py:159-169, :211-219).  See conv_sw.cuh.
#define DLWP_SW_TU_FLAGS g_tc_flags_net_a
#include "conv_sw.cuh"

namespace dlwp {
static const SwFolded kTable[] = {
    //              KH KWE NC NCOLS KS D CBLK ACT              OUT FULL
    SW_FOLDED_ENTRY(3, 1, 8, 32, 2, 2, 4, DLWP_ACT_TANH, 1, 1, "Net A conv1: 6->32 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY(5, 5, 6, 32, 2, 1, 1, DLWP_ACT_LINEAR, 3, 1, "Net A conv2: 32->6 5x5 linear, fp32 series + P feedback"),
    SW_FOLDED_ENTRY(5, 5, 6, 32, 2, 1, 1, DLWP_ACT_LINEAR, 2, 1, "Net A conv2: 32->6 5x5 linear, fp32 only"),
    SW_FOLDED_ENTRY_X(5, 5, 6, 32, 2, 1, 1, DLWP_ACT_LINEAR, 3, 1,
                      "Net A conv2: 32->6 5x5 linear, fp32 series + P feedback + latitude-band neighbours' halo rows", 0, 1),
};
const SwFolded* sw_folded_net_a(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
int sw_flags_net_a() { return sw_tu_flags_read_clear(); }
void sw_counters_net_a(unsigned long long* acc8) { sw_tu_counters_read_clear(acc8); }
}  // namespace dlwp
