// Folded sliding-window kernel instances for the non-skip encoder-decoder of examples/train.py:159-219 and
// examples/train_functional.py:222-245 (C -> 32 -> pool -> 64 -> pool -> 128 -> up -> 64 -> up -> 32 -> C_out); the first
// and last layers coincide with Net B's (conv_sw_net_b.cu).  The 64->128 and 128->64 layers are not here: their split
// weight images (294 KB) exceed the shared memory the sliding-window kernel keeps weights resident in, so that net stays
// on the fp32 FFMA kernels as a whole.  See conv_sw.cuh.
#define DLWP_SW_TU_FLAGS g_tc_flags_net_basic
#include "conv_sw.cuh"

namespace dlwp {
static const SwFolded kTable[] = {
    //              KH KWE NC NCOLS KS D CBLK ACT              OUT FULL
    SW_FOLDED_ENTRY(3, 1, 8, 64, 6, 1, 8, DLWP_ACT_TANH, 1, 1, "basic conv 2: 32->64 3x3 tanh"),
    SW_FOLDED_ENTRY(3, 1, 8, 32, 12, 2, 4, DLWP_ACT_TANH, 1, 1, "basic conv 5: 64->32 3x3 dil 2 tanh"),
};
const SwFolded* sw_folded_net_basic(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
int sw_flags_net_basic() { return sw_tu_flags_read_clear(); }
void sw_counters_net_basic(unsigned long long* acc8) { sw_tu_counters_read_clear(acc8); }
}  // namespace dlwp
