// Sliding-window kernel instances of the plain-bf16 mode (DlwpPlanOptions.precision = 1; BASELINE.json configs[2]: "U-Net
// DLWPFunctional, 12-chan 1 degree grid, bf16"): activations and weights are bf16 (one plane per 8-channel chunk), ONE
// tcgen05 MMA pass per K step, fp32 accumulation in TMEM, bias / tanh in fp32.  Generic instances for every layer shape the
// planner accepts, folded ones for Net B's layers.  See conv_sw.cuh.
#define DLWP_SW_TU_FLAGS g_tc_flags_bf16
#include "conv_sw.cuh"

namespace dlwp {
static const SwFolded kTable[] = {
    //                   KH KWE NC NCOLS KS D CBLK ACT              OUT FULL
    SW_FOLDED_ENTRY_BF16(3, 1, 8, 32, 2, 2, 4, DLWP_ACT_TANH, 1, 1, "bf16 Net A conv1: 6->32 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY_BF16(5, 5, 6, 32, 2, 1, 1, DLWP_ACT_LINEAR, 3, 1, "bf16 Net A conv2: 32->6 5x5 linear, fp32 series + P feedback"),
    SW_FOLDED_ENTRY_X(5, 5, 6, 32, 2, 1, 1, DLWP_ACT_LINEAR, 3, 1,
                      "bf16 Net A conv2: 32->6 5x5 linear, fp32 series + P feedback + latitude-band neighbours' halo rows", 1, 1),
    SW_FOLDED_ENTRY_BF16(3, 1, 8, 32, 3, 2, 4, DLWP_ACT_TANH, 1, 1, "bf16 Net B conv_2d_1: 12->32 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY_BF16(3, 1, 8, 64, 3, 1, 8, DLWP_ACT_TANH, 1, 1, "bf16 Net B conv_2d_2: 16->64 3x3 tanh"),
    SW_FOLDED_ENTRY_BF16(3, 1, 8, 128, 6, 1, 16, DLWP_ACT_TANH, 1, 1, "bf16 Net B conv_2d_3: 32->128 3x3 tanh"),
    SW_FOLDED_ENTRY_BF16(3, 3, 8, 96, 8, 1, 4, DLWP_ACT_TANH, 1, 1, "bf16 Net B conv_2d_4: 128->32 3x3 tanh"),
    SW_FOLDED_ENTRY_BF16(3, 3, 8, 48, 4, 2, 2, DLWP_ACT_TANH, 1, 1, "bf16 Net B conv_2d_5: 64->16 3x3 dil 2 tanh"),
    SW_FOLDED_ENTRY_BF16(5, 5, 8, 80, 2, 1, 2, DLWP_ACT_LINEAR, 3, 0, "bf16 Net B conv_2d_6: 32->12 5x5 linear, fp32 series + P feedback"),
    SW_FOLDED_ENTRY_X(5, 5, 8, 80, 2, 1, 2, DLWP_ACT_LINEAR, 3, 0,
                      "bf16 Net B conv_2d_6: 32->12 5x5 linear, fp32 series + P feedback + latitude-band neighbours' halo rows", 1, 1),
};
const SwFolded* sw_folded_bf16(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
int sw_flags_bf16() { return sw_tu_flags_read_clear(); }
void sw_counters_bf16(unsigned long long* acc12) { sw_tu_counters_read_clear(acc12); }

bool sw_launch_generic_bf16(int kh, int kw_eff, int nc, const SwParams& p, const CUtensorMap& map_full,
                            const CUtensorMap& map_pair, int grid, size_t smem, cudaStream_t stream) {
    if (kh == 3 && kw_eff == 1) sw_launch_one<3, 1, 8, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else if (kh == 5 && kw_eff == 1) sw_launch_one<5, 1, 8, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else if (kh == 3 && kw_eff == 3 && nc == 8) sw_launch_one<3, 3, 8, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else if (kh == 3 && kw_eff == 3) sw_launch_one<3, 3, 6, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else if (kh == 5 && kw_eff == 5 && nc == 8) sw_launch_one<5, 5, 8, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else if (kh == 5 && kw_eff == 5) sw_launch_one<5, 5, 6, SwGenericBf16>(p, map_full, map_pair, grid, smem, stream);
    else return false;
    return true;
}
void sw_preload_generic_bf16() {
    sw_preload_one<3, 1, 8, SwGenericBf16>(); sw_preload_one<5, 1, 8, SwGenericBf16>(); sw_preload_one<3, 3, 8, SwGenericBf16>();
    sw_preload_one<3, 3, 6, SwGenericBf16>(); sw_preload_one<5, 5, 8, SwGenericBf16>(); sw_preload_one<5, 5, 6, SwGenericBf16>();
}
}  // namespace dlwp
