// k_bitslice.cu -- placeholder until the bit-sliced kernel lands.
#include "sm_common.cuh"
namespace smb {
bool bitslice_supports(int, int) { return false; }
int launch_bitslice(const HotArgs &, int, cudaStream_t)
{
    set_error("bit-sliced kernel not built");
    return SM_ERR_STATE;
}
}  // namespace smb
