// bf16 instantiations of the halo-tile convolution kernel (conv_halo.cuh)
#include <cstring>
#include "conv_halo.cuh"
namespace biu {
BIU_DEFINE_HALO_DISPATCH(halo_dispatch_bf16, 2)
BIU_DEFINE_HALO_PAIRS(halo_max_pairs_bf16, 2)
}  // namespace biu
