// tf32 (fp32 storage) instantiations of the halo-tile convolution kernel (conv_halo.cuh)
#include <cstring>
#include "conv_halo.cuh"
namespace biu {
BIU_DEFINE_HALO_DISPATCH(halo_dispatch_tf32, 4)
BIU_DEFINE_HALO_PAIRS(halo_max_pairs_tf32, 4)
}  // namespace biu
