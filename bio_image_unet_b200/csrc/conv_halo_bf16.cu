// bf16 instantiations of the halo-tile convolution kernel (conv_halo.cuh)
#include "conv_halo.cuh"
namespace biu {
BIU_DEFINE_HALO_DISPATCH(halo_dispatch_bf16, 2)
}  // namespace biu
