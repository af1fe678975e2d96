// tf32 (fp32 storage) instantiations of the row-streaming folded-tap convolution kernel (conv_rows.cuh)
#include "conv_rows.cuh"
namespace biu {
BIU_DEFINE_ROWS_DISPATCH(rows_dispatch_tf32, 4)
}  // namespace biu
