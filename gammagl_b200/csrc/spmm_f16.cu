// spmm_f16.cu -- f16 storage / fp32 accumulate instantiations.
#include "spmm_impl.cuh"

namespace ggl {
template <> int launch_spmm<__half>(const SpmmArgs& p) { return launch_16bit<__half>(p); }
}  // namespace ggl
