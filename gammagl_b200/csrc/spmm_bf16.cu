// spmm_bf16.cu -- bf16 storage / fp32 accumulate instantiations (halves the gathered bytes per edge).
#include "spmm_impl.cuh"

namespace ggl {
template <> int launch_spmm<__nv_bfloat16>(const SpmmArgs& p) { return launch_16bit<__nv_bfloat16>(p); }
}  // namespace ggl
