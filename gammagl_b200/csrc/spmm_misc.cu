// spmm_misc.cu -- the remaining element types the reference's segment ops accept
// (AT_DISPATCH_ALL_TYPES_AND2(Half,BFloat16), cpu/segment_sum_cpu.cpp:32-34): f64 and the integer
// types (degree() counts in int64, gammagl/utils/degree.py:30-40).  Scalar lanes; correctness first.
#include "spmm_impl.cuh"

namespace ggl {

template <typename T, bool W>
static int launch_scalar(const SpmmArgs& p) {
  const int lpr = lpr_for(p.K, 1);
  if (lpr == 1) return launch_modes<T, 1, 1>(p, W);
  if (lpr <= 4) return launch_modes<T, 1, 4>(p, W);
  return launch_modes<T, 1, 32>(p, W);
}

template <> int launch_spmm<double>(const SpmmArgs& p) { return launch_scalar<double, false>(p); }
template <> int launch_spmm<int8_t>(const SpmmArgs& p) { return launch_scalar<int8_t, false>(p); }
template <> int launch_spmm<uint8_t>(const SpmmArgs& p) { return launch_scalar<uint8_t, false>(p); }
template <> int launch_spmm<int16_t>(const SpmmArgs& p) { return launch_scalar<int16_t, false>(p); }
template <> int launch_spmm<int32_t>(const SpmmArgs& p) { return launch_scalar<int32_t, false>(p); }
template <> int launch_spmm<int64_t>(const SpmmArgs& p) { return launch_scalar<int64_t, false>(p); }

}  // namespace ggl
