// spmm_bf16.cu -- bf16 storage / fp32 accumulate instantiations (halves the gathered bytes per edge).
#include "spmm_impl.cuh"

namespace ggl {

template <>
int launch_spmm<__nv_bfloat16>(const SpmmArgs& p) {
  using T = __nv_bfloat16;
  const int64_t K = p.K;
  const int64_t C = p.heads > 1 ? K / p.heads : K;
  const uintptr_t al = (uintptr_t)p.x | (uintptr_t)p.out;
  int vec = 1;
  if (K > 4) {
    if (K % 8 == 0 && C % 8 == 0 && al % 16 == 0) vec = 8;
    else if (K % 4 == 0 && C % 4 == 0 && al % 8 == 0) vec = 4;
    else if (K % 2 == 0 && C % 2 == 0 && al % 4 == 0) vec = 2;
  }
  const int lpr = lpr_for(K, vec);
  if (vec == 8) {
    if (lpr <= 4) return launch_modes<T, 8, 4>(p, true);
    if (lpr == 8) return launch_modes<T, 8, 8>(p, true);
    if (lpr == 16) return launch_modes<T, 8, 16>(p, true);
    return launch_modes<T, 8, 32>(p, true);
  }
  if (vec == 4) return launch_modes<T, 4, 32>(p, true);
  if (vec == 2) return launch_modes<T, 2, 32>(p, true);
  if (lpr == 1) return launch_modes<T, 1, 1>(p, true);
  if (lpr <= 4) return launch_modes<T, 1, 4>(p, true);
  return launch_modes<T, 1, 32>(p, true);
}

}  // namespace ggl
