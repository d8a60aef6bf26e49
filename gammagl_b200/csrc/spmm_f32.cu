// spmm_f32.cu -- fp32 instantiations of the aggregation kernels (the headline path: K=100/128 fp32).
#include "spmm_impl.cuh"

namespace ggl {

template <>
int launch_spmm<float>(const SpmmArgs& p) {
  const int64_t K = p.K;
  const int64_t C = p.heads > 1 ? K / p.heads : K;
  const uintptr_t al = (uintptr_t)p.x | (uintptr_t)p.out;
  int vec = 1;
  if (K > 4) {
    if (K % 4 == 0 && C % 4 == 0 && al % 16 == 0) vec = 4;
    // per-head weights with a head width that is not a multiple of 4 (H=4, C=25 ...): 16-byte lanes that may span two heads
    // (WMODE 3) instead of four scalar column tiles -- the row is still one aligned 16-byte-multiple bulk copy
    else if (p.heads > 1 && K % 4 == 0 && C >= 4 && al % 16 == 0 && p.x_table == nullptr) vec = 4;
    else if (K % 2 == 0 && C % 2 == 0 && al % 8 == 0) vec = 2;
  }
  const int lpr = lpr_for(K, vec);
  if (vec == 4) {
    if (lpr <= 4) return launch_modes<float, 4, 4>(p, true);
    if (lpr == 8) return launch_modes<float, 4, 8>(p, true);
    if (lpr == 16) return launch_modes<float, 4, 16>(p, true);
    return launch_modes<float, 4, 32>(p, true);
  }
  if (vec == 2) return launch_modes<float, 2, 32>(p, true);
  if (lpr == 1) return launch_modes<float, 1, 1>(p, true);
  if (lpr <= 4) return launch_modes<float, 1, 4>(p, true);
  return launch_modes<float, 1, 32>(p, true);
}

}  // namespace ggl
