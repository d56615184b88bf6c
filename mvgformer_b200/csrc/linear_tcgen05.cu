// Dense projection on tcgen05 tensor cores (placeholder until the kernel lands).
#include "common.cuh"

extern "C" int mvg_linear_bf16(const void* A, const void* W, const float* bias, void* out,
                               int out_dtype, int64_t M, int Nout, int K, int64_t ldo, int relu,
                               void* stream) {
  (void)A; (void)W; (void)bias; (void)out; (void)out_dtype; (void)M; (void)Nout; (void)K;
  (void)ldo; (void)relu; (void)stream;
  mvg::set_error("mvg_linear_bf16: tcgen05 kernel not built in this revision");
  return MVG_EUNSUPPORTED;
}
