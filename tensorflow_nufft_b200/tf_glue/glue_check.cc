// glue_check.cc -- translation unit that type-checks nufft_kernels_b200.cc without TensorFlow:
//   g++ -std=c++17 -fsyntax-only -Ioracle/ref_build/shim -I/root/reference -Iinclude ... glue_check.cc
// (run by __graft_entry__.build() when the reference tree is present). The reference's own
// nufft_plan.h supplies TransformType / FftDirection / Options / GPUDevice exactly as the OpKernel
// sees them; the TF framework types come from the stand-in headers in oracle/ref_build/shim/.
// The reference headers are parsed in their CPU configuration (their GOOGLE_CUDA sections need
// cuFFT and StreamExecutor headers); GOOGLE_CUDA is defined afterwards, for the glue only.
#include "tensorflow_nufft/cc/kernels/nufft_plan.h"

namespace tensorflow {
namespace nufft {
enum class OpType { NUFFT, INTERP, SPREAD };   // nufft_kernels.cc:37 (local to that file)
}  // namespace nufft
}  // namespace tensorflow

#define GOOGLE_CUDA 1
#define B200NUFFT_WITH_TENSORFLOW 1
#include "nufft_kernels_b200.cc"
