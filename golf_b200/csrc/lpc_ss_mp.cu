// lpc_ss_mp.cu -- one translation unit per tap-count bucket GOLF_MP (compiled 8x in
// parallel by golf_b200/build.py): explicit instantiations of the chunk-response and
// chunk-solve kernels for the direct (forward) and transposed (adjoint) forms.
#include "lpc_ss.cuh"

#ifndef GOLF_MP
#error "compile with -DGOLF_MP=<4|8|12|16|20|24|32|40>"
#endif

namespace golf {
template int launch_mp<GOLF_MP, 0>(const SsParams&, bool, int, cudaStream_t);
template int launch_mp<GOLF_MP, 1>(const SsParams&, bool, int, cudaStream_t);
}  // namespace golf
