// batched_shapes.cu -- straight-line instantiations of the one-GP-per-CTA kernel for the expression
// shapes of shapes.cuh (the reference's kernel list): the same kernel template as batched.cu with a
// StaticShape policy, so the element evaluation carries no interpretation (ncu on the interpreter:
// ~700 issued instructions per gradient element of which 25 % FP64, and an 83 % instruction-cache
// hit rate on a 1.1 MB kernel; the static kernels are a few thousand instructions).
#include "batched_kernel.cuh"

namespace gpb {

int launch_batched_static(int shape, int dp, GPB_BATCHED_PARAMS) {
#define GPB_ONE_DP(DPV)                                                                  \
    if constexpr ((SH_DPMASK & (DPV)) != 0) return launch_batched_dp<DPV, true, SH>(GPB_BATCHED_ARGS); \
    else return -100;
#define GPB_SHAPE_BODY_                        \
    switch (dp) {                              \
        case 1: { GPB_ONE_DP(1) }              \
        case 2: { GPB_ONE_DP(2) }              \
        case 4: { GPB_ONE_DP(4) }              \
        case 8: { GPB_ONE_DP(8) }              \
        default: { GPB_ONE_DP(16) }            \
    }
    switch (shape) {
        GPB_SHAPE_LIST(GPB_SHAPE_CASE_)
        default: return -100;
    }
#undef GPB_SHAPE_BODY_
#undef GPB_ONE_DP
}

}  // namespace gpb
