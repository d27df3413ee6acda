// batched_generic.cu -- the one-GP-per-CTA kernel for arbitrary expressions (per-dimension
// lengthscales, more than four leaves): run-time interpreter with a per-thread gradient array.
// Its own translation unit only to keep the build parallel (see batched_kernel.cuh).
#include "batched_kernel.cuh"

namespace gpb {

int launch_batched_generic(int dp, GPB_BATCHED_PARAMS) {
    switch (dp) {
        case 1: return launch_batched_dp<1, false>(GPB_BATCHED_ARGS);
        case 2: return launch_batched_dp<2, false>(GPB_BATCHED_ARGS);
        case 4: return launch_batched_dp<4, false>(GPB_BATCHED_ARGS);
        case 8: return launch_batched_dp<8, false>(GPB_BATCHED_ARGS);
        default: return launch_batched_dp<16, false>(GPB_BATCHED_ARGS);
    }
}

}  // namespace gpb
