// ts_step_inst.cu -- instantiates the step / valid-move kernels for one board size.
// Compiled once per size with -DTS_S=<size> so the sizes build in parallel.
#include "ts_step.cuh"
#include "ts_valid.cuh"

#ifndef TS_S
#error "compile with -DTS_S=<board size>"
#endif
#define TS_CAT2(a, b) a##b
#define TS_CAT(a, b) TS_CAT2(a, b)

namespace ts {

cudaError_t TS_CAT(step_dispatch_s, TS_S)(const ts_step_args& a, cudaStream_t st) {
    switch (a.n_tiles) {
        case 1: return launch_step<TS_S, 1>(a, st);
        case 2: return launch_step<TS_S, 2>(a, st);
        case 3: return launch_step<TS_S, 3>(a, st);
        case 4: return launch_step<TS_S, 4>(a, st);
        case 5: return launch_step<TS_S, 5>(a, st);
        case 6: return launch_step<TS_S, 6>(a, st);
        case 7: return launch_step<TS_S, 7>(a, st);
        case 8: return launch_step<TS_S, 8>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t TS_CAT(valid_dispatch_s, TS_S)(const ts_valid_args& a, cudaStream_t st) {
    switch (a.n_tiles) {
        case 1: return launch_valid<TS_S, 1>(a, st);
        case 2: return launch_valid<TS_S, 2>(a, st);
        case 3: return launch_valid<TS_S, 3>(a, st);
        case 4: return launch_valid<TS_S, 4>(a, st);
        case 5: return launch_valid<TS_S, 5>(a, st);
        case 6: return launch_valid<TS_S, 6>(a, st);
        case 7: return launch_valid<TS_S, 7>(a, st);
        case 8: return launch_valid<TS_S, 8>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t TS_CAT(goal_dispatch_s, TS_S)(const ts_goal_args& a, cudaStream_t st) {
    switch (a.n_tiles) {
        case 1: return launch_goal<TS_S, 1>(a, st);
        case 2: return launch_goal<TS_S, 2>(a, st);
        case 3: return launch_goal<TS_S, 3>(a, st);
        case 4: return launch_goal<TS_S, 4>(a, st);
        case 5: return launch_goal<TS_S, 5>(a, st);
        case 6: return launch_goal<TS_S, 6>(a, st);
        case 7: return launch_goal<TS_S, 7>(a, st);
        case 8: return launch_goal<TS_S, 8>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ts
