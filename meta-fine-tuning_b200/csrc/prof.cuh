// Launch accounting and optional per-kernel-category CUDA-event timing.
// Every kernel launch of the library sits inside a ProfScope: it always bumps the
// launch counter (bench.py reports it as gpu_launches) and, when profiling is
// enabled (mft_prof_enable), brackets the launch with two events on the launch
// stream so that bench.py can report per-category device time without a profiler.
#pragma once

#include <cuda_runtime.h>

namespace mft {

enum ProfCat {
    PC_MISC = 0,
    PC_FWD_L1, PC_FWD_L2, PC_FWD_L3, PC_FWD_L4,
    PC_SCORE, PC_SOFTMAX,
    PC_SOFTMAX_BWD, PC_DY4,
    PC_DH,
    PC_WGRAD_L1, PC_WGRAD_L2, PC_WGRAD_L3, PC_WGRAD_L4,
    PC_DGRAD_L1, PC_DGRAD_L2, PC_DGRAD_L3, PC_DGRAD_L4,
    PC_FINALIZE,
    PC_GCONV_FWD, PC_GCONV_BWD,
    PC_PREP,
    PC_COUNT
};

const char* prof_name(int cat);

struct ProfScope {
    int slot;
    cudaStream_t st;
    ProfScope(int cat, cudaStream_t stream);
    ~ProfScope();
};

}  // namespace mft
