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
    PC_BWD_REGION,
    PC_SPAN_WC_FWD, PC_SPAN_GC_FWD, PC_SPAN_GC_BWD, PC_SPAN_WC_BWD,   // main-stream spans of the phases of gnn_fwd / gnn_bwd      // main-stream span of one Wcompute's wgrad + dgrad launches (they overlap on two streams)
    PC_COUNT
};

const char* prof_name(int cat);
void set_pdl_level(int v);

struct ProfScope {
    int slot;
    cudaStream_t st;
    ProfScope(int cat, cudaStream_t stream, bool count_launch = true);
    ~ProfScope();
};

}  // namespace mft

// ---------------------------------------------------------------------------------------------
// Side branches for independent small kernels (the Gconv products run on 16-60 CTAs each; issued
// back to back on one stream they leave most of the 148 SMs idle).  A Branches object forks up to
// kSideStreams library-owned non-blocking streams off the caller's stream with events and joins
// them back before the library call returns, so the caller still sees plain stream order.  Works
// the same eagerly and under stream capture (the event edges become graph dependencies).  The
// streams and events are created once per device on first use; nothing is allocated per call.
// ---------------------------------------------------------------------------------------------
namespace mft {

constexpr int kSideStreams = 5;   // 0, 1: inside one Wcompute / Gconv call; 2, 3, 4: gnn_fwd / gnn_bwd across calls

class Branches {
public:
    explicit Branches(cudaStream_t main);
    bool ok() const { return ok_; }
    // side stream i, ordered after everything enqueued on the main stream so far
    cudaStream_t fork(int i);
    // side stream i as it is (forked or not)
    cudaStream_t stream(int i) const;
    // make side stream i wait for the main stream's work enqueued so far
    void sync_to_main(int i);
    // main stream waits for side stream i
    void join(int i);
    ~Branches();   // joins whatever is still forked

private:
    cudaStream_t main_;
    void* pool_;
    bool forked_[kSideStreams];
    bool ok_;
};

}  // namespace mft
