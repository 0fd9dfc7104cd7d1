#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "prof.cuh"

namespace mft {

static const char* kNames[PC_COUNT] = {
    "misc",
    "fwd_gemm_l1", "fwd_gemm_l2", "fwd_gemm_l3", "fwd_gemm_l4",
    "score", "softmax",
    "softmax_bwd", "dy4",
    "bn_bwd_dh",
    "wgrad_l1", "wgrad_l2", "wgrad_l3", "wgrad_l4",
    "dgrad_l1", "dgrad_l2", "dgrad_l3", "dgrad_l4",
    "finalize_grads",
    "gconv_fwd", "gconv_bwd",
    "prep",
    "bwd_gemm_region",
    "span_wcompute_fwd", "span_gconv_fwd", "span_gconv_bwd", "span_wcompute_bwd",
};

const char* prof_name(int cat) { return (cat >= 0 && cat < PC_COUNT) ? kNames[cat] : "?"; }

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<bool> g_enabled{false};
static std::mutex g_mu;
struct Slot { cudaEvent_t a, b; int cat; };
static std::vector<Slot> g_slots;
static size_t g_used = 0;
constexpr size_t kMaxSlots = 16384;

bool prof_enabled() { return g_enabled.load(std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};
int pdl_level() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("MFT_PDL");
        v = (e && *e) ? atoi(e) : 2;
        if (v < 0) v = 0;
        g_pdl.store(v);
    }
    return v;
}
void set_pdl_level(int v) { g_pdl.store(v < 0 ? 0 : v); }

ProfScope::ProfScope(int cat, cudaStream_t stream, bool count_launch) : slot(-1), st(stream) {
    if (count_launch) g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_used >= kMaxSlots) return;
    if (g_used >= g_slots.size()) {
        Slot s;
        if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
        g_slots.push_back(s);
    }
    slot = (int)g_used++;
    g_slots[slot].cat = cat;
    cudaEventRecord(g_slots[slot].a, st);
}

ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_slots[slot].b, st);
}

}  // namespace mft

using namespace mft;

extern "C" {

unsigned long long mft_launch_count(void) { return g_launches.load(); }

int mft_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_enabled.store(on != 0);
    g_used = 0;
    return MFT_OK;
}

int mft_set_pdl(int level) {
    int before = pdl_level();
    set_pdl_level(level);
    return before;
}

int mft_prof_categories(void) { return PC_COUNT; }
const char* mft_prof_name(int cat) { return prof_name(cat); }

// Synchronises the device, then sums the elapsed time of every recorded scope per category.
int mft_prof_collect(float* ms, int* counts, int n) {
    MFT_CHECK_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < n; ++i) { ms[i] = 0.f; counts[i] = 0; }
    for (size_t i = 0; i < g_used; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_slots[i].a, g_slots[i].b) != cudaSuccess) continue;
        int c = g_slots[i].cat;
        if (c >= 0 && c < n) { ms[c] += t; counts[c] += 1; }
    }
    g_used = 0;
    return MFT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------- side branches
namespace mft {

namespace {
struct SidePool {
    bool ready = false;
    cudaStream_t s[kSideStreams];
    cudaEvent_t main_ev[kSideStreams];   // recorded on the main stream, waited on by side i
    cudaEvent_t side_ev[kSideStreams];   // recorded on side i, waited on by the main stream
};
SidePool g_pools[64];
std::mutex g_pool_mu;

SidePool* pool_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    SidePool& p = g_pools[dev];
    if (!p.ready) {
        for (int i = 0; i < kSideStreams; ++i) {
            if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&p.main_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&p.side_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        p.ready = true;
    }
    return &p;
}
}  // namespace

Branches::Branches(cudaStream_t main) : main_(main), pool_(pool_for_current_device()), ok_(true) {
    for (int i = 0; i < kSideStreams; ++i) forked_[i] = false;
    if (pool_ == nullptr) ok_ = false;
}

void Branches::sync_to_main(int i) {
    SidePool* p = static_cast<SidePool*>(pool_);
    if (!p) return;
    if (cudaEventRecord(p->main_ev[i], main_) != cudaSuccess) ok_ = false;
    if (cudaStreamWaitEvent(p->s[i], p->main_ev[i], 0) != cudaSuccess) ok_ = false;
}

cudaStream_t Branches::fork(int i) {
    SidePool* p = static_cast<SidePool*>(pool_);
    if (!p) return main_;               // no pool: everything stays on the caller's stream
    sync_to_main(i);
    forked_[i] = true;
    return p->s[i];
}

cudaStream_t Branches::stream(int i) const {
    SidePool* p = static_cast<SidePool*>(pool_);
    return p ? p->s[i] : main_;
}

void Branches::join(int i) {
    SidePool* p = static_cast<SidePool*>(pool_);
    if (!p || !forked_[i]) return;
    if (cudaEventRecord(p->side_ev[i], p->s[i]) != cudaSuccess) ok_ = false;
    if (cudaStreamWaitEvent(main_, p->side_ev[i], 0) != cudaSuccess) ok_ = false;
    forked_[i] = false;
}

Branches::~Branches() {
    for (int i = 0; i < kSideStreams; ++i) join(i);
}

}  // namespace mft
