// support.cu -- error string + launch counter shared by every launcher.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"

namespace itb {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = std::getenv("ITB_NO_PDL");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
int *stream_tickets(cudaStream_t st, int n) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, int *> table;
    constexpr int kPool = 65536;
    if (n > kPool) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int *&p = table[{dev, st}];
    if (!p) {
        if (cudaMalloc(&p, kPool * sizeof(int)) != cudaSuccess) return p = nullptr;
        if (cudaMemset(p, 0, kPool * sizeof(int)) != cudaSuccess) return nullptr;
    }
    return p;
}
long long launches() { return g_launches.load(std::memory_order_relaxed); }

// L2 prefetch hint: "the kernel launched AFTER the next one will stream these bytes" -- consumed (and cleared) by the next launcher
// that knows how to use it (the decode attention kernel; measured slower than no hint, see below -- opt-in)
static thread_local const void *g_pf_ptr = nullptr;
static thread_local long long g_pf_bytes = 0;
void take_prefetch_hint(const void *&ptr, long long &bytes) {
    // MEASURED (B200, C3 decode step, 20 replays): with the hint 3.735 ms / step, without 3.607 -- the attention kernel is latency-
    // sensitive, the extra 33.5 MB per layer slow it by more than the o-projection gains from L2.  Opt-in: ITB_L2_PREFETCH=1
    static const bool off = [] {
        const char *e = std::getenv("ITB_L2_PREFETCH");
        return !(e && e[0] == '1');
    }();
    ptr = off ? nullptr : g_pf_ptr;
    bytes = off ? 0 : g_pf_bytes;
    g_pf_ptr = nullptr;
    g_pf_bytes = 0;
}
}  // namespace itb

extern "C" void it_b200_l2_prefetch_hint(const void *ptr, long long bytes) {
    itb::g_pf_ptr = ptr;
    itb::g_pf_bytes = ptr && bytes > 0 ? bytes : 0;
}

extern "C" const char *it_b200_last_error(void) { return itb::g_err; }
extern "C" int it_b200_version(void) { return 1; }
extern "C" long long it_b200_launch_count(void) { return itb::launches(); }
