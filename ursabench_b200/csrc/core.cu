// Error plumbing, device info and ABI version of libursa_b200.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ursa {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return URSA_ERR_CUDA;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-kernel event timing ------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int kind; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;          // records of the current session
static std::vector<cudaEvent_t> g_prof_pool; // recycled events
static size_t g_prof_cap = 0;

bool prof_on() { return g_prof_on.load(std::memory_order_relaxed); }

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void prof_mark(int kind, cudaStream_t st, bool begin) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on.load()) return;
    if (begin) {
        if (g_prof.size() >= g_prof_cap) return;
        ProfRec r{prof_event(), nullptr, kind};
        if (r.a) cudaEventRecord(r.a, st);
        g_prof.push_back(r);
    } else {
        for (size_t i = g_prof.size(); i-- > 0;)
            if (g_prof[i].kind == kind && g_prof[i].b == nullptr) {
                g_prof[i].b = prof_event();
                if (g_prof[i].b) cudaEventRecord(g_prof[i].b, st);
                break;
            }
    }
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace ursa

extern "C" int ursa_abi_version(void) { return URSA_ABI_VERSION; }

extern "C" const char *ursa_last_error(void) { return ursa::g_err; }

extern "C" int ursa_profile_begin(int capacity) {
    std::lock_guard<std::mutex> lk(ursa::g_prof_mu);
    URSA_REQUIRE(capacity > 0, "ursa_profile_begin: capacity must be positive");
    for (auto &r : ursa::g_prof) { if (r.a) ursa::g_prof_pool.push_back(r.a); if (r.b) ursa::g_prof_pool.push_back(r.b); }
    ursa::g_prof.clear();
    ursa::g_prof_cap = (size_t)capacity;
    ursa::g_prof_on.store(true);
    return URSA_OK;
}

extern "C" int ursa_profile_end(double *ms_sum, int64_t *count, int n_kinds) {
    URSA_REQUIRE(ms_sum && count && n_kinds > 0, "ursa_profile_end: null pointer");
    ursa::g_prof_on.store(false);
    std::lock_guard<std::mutex> lk(ursa::g_prof_mu);
    for (int i = 0; i < n_kinds; ++i) { ms_sum[i] = 0.0; count[i] = 0; }
    for (auto &r : ursa::g_prof) {
        if (r.a && r.b && r.kind >= 0 && r.kind < n_kinds) {
            URSA_CUDA(cudaEventSynchronize(r.b));
            float ms = 0.f;
            URSA_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
            ms_sum[r.kind] += ms;
            count[r.kind] += 1;
        }
        if (r.a) ursa::g_prof_pool.push_back(r.a);
        if (r.b) ursa::g_prof_pool.push_back(r.b);
    }
    ursa::g_prof.clear();
    return URSA_OK;
}

extern "C" uint64_t ursa_launch_count(void) { return ursa::g_launches.load(std::memory_order_relaxed); }

extern "C" int ursa_device_info(int *sm, int *major, int *minor) {
    int dev = 0;
    URSA_CUDA(cudaGetDevice(&dev));
    if (sm) URSA_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (major) URSA_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    if (minor) URSA_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return URSA_OK;
}
