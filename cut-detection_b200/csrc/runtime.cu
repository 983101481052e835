// Error plumbing, device checks, launch accounting and the per-kernel event profiler of the C ABI.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace cutdet {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_error, sizeof(g_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return CUTDET_ECUDA;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

namespace {
std::atomic<long long> g_launches{0};
std::atomic<bool> g_profiling{false};
std::mutex g_prof_mutex;
struct ProfRecord { const char *name; cudaEvent_t start, stop; };
std::vector<ProfRecord> g_records;
}  // namespace

KernelScope::KernelScope(const char *name, cudaStream_t stream) : slot_(-1), stream_(stream) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    ProfRecord r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, stream);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    slot_ = (int)g_records.size();
    g_records.push_back(r);
}

KernelScope::~KernelScope() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    cudaEventRecord(g_records[slot_].stop, stream_);
}

}  // namespace cutdet

extern "C" long long cutdet_launch_count(void) { return cutdet::g_launches.load(); }

extern "C" int cutdet_profile_begin(void) {
    std::lock_guard<std::mutex> lock(cutdet::g_prof_mutex);
    for (auto &r : cutdet::g_records) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    cutdet::g_records.clear();
    cutdet::g_profiling.store(true);
    return CUTDET_OK;
}

extern "C" int cutdet_profile_end(char *json_out, size_t capacity) {
    using namespace cutdet;
    g_profiling.store(false);
    CUTDET_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto &r : g_records) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            auto &a = agg[r.name];
            a.first += 1;
            a.second += ms;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_records.clear();
    std::string out = "{";
    bool first = true;
    for (auto &kv : agg) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
                 kv.second.first, kv.second.second);
        out += buf;
        first = false;
    }
    out += "}";
    if (!json_out || out.size() + 1 > capacity) return fail(CUTDET_ECAPACITY, "profile_end: output buffer too small");
    memcpy(json_out, out.c_str(), out.size() + 1);
    return CUTDET_OK;
}

extern "C" int cutdet_abi_version(void) { return CUTDET_ABI_VERSION; }

extern "C" const char *cutdet_last_error(void) { return cutdet::g_error; }

extern "C" int cutdet_device_check(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    CUTDET_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0, sms = 0;
    CUTDET_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUTDET_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    CUTDET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (major != 10)
        return cutdet::fail(CUTDET_EUNSUPPORTED,
                            "libcutdet_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return CUTDET_OK;
}

extern "C" int cutdet_target_size(int width, int height, int resize, int *new_width, int *new_height) {
    CUTDET_REQUIRE(width > 0 && height > 0 && resize > 0 && new_width && new_height, "cutdet_target_size: bad argument");
    // frameID/data.py:199-202: int(height * (new_width / width)) in Python floats (doubles)
    *new_width = resize;
    *new_height = (int)((double)height * ((double)resize / (double)width));
    return CUTDET_OK;
}
