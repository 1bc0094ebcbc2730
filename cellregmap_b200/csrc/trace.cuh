// CRM_TRACE=1: per-phase device times (CUDA events on the launching stream) and host enqueue times, printed to stderr at the end of a call.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <utility>
#include <vector>

namespace crm {

// CRM_TRACE=1: per-phase device times (CUDA events on the launching stream) printed to stderr at the end of a call
static inline double host_ms() { static const auto t0 = std::chrono::steady_clock::now(); return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
static inline bool trace_on() { static const bool on = [] { const char* v = getenv("CRM_TRACE"); return v && atoi(v) != 0; }(); return on; }
struct PhaseTrace {
    bool on;
    cudaStream_t st;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    std::vector<double> host_at;     // host clock when each mark was enqueued: tells a slow phase (device) from a late one (host)
    explicit PhaseTrace(cudaStream_t s) : st(s) { static const bool env = [] { const char* v = getenv("CRM_TRACE"); return v && atoi(v) != 0; }(); on = env; mark("begin"); }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); marks.emplace_back(name, e); host_at.push_back(host_ms());
    }
    void report(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "[crm trace] %s:", what);
        for (size_t i = 1; i < marks.size(); i++) { float ms = 0.f; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second); fprintf(stderr, " %s %.2f ms (host +%.2f) |", marks[i].first, ms, host_at[i] - host_at[i - 1]); }
        fprintf(stderr, "\n");
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear(); host_at.clear();
    }
};


// CRM_TRACE=1: reports a host-side section that took longer than `limit_ms` (where does an occasional long step lose its time?)
struct SlowSection {
    const char* name; double t0, limit;
    explicit SlowSection(const char* n, double limit_ms = 3.0) : name(n), t0(trace_on() ? host_ms() : 0.0), limit(limit_ms) {}
    ~SlowSection() {
        if (!trace_on()) return;
        const double dt = host_ms() - t0;
        if (dt > limit) fprintf(stderr, "[crm trace] host %.1f ms: SLOW host section '%s' took %.1f ms\n", host_ms(), name, dt);
    }
};

}  // namespace crm
