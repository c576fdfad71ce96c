#include "session.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace spsp_host {

static const std::chrono::steady_clock::time_point g_loaded = std::chrono::steady_clock::now();
void trace(const char *what)
{
    static const bool on = getenv("SPSP_TRACE") != nullptr;
    if (!on) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_loaded).count();
    fprintf(stderr, "[spsp %9.1f ms] %s\n", ms, what);
}

void throw_spsp(const std::string &what) { throw std::runtime_error(what + ": " + spsp_last_error()); }

DeviceSession::DeviceSession(int device, int k, int m, uint64_t threshold, int n_slots)
    : device_(device), n_slots_(n_slots)
{
    if (spsp_create(device, k, m, threshold, n_slots, &ctx_) != 0) throw_spsp("spsp_create");
}

DeviceSession::~DeviceSession() { spsp_destroy(ctx_); }

uint64_t DeviceSession::launches() const
{
    uint64_t n = 0;
    spsp_launch_count(ctx_, &n);
    return n;
}

}  // namespace spsp_host
