// Host-side owner of one device context (include/spsp.h) shared by the
// worker threads of a process: one slot (CUDA stream) per worker.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "spsp.h"

namespace spsp_host {

class DeviceSession {
public:
    // Throws std::runtime_error when no CUDA device is usable: there is no CPU path.
    DeviceSession(int device, int k, int m, uint64_t threshold, int n_slots);
    ~DeviceSession();
    DeviceSession(const DeviceSession &) = delete;
    DeviceSession &operator=(const DeviceSession &) = delete;
    spsp_ctx *ctx() const { return ctx_; }
    int device() const { return device_; }
    int n_slots() const { return n_slots_; }
    uint64_t launches() const;

private:
    spsp_ctx *ctx_ = nullptr;
    int device_, n_slots_;
};

[[noreturn]] void throw_spsp(const std::string &what);
// SPSP_TRACE=1: milliseconds since the library was loaded + a label, on stderr (where a CLI run spends its time).
void trace(const char *what);

}  // namespace spsp_host
