// Process-wide count of the kernels this library launched (C ABI: hint_launch_count).  bench.py reports the difference over
// its timed region as "gpu_launches": a count, not a formula.
#pragma once
#include <atomic>

namespace hint {
extern std::atomic<unsigned long long> g_launches;
}
#define HINT_LAUNCHED() ::hint::g_launches.fetch_add(1ull, std::memory_order_relaxed)
