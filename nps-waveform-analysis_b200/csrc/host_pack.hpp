// Host-side lossless transport packing for the binary64 host layout (npswf_analyze_batch).
//
// The reference's samples are ADC counts times ADCtomV = 1000/4096 (T2:357), so a trace of doubles normally holds
// 16-bit information per sample, and the host->device copy (950 400 B per event as binary64) is what bounds the
// end-to-end rate.  A small pool of host threads therefore rewrites each chunk as int16 counts into a pinned
// staging buffer -- but only if that is provably lossless: every sample x must satisfy
//     double(k) * lsb == x   with   k = round(x / lsb),  |k| <= 32767
// (the product is what the device computes from the counts, widen_counts_kernel).  One sample that fails (off the
// lattice, out of range, NaN, infinity) sends the whole chunk over as the caller's doubles instead.  The kernels see
// the same real numbers either way; nothing is computed on the host.  The one bit pattern that is not preserved is
// the sign of a zero sample (-0.0 arrives as +0.0).  No output can tell: the pipeline only compares samples, subtracts
// them, takes |.| of them and adds their products into sums that start at +0.0 -- operations in which a zero's sign
// never reaches a non-zero result -- and it never divides by a sample (tests: the synthetic sets are full of -0.0).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <sched.h>
#include <pthread.h>

namespace npswf {

// Restrict the calling thread to `cpus` (the cores of a GPU's NUMA node); an empty list is a no-op.
inline void bind_this_thread(const std::vector<int> &cpus)
{
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus)
        if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
    (void)pthread_setaffinity_np(pthread_self(), sizeof set, &set);
}

// counts[i] = round(x[i] / lsb) for i in [0, n); returns true iff every sample is reproduced bit for bit.
// (host_pack.cpp: AVX2 when the CPU has it, portable C++ otherwise)
bool pack_counts_range(const double *x, int16_t *out, size_t n, double lsb, double inv_lsb);

// Persistent fork-join pool: pack() splits one chunk over the workers and the calling thread.
class PackPool {
public:
    // cpus: the cores the workers are confined to (those next to the GPU the packed chunks go to); empty = anywhere
    explicit PackPool(int n_threads, std::vector<int> cpus = {}) : n_(n_threads < 1 ? 1 : n_threads), cpus_(std::move(cpus))
    {
        for (int t = 1; t < n_; t++) workers_.emplace_back([this, t] { bind_this_thread(cpus_); loop(t); });
    }
    ~PackPool()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    PackPool(const PackPool &) = delete;
    PackPool &operator=(const PackPool &) = delete;
    int threads() const { return n_; }

    bool pack(const double *x, int16_t *out, size_t n, double lsb)
    {
        x_ = x; out_ = out; count_ = n; lsb_ = lsb; inv_ = 1.0 / lsb;
        ok_.store(true, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> g(mu_);
            pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        run_part(0);
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
        return ok_.load(std::memory_order_relaxed);
    }

private:
    void run_part(int t)
    {
        // slices of whole 4 KB pages of the source, so that neighbouring threads do not share cache lines
        const size_t per = ((count_ + (size_t)n_ - 1) / (size_t)n_ + 511) & ~(size_t)511;
        const size_t lo = std::min(count_, per * (size_t)t), hi = std::min(count_, lo + per);
        // sub-blocks so that a failing chunk is abandoned early by everyone
        for (size_t a = lo; a < hi; a += 65536) {
            if (!ok_.load(std::memory_order_relaxed)) return;
            const size_t b = std::min(hi, a + 65536);
            if (!pack_counts_range(x_ + a, out_ + a, b - a, lsb_, inv_)) ok_.store(false, std::memory_order_relaxed);
        }
    }
    void loop(int t)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            run_part(t);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }

    const int n_;
    const std::vector<int> cpus_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    const double *x_ = nullptr;
    int16_t *out_ = nullptr;
    size_t count_ = 0;
    double lsb_ = 0, inv_ = 0;
    std::atomic<bool> ok_{true};
};

}  // namespace npswf
