// wire.cpp -- see wire.hpp
#include "wire.hpp"

#include <immintrin.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace sloth {

namespace {

void fill_plain(uint32_t* p, size_t n, uint32_t v)
{
    for (size_t i = 0; i < n; ++i) p[i] = v;
}

// 32-byte non-temporal stores for the aligned middle of a long run (no read-for-ownership of the destination lines)
__attribute__((target("avx2"))) void fill_stream_avx2(uint32_t* p, size_t n, uint32_t v)
{
    while (n && (reinterpret_cast<uintptr_t>(p) & 31u)) { *p++ = v; --n; }
    const __m256i x = _mm256_set1_epi32((int)v);
    size_t blocks = n / 8;
    for (; blocks >= 4; blocks -= 4, p += 32) {
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p), x);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + 8), x);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + 16), x);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + 24), x);
    }
    for (; blocks; --blocks, p += 8) _mm256_stream_si256(reinterpret_cast<__m256i*>(p), x);
    for (size_t i = 0; i < (n & 7u); ++i) p[i] = v;
}

bool have_avx2()
{
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
}

}  // namespace

void expand_runs(const Run* runs, size_t n_runs, size_t cell_begin, size_t cell_end, uint32_t* cells, size_t n_cells)
{
    cell_end = std::min(cell_end, n_cells);
    if (n_runs == 0 || cell_begin >= cell_end) return;
    // the run that holds cell_begin: the last one whose start is <= cell_begin
    size_t lo = 0, hi = n_runs;
    while (hi - lo > 1) {
        const size_t mid = lo + (hi - lo) / 2;
        if (runs[mid].start <= cell_begin) lo = mid; else hi = mid;
    }
    const bool avx2 = have_avx2();
    bool streamed = false;
    for (size_t i = lo; i < n_runs && runs[i].start < cell_end; ++i) {
        const size_t a = std::max<size_t>(runs[i].start, cell_begin);
        const size_t b = std::min<size_t>(i + 1 < n_runs ? runs[i + 1].start : n_cells, cell_end);
        if (b <= a) continue;   // never for runs the device wrote; keeps a corrupt list inside the range
        const size_t n = b - a;
        if (avx2 && n >= 64) { fill_stream_avx2(cells + a, n, runs[i].cell); streamed = true; }
        else fill_plain(cells + a, n, runs[i].cell);
    }
    if (streamed) _mm_sfence();   // streaming stores are weakly ordered: fence before the range is handed over
}

struct WirePool::Impl {
    std::mutex mu;
    std::condition_variable cv_job, cv_idle;
    std::deque<std::function<void()>> jobs;
    size_t running = 0;
    bool stop = false;
    std::vector<std::thread> threads;
};

WirePool::WirePool(unsigned n_threads) : impl_(new Impl), n_threads_(std::max(1u, n_threads))
{
    for (unsigned t = 0; t < n_threads_; ++t)
        impl_->threads.emplace_back([this] {
            Impl& s = *impl_;
            std::unique_lock<std::mutex> lk(s.mu);
            for (;;) {
                s.cv_job.wait(lk, [&] { return s.stop || !s.jobs.empty(); });
                if (s.jobs.empty()) return;   // stop, and nothing left to do
                std::function<void()> job = std::move(s.jobs.front());
                s.jobs.pop_front();
                ++s.running;
                lk.unlock();
                job();
                lk.lock();
                --s.running;
                if (s.jobs.empty() && s.running == 0) s.cv_idle.notify_all();
            }
        });
}

WirePool::~WirePool()
{
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_job.notify_all();
    for (std::thread& t : impl_->threads) t.join();
    delete impl_;
}

void WirePool::submit(std::function<void()> job)
{
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->jobs.push_back(std::move(job));
    }
    impl_->cv_job.notify_one();
}

void WirePool::wait_idle()
{
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->cv_idle.wait(lk, [&] { return impl_->jobs.empty() && impl_->running == 0; });
}

unsigned wire_default_threads()
{
    if (const char* e = std::getenv("SLOTH_WIRE_THREADS")) return (unsigned)std::min(64, std::max(1, std::atoi(e)));
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 8;
    unsigned ranks = 1;
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, std::atoi(e));
    return std::min(32u, std::max(2u, hw / ranks));
}

}  // namespace sloth
