// wire.hpp -- host side of the span wire format (csrc/spans.cuh): rebuilds the 4-byte cells of a frame
// (Context.frame_buffer, context.rs:16) from its (start, cell) runs, on a small pool of worker threads so that the
// thread driving the GPU never touches the 33 MB of a 4K frame itself.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace sloth {

struct Run {
    uint32_t start;   // first cell of the run
    uint32_t cell;    // glyph | r << 8 | g << 16 | b << 24
};

// Rebuilds cells [cell_begin, cell_end) of a frame with n_cells cells from its n_runs runs (ascending starts, the first
// at 0; run i covers [start_i, start_{i+1}), the last one ends at n_cells).  Streaming stores for long runs: the
// destination is written once and not read back.
void expand_runs(const Run* runs, size_t n_runs, size_t cell_begin, size_t cell_end, uint32_t* cells, size_t n_cells);

// Fixed pool of worker threads with one FIFO of jobs.
class WirePool {
public:
    explicit WirePool(unsigned n_threads);
    ~WirePool();
    unsigned size() const { return n_threads_; }
    void submit(std::function<void()> job);
    void wait_idle();   // until every submitted job has finished
private:
    struct Impl;
    Impl* impl_;
    unsigned n_threads_;
};

// threads a context should use: SLOTH_WIRE_THREADS, else the hardware threads divided among the local ranks
// (LOCAL_WORLD_SIZE, set by torchrun), at least 2 and at most 32
unsigned wire_default_threads();

}  // namespace sloth
