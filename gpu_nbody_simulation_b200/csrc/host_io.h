// Host-side helpers of libbh.so: the reference's text formats and the canonical node table.
#pragma once
#include <stdint.h>

#include <vector>

namespace bh {

struct HostTree {
    int finest = 0;
    uint64_t level_off[17] = {};
    double bounds[4] = {};
    std::vector<double> mass, comx, comy;
    std::vector<uint32_t> count, first;
    std::vector<uint32_t> sidx;   // body (internal index) per sorted position
    std::vector<uint32_t> perm;   // internal -> original index (empty = identity)
    std::vector<double> pos;      // internal order, xy interleaved
};

// DFS pre-order, children 0->3; row = {depth, xmin, xmax, ymin, ymax, mass, comx, comy, occupant, internal}.
// Writes at most cap_rows rows; returns the total number of nodes.
int64_t canonical_rows(const HostTree& t, double* out_rows, int64_t cap_rows);
int dump_quadtree_txt(const HostTree& t, const char* path);
int load_text(const char* mf, const char* pf, const char* vf, int64_t n, double* mass, double* pos, double* vel);
int append_positions_txt(const char* path, const double* pos, int64_t n, double time, int truncate);
int write_init_files(const char* mf, const char* pf, const char* vf, int64_t n, const double* mass, const double* pos,
                     const double* vel);
double round6(double v);   // value after a round trip through the "%.6g" text format

// Trajectory file (savePositions format, project.cu:855-863) written by a background thread: the producer hands over
// frames (n bodies in a pinned host buffer, a time stamp, the CUDA event recorded after the copy into the buffer); the
// thread waits for the event, formats the frame and appends it to ONE open file while the producer keeps stepping.
class FrameWriter {
public:
    FrameWriter(const char* path, int64_t n, int device);
    ~FrameWriter();                        // flushes, joins
    bool ok() const { return f_ != nullptr && !failed_; }
    // blocks until buffer `slot` (0 / 1) is no longer being written out
    void acquire(int slot);
    // queue buffer `slot` = host positions [n][2]; `ready` is a cudaEvent_t recorded after the copy into it
    void submit(int slot, const double* host, double time, void* ready_event);
    void finish();                         // wait until everything queued is on disk
private:
    struct Impl;
    Impl* impl_;
    void* f_;
    bool failed_ = false;
};

}  // namespace bh
