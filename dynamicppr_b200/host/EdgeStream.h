// EdgeStream.h -- the host side of the stream: .bin reader + window/batch arithmetic.
//
// Mirrors what the hot path needs from SlidingGraphVec (SlidingGraphVec.h:28-96, 219-275):
//   * .bin format: int32 vertex_count, then (int32 v1, int32 v2) records in stream order; M is
//     derived from the file size (encoder/GraphEncoder.h:86-95).
//   * W = (int)(M * window_ratio); mode 0: B = (size_t)(ratio * W), total = B * count; mode 1:
//     count = ceil(total / B); total capped at M - W (count NOT recomputed); the loop additionally
//     stops when fewer than B edges remain (:221).  Same truncations, same printed lines.
//   * next_batch() hands out the B ARRIVING edges (the reference's `new_stream`); the expiring ones
//     are not re-read from the file -- the device keeps the window in arrival order.
// The file is memory-mapped instead of fread()-ing four bytes at a time; the reference's host
// adjacency vectors (col_ind / in_col_ind) are not built at all: the window graph lives on the GPU.
#pragma once
#include <cstdint>
#include <cstdio>
#include <fcntl.h>
#include <iostream>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "Arguments.h"

namespace dppr_host {

class EdgeStream {
public:
    EdgeStream(const Settings &s, bool whole_file_window) {
        std::cout << "read filename=" << s.data_file << std::endl;
        fd_ = ::open(s.data_file.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("cannot open " + s.data_file);
        struct stat sb;
        if (fstat(fd_, &sb) != 0 || sb.st_size < 4 || (sb.st_size - 4) % 8 != 0)
            throw std::runtime_error(s.data_file + " is not an encoder .bin file (int32 V + (int32,int32) records)");
        size_ = (size_t)sb.st_size;
        map_ = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (map_ == MAP_FAILED) throw std::runtime_error("mmap failed for " + s.data_file);
        vertex_count = *(const int32_t *)map_;
        pairs_ = (const int32_t *)((const char *)map_ + 4);
        stream_length = (size_ - 4) / 8;
        std::cout << "vertex_count=" << vertex_count << std::endl;
        directed = s.directed != 0;

        if (whole_file_window) {  // -y 0: static graph = every edge of the file (GraphVec.h:28-82)
            window = (int64_t)stream_length;
            per_batch = 0; batch_count = 0; total = 0;
            std::cout << "edge_stream_count=" << stream_length << std::endl;
            std::cout << "edge_count=" << (directed ? window : 2 * window) << std::endl;
            pos_ = (size_t)window;
            return;
        }
        window = (int32_t)((double)stream_length * s.window_ratio);
        if (s.workload_type == 0) {
            per_batch = (size_t)(s.batch_ratio * (double)window);
            batch_count = (size_t)s.batch_count;
            total = per_batch * batch_count;
        } else {
            per_batch = (size_t)s.per_batch;
            total = (size_t)s.total_edges;
            batch_count = (total + per_batch - 1) / per_batch;
        }
        if (total > stream_length - (size_t)window) total = stream_length - (size_t)window;
        std::cout << "after workload config: gStreamUpdateCountPerBatch=" << per_batch << ",gStreamBatchCount=" << batch_count
                  << ",gStreamUpdateCountTotal=" << total << std::endl;
        std::cout << "sliding window size=" << window << ",gStreamUpdateCountPerBatch=" << per_batch << std::endl;
        std::cout << "edge_count=" << (directed ? window : 2 * window) << std::endl;
        pos_ = (size_t)window;
    }
    ~EdgeStream() {
        if (map_ && map_ != MAP_FAILED) ::munmap(map_, size_);
        if (fd_ >= 0) ::close(fd_);
    }
    EdgeStream(const EdgeStream &) = delete;
    EdgeStream &operator=(const EdgeStream &) = delete;

    const int32_t *initial_window() const { return pairs_; }
    const int32_t *all_pairs() const { return pairs_; }
    // the W edges the window holds after `done` batches: stream records [done * B, done * B + W)
    const int32_t *window_after(size_t done) const { return pairs_ + 2 * done * per_batch; }
    // nullptr when fewer than per_batch edges remain (StreamUpdates returning true, SlidingGraphVec.h:221)
    const int32_t *next_batch() {
        if (per_batch == 0 || pos_ + per_batch > stream_length) return nullptr;
        const int32_t *p = pairs_ + 2 * pos_;
        pos_ += per_batch;
        return p;
    }

    int32_t vertex_count = 0;
    bool directed = true;
    size_t stream_length = 0;  // M
    int64_t window = 0;        // W
    size_t per_batch = 0;      // B   (gStreamUpdateCountPerBatch)
    size_t batch_count = 0;    //     (gStreamBatchCount)
    size_t total = 0;          //     (gStreamUpdateCountTotal)

private:
    int fd_ = -1;
    void *map_ = nullptr;
    size_t size_ = 0, pos_ = 0;
    const int32_t *pairs_ = nullptr;
};

}  // namespace dppr_host
