// SourcePicker.h -- the reference's source-selection rule (workload/Graph.h:178-227, workload/Workload.cpp:29-63) on top
// of the device degree ranking of include/dppr.h (dppr_rank_by_degree): SURVEY 8f row f2.
//
// Buckets: "top10" = exactly the ranks 0..9 by degree; "top1000" = 10 distinct vertices drawn from ranks [10, 1000);
// "top1000000" = 10 distinct vertices drawn from ranks [1000, 10^6) (clipped to V); a drawn vertex must have both an out-
// and an in-edge (workload/Graph.h:211-214).  Draws use rand() % (rank_ed - rank_st) + rank_st with the C library's default
// seed, exactly like the reference, so on a degree sequence without ties inside the bucket the two tools write the same ids
// (the reference's std::sort leaves the order of equal degrees unspecified; here ties go by ascending id).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
#include "dppr.h"

namespace dppr_host {

struct DegreeRanking {
    std::vector<int32_t> order, out_degree, in_degree;
};

// pairs: n (int32, int32) records in host memory (e.g. the mmap'ed payload of a .bin file)
inline DegreeRanking rank_by_degree(int device, int32_t V, bool directed, bool by_out_degree, const int32_t *pairs, int64_t n) {
    DegreeRanking r;
    r.order.resize((size_t)V); r.out_degree.resize((size_t)V); r.in_degree.resize((size_t)V);
    const int rc = dppr_rank_by_degree(device, V, directed ? 1 : 0, by_out_degree ? 1 : 0, pairs, n, 0, r.order.data(),
                                       r.out_degree.data(), r.in_degree.data());
    if (rc != DPPR_OK) throw std::runtime_error(std::string("dppr_rank_by_degree failed: ") + dppr_last_error(nullptr));
    return r;
}

// ChooseVertexDegreeRange (workload/Graph.h:178-227), including its progress line per chosen vertex
inline std::vector<int32_t> choose_degree_range(const DegreeRanking &r, size_t num, int64_t rank_st, int64_t rank_ed, bool verbose = true) {
    const int64_t V = (int64_t)r.order.size();
    if (rank_ed > V) rank_ed = V;
    if (rank_ed - rank_st < (int64_t)num) throw std::runtime_error("degree bucket holds fewer vertices than requested");
    std::vector<int32_t> ids(num);
    if (rank_ed - rank_st == (int64_t)num) {
        for (size_t i = 0; i < num; ++i) ids[i] = r.order[(size_t)(rank_st + (int64_t)i)];
    } else {
        // (the reference loops for ever if the bucket holds fewer than `num` connected vertices; bail out instead)
        for (size_t i = 0; i < num; ++i) {
            for (long tries = 0;; ++tries) {
                if (tries > 100000000L) throw std::runtime_error("degree bucket holds too few connected vertices");
                const int64_t p = std::rand() % (rank_ed - rank_st) + rank_st;
                const int32_t u = r.order[(size_t)p];
                bool valid = true;
                for (size_t j = 0; j < i; ++j)
                    if (ids[j] == u) { valid = false; break; }
                if (r.out_degree[(size_t)u] == 0 || r.in_degree[(size_t)u] == 0) valid = false;
                if (valid) { ids[i] = u; break; }
            }
        }
    }
    if (verbose)
        for (size_t i = 0; i < num; ++i)
            std::cout << "u=" << ids[i] << ",deg=" << r.out_degree[(size_t)ids[i]] << ",in_deg=" << r.in_degree[(size_t)ids[i]] << std::endl;
    return ids;
}

inline void bucket_ranks(const std::string &bucket, int64_t &rank_st, int64_t &rank_ed) {
    if (bucket == "top10") { rank_st = 0; rank_ed = 10; }
    else if (bucket == "top1000") { rank_st = 10; rank_ed = 1000; }
    else if (bucket == "top1000000") { rank_st = 1000; rank_ed = 1000000; }
    else throw std::runtime_error("unknown source bucket " + bucket);
}

}  // namespace dppr_host
