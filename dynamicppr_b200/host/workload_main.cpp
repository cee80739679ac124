// workload -- drop-in for the reference's source-picking tool (workload/Workload.cpp:29-63):
//     ./workload filename directed is_window is_choose_outdegree
// reads an encoder .bin file, ranks the vertices by out- (or in-) degree over the whole file (or over its first window, 10 %
// of the stream: workload/Graph.h:28-84) and writes ten ids per bucket to <basename>_top[window][rev]{10,1000,1000000}.txt in
// the current directory -- same file names, same bucket rule, same draws (SourcePicker.h).  The degree histograms and the
// ranking run on the GPU (dppr_rank_by_degree), exact at Twitter scale; the reference reads the file four bytes at a time.
#include <fcntl.h>
#include <fstream>
#include <iostream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "SourcePicker.h"

int main(int argc, char *argv[]) {
    using namespace dppr_host;
    if (argc != 5) {
        std::cout << "./workload filename directed is_window is_choose_outdegree" << std::endl;
        return -1;
    }
    std::string input_filename(argv[1]);
    const int directed = std::atoi(argv[2]), is_window = std::atoi(argv[3]), is_out_degree = std::atoi(argv[4]);
    std::cout << "is_directed=" << directed << ",window=" << is_window << ",is_outdegree=" << is_out_degree << std::endl;
    try {
        const int fd = ::open(input_filename.c_str(), O_RDONLY);
        struct stat sb;
        if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size < 4 || (sb.st_size - 4) % 8 != 0)
            throw std::runtime_error(input_filename + " is not an encoder .bin file");
        void *map = ::mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) throw std::runtime_error("mmap failed");
        const int32_t V = *(const int32_t *)map;
        const int32_t *pairs = (const int32_t *)((const char *)map + 4);
        int64_t n = (int64_t)((sb.st_size - 4) / 8);
        std::cout << (is_window ? "read window graph filename=" : "readfile filename=") << input_filename << std::endl;
        std::cout << "vertex_count=" << V << std::endl;
        if (is_window) n = (int64_t)((double)n * 0.1);  // gWindowRatio's default (Meta.cpp); the tool takes no -w
        const DegreeRanking r = rank_by_degree(0, V, directed != 0, is_out_degree != 0, pairs, n);
        std::cout << "finish sort" << std::endl;
        const std::string base = input_filename.substr(input_filename.rfind('/') + 1);
        const std::string prefix = std::string("top") + (is_window ? "window" : "") + (is_out_degree ? "" : "rev");
        for (const char *bucket : {"10", "1000", "1000000"}) {
            const std::string fname = base + "_" + prefix + bucket + ".txt";
            std::cout << "top" << bucket << " filename=" << fname << std::endl;
            int64_t st, ed;
            bucket_ranks(std::string("top") + bucket, st, ed);
            const std::vector<int32_t> ids = choose_degree_range(r, 10, st, ed);
            std::ofstream f(fname.c_str());
            for (int32_t u : ids) f << u << std::endl;
        }
        ::munmap(map, (size_t)sb.st_size);
        ::close(fd);
    } catch (const std::exception &e) {
        std::cout << "error: " << e.what() << std::endl;
        return -1;
    }
    return 0;
}
