// Arguments.h -- the reference CLI contract (Arguments.h:66-86, Meta.h:47-67, util/CommandLine.h:52-86):
// same 14 flags, same defaults, same validation rule (invalid -> "invalid arguments" + usage +
// exit(-1)), same echo lines.  Flag lookup is a positional string match whose value is the next
// argv entry, parsed with atoi / sscanf("%lf") like the reference.  Settings live in one struct
// instead of mutable globals.  Extension flags (ignored by the reference) are long options.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

namespace dppr_host {

constexpr double kAlpha = 0.15;  // Meta.h:31

struct Settings {
    std::string data_file;        // -d
    int app_type = 0;             // -a  (0 = reverse push, the only algorithm)
    int directed = -1;            // -i
    int dynamic = -1;             // -y
    double window_ratio = 0.1;    // -w
    int workload_type = 0;        // -n  0 = batch as ratio of window, 1 = batch as edge count
    double batch_ratio = -1.0;    // -r
    long batch_count = 0;         // -b
    long per_batch = 0;           // -c
    long total_edges = 0;         // -l
    int source = 1;               // -s
    int threads = 1;              // -t  (CPU build only; accepted and ignored)
    int variant = 0;              // -o
    double tolerance = 1e-9;      // -e
    // extensions
    std::string sources_file;     // --sources <file>: one vertex id per line (multi-source run)
    int device = 0;               // --device <ordinal>
    int stepwise = 0;             // --stepwise 1: one launch per push iteration
    std::string dump_file;        // --dump <file>: final estimates of every source, raw float64
    int stats = 0;                // --stats 1: per-batch counters on stdout
    double pool_factor = 0.0;     // --pool-factor <x>: adjacency pool slots per window CSR entry (0 = library default, 8)
    int validate = 0;             // --validate 1: the reference's -DVALIDATE checks after the initial solve and after every batch, on
                                  //   the device (window graph vs the file's window, residual bound, push invariant); exit(-1) on a violation
    std::string pick;             // --pick top10|top1000|top1000000: choose the source like the reference's `workload` tool does
                                  //   (workload/Workload.cpp:47-55; 4th id of the bucket, as scripts/gpu.sh:17 does) instead of -s
    int progress = 1;             // --progress 0: print the keys at the end only (the default follows the reference's cadence,
                                  //   gpu/PPRGPU.cuh:116: before every batch when a batch holds more than 100 edges, else every 100)
};

class ArgScanner {
public:
    ArgScanner(int argc, char **argv) : argc_(argc), argv_(argv) {}
    const char *find(const std::string &flag) const {
        for (int i = 1; i < argc_; ++i)
            if (flag == argv_[i]) return (i + 1 < argc_) ? argv_[i + 1] : "";
        return nullptr;
    }
    std::string str(const std::string &flag, const std::string &dflt) const {
        const char *v = find(flag);
        return v ? std::string(v) : dflt;
    }
    long integer(const std::string &flag, long dflt) const {
        const char *v = find(flag);
        return v ? std::atol(v) : dflt;
    }
    double real(const std::string &flag, double dflt) const {
        const char *v = find(flag);
        if (!v) return dflt;
        double x = dflt;
        if (std::sscanf(v, "%lf", &x) == EOF) {
            std::cout << "usage: " << argv_[0] << " bad argument" << std::endl;
            std::abort();
        }
        return x;
    }
private:
    int argc_;
    char **argv_;
};

inline void print_usage() {
    std::cout << "==========[USAGE]==========" << std::endl;
    std::cout << "-d: gDataFileName" << std::endl;
    std::cout << "-a: gAppType" << std::endl;
    std::cout << 0 << ":rev push" << std::endl;
    std::cout << "-i: gIsDirected" << std::endl;
    std::cout << "-y: gIsDynamic" << std::endl;
    std::cout << "-w: gWindowRatio" << std::endl;
    std::cout << "-n: gWorkloadConfigType" << std::endl;
    std::cout << 0 << ": SLIDE_WINDOW_RATIO, " << 1 << ": SLIDE_BATCH_SIZE" << std::endl;
    std::cout << "-r: gStreamUpdateCountVersusWindowRatio" << std::endl;
    std::cout << "-b: gStreamBatchCount" << std::endl;
    std::cout << "-c: gStreamUpdateCountPerBatch" << std::endl;
    std::cout << "-l: gStreamUpdateCountTotal" << std::endl;
    std::cout << "-s: gSourceVertexId" << std::endl;
    std::cout << "-t: gThreadNum" << std::endl;
    std::cout << "-o: gVariant" << std::endl;
    std::cout << 0 << ": optimized, " << 1 << ": fast frontier, " << 2 << ": eager, " << 3 << ": VANILLA" << std::endl;
    std::cout << "-e: error tolerance" << std::endl;
    std::cout << "extensions: --sources <file> --device <n> --stepwise <0|1> --dump <file> --stats <0|1> --pool-factor <x>" << std::endl;
    std::cout << "            --validate <0|1> --pick <top10|top1000|top1000000> --progress <0|1>" << std::endl;
    std::cout << "EXAMPLE: ./pagerank -d ../data/com-dblp.ungraph.bin -a 0 -i 0 -y 1 -w 0.1 -n 0 -r 0.01 -b 1000 -s 1" << std::endl;
    std::cout << "EXAMPLE: ./pagerank -d ../data/com-dblp.ungraph.bin -a 0 -i 0 -y 1 -w 0.1 -n 1 -c 100 -l 10000 -s 1" << std::endl;
}

// the reference's validity rule (Arguments.h:42-64), plus the variant range the reference leaves unchecked
inline bool settings_valid(const Settings &s) {
    if (s.app_type < 0 || s.app_type > 1) return false;
    if (s.directed < 0 || s.dynamic < 0 || s.data_file.empty()) return false;
    if (s.workload_type == 0) {
        if (s.batch_ratio < 0.0 || s.batch_count == 0) return false;
    } else if (s.workload_type == 1) {
        if (s.per_batch == 0 || s.total_edges == 0) return false;
    } else {
        return false;
    }
    if (s.variant < 0 || s.variant > 3) return false;
    if (!s.pick.empty() && s.pick != "top10" && s.pick != "top1000" && s.pick != "top1000000") return false;
    return true;
}

inline Settings parse_arguments(int argc, char **argv) {
    ArgScanner a(argc, argv);
    Settings s;
    s.data_file = a.str("-d", "");
    s.app_type = (int)a.integer("-a", 0);
    s.directed = (int)a.integer("-i", -1);
    s.dynamic = (int)a.integer("-y", -1);
    s.window_ratio = a.real("-w", 0.1);
    s.workload_type = (int)a.integer("-n", 0);
    s.batch_ratio = a.real("-r", -1.0);
    s.batch_count = a.integer("-b", 0);
    s.per_batch = a.integer("-c", 0);
    s.total_edges = a.integer("-l", 0);
    s.source = (int)a.integer("-s", 1);
    s.threads = (int)a.integer("-t", 1);
    s.variant = (int)a.integer("-o", 0);
    s.tolerance = a.real("-e", 0.000000001);
    s.sources_file = a.str("--sources", "");
    s.device = (int)a.integer("--device", 0);
    s.stepwise = (int)a.integer("--stepwise", 0);
    s.dump_file = a.str("--dump", "");
    s.stats = (int)a.integer("--stats", 0);
    s.pool_factor = a.real("--pool-factor", 0.0);
    s.validate = (int)a.integer("--validate", 0);
    s.pick = a.str("--pick", "");
    s.progress = (int)a.integer("--progress", 1);
    if (!settings_valid(s)) {
        std::cout << "invalid arguments" << std::endl;
        print_usage();
        std::exit(-1);
    }
    return s;
}

// same echo as the reference's PrintArguments (Arguments.h:34-40); per-batch values are still 0 here,
// exactly as in the reference, because they are derived when the stream is opened
inline void print_arguments(const Settings &s) {
    std::cout << "gAppType=" << s.app_type << ",gIsDirected=" << s.directed << ",gIsDynamic=" << s.dynamic << std::endl;
    std::cout << "gWindowRatio=" << s.window_ratio << ",gWorkloadConfigType=" << s.workload_type
              << ",gStreamUpdateCountVersusWindowRatio=" << s.batch_ratio << ",gStreamBatchCount=" << s.batch_count
              << ",gStreamUpdateCountPerBatch=" << s.per_batch << ",gStreamUpdateCountTotal=" << s.total_edges << std::endl;
    std::cout << "gSourceVertexId=" << s.source << std::endl;
    std::cout << "gThreadNum=" << s.threads << ",gVariant=" << s.variant << std::endl;
    std::cout << "error=" << s.tolerance << ",ALPHA=" << kAlpha << std::endl;
}

}  // namespace dppr_host
