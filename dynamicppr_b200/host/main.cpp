// pagerank -- drop-in for the reference's gpu/pagerank executable (gpu/PPRGPUMain.cu:8-39):
// same flags, same .bin input, same stdout keys; the work happens in libdppr.so (sm_100a CUDA).
#include <iostream>
#include "Arguments.h"
#include "EdgeStream.h"
#include "PPRDriver.h"

int main(int argc, char *argv[]) {
    using namespace dppr_host;
    Settings s = parse_arguments(argc, argv);
    print_arguments(s);
    try {
        EdgeStream stream(s, /*whole_file_window=*/s.dynamic == 0);
        PPRDriver driver(s, stream);
        if (s.dynamic) driver.DynamicExecute();
        else driver.Execute();
    } catch (const std::exception &e) {
        std::cout << "error: " << e.what() << std::endl;
        return -1;
    }
    return 0;
}
