# Builds the product: libdppr.so (hand-written sm_100a kernels behind the C ABI of include/dppr.h)
# and the `pagerank` CLI host.  The checker lives in oracle/ and has its own Makefile.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  := /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -ccbin $(CXX)
CSRC := dynamicppr_b200/csrc
HOST := dynamicppr_b200/host
LIBDIR := dynamicppr_b200/lib
BINDIR := dynamicppr_b200/bin
HDRS := $(wildcard $(CSRC)/*.cuh) include/dppr.h

.PHONY: all lib cli clean oracle
all: lib cli

lib: $(LIBDIR)/libdppr.so
$(LIBDIR)/libdppr.so: $(CSRC)/engine.cu $(CSRC)/capi.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) $(PTXAS_V) $(DPPR_DEFS) -shared -o $@ $(CSRC)/capi.cu -lcudart

cli: $(BINDIR)/pagerank $(BINDIR)/workload
$(BINDIR)/pagerank: $(HOST)/main.cpp $(wildcard $(HOST)/*.h) include/dppr.h $(LIBDIR)/libdppr.so
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -I$(HOST) -o $@ $(HOST)/main.cpp -L$(LIBDIR) -ldppr -Wl,-rpath,'$$ORIGIN/../lib' -lpthread
# drop-in for the reference's source picker (workload/Workload.cpp)
$(BINDIR)/workload: $(HOST)/workload_main.cpp $(HOST)/SourcePicker.h include/dppr.h $(LIBDIR)/libdppr.so
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -I$(HOST) -o $@ $(HOST)/workload_main.cpp -L$(LIBDIR) -ldppr -Wl,-rpath,'$$ORIGIN/../lib' -lpthread

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIBDIR) $(BINDIR)
