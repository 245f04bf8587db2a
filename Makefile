# Top-level build: the CUDA engine (C ABI), the host CLI and the checkers under oracle/.
#   make            -> pangenomesim_b200/lib/libpgs.so, pangenomesim_b200/bin/pangenomesim-b200, oracle
#   make lib|cli|oracle|clean
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-ffp-contract=off -Xptxas -v
CSRC     := pangenomesim_b200/csrc
HOST     := pangenomesim_b200/host
LIBDIR   := pangenomesim_b200/lib
BINDIR   := pangenomesim_b200/bin
OBJDIR   := build/obj

LIB_SRCS := $(CSRC)/engine.cu $(CSRC)/plan.cu $(CSRC)/kernels.cu $(CSRC)/mismatch.cu $(CSRC)/mismatch_tc.cu $(CSRC)/format.cu $(CSRC)/edge_table.cpp $(CSRC)/expand.cpp
LIB_HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/pgs.h
LIB_OBJS := $(OBJDIR)/engine.o $(OBJDIR)/plan.o $(OBJDIR)/kernels.o $(OBJDIR)/mismatch.o $(OBJDIR)/mismatch_tc.o $(OBJDIR)/format.o $(OBJDIR)/edge_table.o $(OBJDIR)/expand.o

all: lib cli oracle

lib: $(LIBDIR)/libpgs.so
cli: $(BINDIR)/pangenomesim-b200 $(LIBDIR)/libpgshost.so
oracle:
	$(MAKE) -C oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(LIB_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)
	@grep -E "error|warning|spill|registers" $(OBJDIR)/$*.ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(LIB_HDRS)
	@mkdir -p $(OBJDIR)
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off -c $< -o $@

# the host-side base expander is a plain byte loop with SIMD variants: optimise it fully
$(OBJDIR)/expand.o: $(CSRC)/expand.cpp $(CSRC)/expand.h
	@mkdir -p $(OBJDIR)
	$(CXX) -O3 -std=c++17 -fPIC -Wall -Wextra -c $< -o $@

$(LIBDIR)/libpgs.so: $(LIB_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(LIB_OBJS) -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64 -ldl

HOST_HDRS := $(wildcard $(HOST)/*.h) include/pgs.h
$(BINDIR)/pangenomesim-b200: $(HOST)/main.cpp $(HOST)/model.cpp $(HOST_HDRS) $(LIBDIR)/libpgs.so
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude $(HOST)/main.cpp $(HOST)/model.cpp -o $@ -L$(LIBDIR) -lpgs \
		-Wl,-rpath,'$$ORIGIN/../lib' -lpthread

# the host model as a small shared library for Python callers (bench.py, tests): no CUDA inside
$(LIBDIR)/libpgshost.so: $(HOST)/capi_host.cpp $(HOST)/model.cpp $(HOST_HDRS)
	@mkdir -p $(LIBDIR)
	$(CXX) -O2 -std=c++17 -fPIC -shared -Wall -Wextra -Iinclude $(HOST)/capi_host.cpp $(HOST)/model.cpp -o $@

clean:
	rm -rf build $(LIBDIR) $(BINDIR)
	$(MAKE) -C oracle clean
.PHONY: all lib cli oracle clean
