# Makefile — builds everything in-tree (the built .so files travel to the GPU box with gpurun).
#   make lib      -> spmv_test_b200/lib/libspmv_b200.so      (the product: C-ABI + sm_100a kernels)
#   make oracle   -> oracle/liboracle.so (+ oracle/_ref/* when /root/reference is present)
#   make harness  -> build/sparse_sgemv                      (drop-in for the reference's executable)
CUDA   ?= /usr/local/cuda
NVCC   ?= $(CUDA)/bin/nvcc
CXX    ?= g++
ARCH   := -gencode arch=compute_100a,code=sm_100a
CSRC   := spmv_test_b200/csrc
HOST   := spmv_test_b200/host
LIBDIR := spmv_test_b200/lib
OBJ    := build/obj
NVFLAGS := -std=c++17 -O3 -lineinfo $(ARCH) -Iinclude -I$(CSRC) -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CXXFLAGS := -std=c++17 -O3 -fPIC -fvisibility=hidden -Iinclude -I$(CSRC) -I$(CUDA)/include

CU_SRCS  := capi wsp asp panel panel_rs strips mg compact pack_dev
CPP_SRCS := pack_host

all: lib oracle harness

lib: $(LIBDIR)/libspmv_b200.so

$(OBJ)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(OBJ)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.hpp include/*.h)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c -o $@ $<

$(LIBDIR)/libspmv_b200.so: $(foreach f,$(CU_SRCS) $(CPP_SRCS),$(OBJ)/$(f).o)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared $(ARCH) -o $@ $^ -cudart static -Xlinker --exclude-libs,ALL

oracle:
	$(MAKE) -C oracle

harness: build/sparse_sgemv

HOST_SRCS := $(wildcard $(HOST)/*.cpp) test/main.cpp
build/sparse_sgemv: $(HOST_SRCS) $(wildcard $(HOST)/include/*.hpp) $(LIBDIR)/libspmv_b200.so
	@mkdir -p build
	$(CXX) -std=c++17 -O2 -I$(HOST)/include -Iinclude -I$(CUDA)/include -o $@ $(HOST_SRCS) -L$(LIBDIR) -lspmv_b200 \
	    -L$(CUDA)/lib64 -lcublas -lcudart -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -Wl,-rpath,$(CUDA)/lib64

clean:
	rm -rf build $(LIBDIR)/libspmv_b200.so
	$(MAKE) -C oracle clean

.PHONY: all lib oracle harness clean
