# Builds libmdbg_b200.so (hand-written sm_100a kernels + C ABI) in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC := rust-mdbg_b200/csrc
OUT := rust-mdbg_b200/libmdbg_b200.so
CU := $(wildcard $(CSRC)/*.cu)
CPP := $(wildcard $(CSRC)/*.cpp)
CC := $(wildcard $(CSRC)/*.cc)
OBJ := $(CU:.cu=.o) $(CPP:.cpp=.o) $(CC:.cc=.o)
HDR := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

CLI := rust-mdbg_b200/rust-mdbg

all: $(OUT) $(CLI) oracle model

$(CLI): rust-mdbg_b200/cli/rust_mdbg_main.cpp $(wildcard rust-mdbg_b200/cli/*.hpp) include/mdbg.h $(OUT)
	g++ -O2 -std=c++17 -Wall -o $@ $< -Lrust-mdbg_b200 -lmdbg_b200 -lz -Wl,-rpath,'$$ORIGIN'

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(CSRC)/%.o: $(CSRC)/%.cpp $(HDR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

# host-only sources (SIMD intrinsics): plain g++; -mssse3 only where the host is x86
HOST_SIMD := $(if $(filter x86_64 i686 amd64,$(shell uname -m)),-mssse3,)
$(CSRC)/%.o: $(CSRC)/%.cc $(HDR)
	g++ -O3 -std=c++17 $(HOST_SIMD) -fPIC -Wall -pthread -c $< -o $@

$(OUT): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl -lz -lpthread

oracle:
	$(MAKE) -C oracle

# CPU model of the bit-sliced K-A kernel body (test infrastructure, see tests/model/)
MODEL := tests/model/libka_bitslice_model.so
model: $(MODEL)
$(MODEL): tests/model/ka_bitslice_model.cpp $(HDR)
	g++ -O2 -std=c++17 -fPIC -Wall -Wno-unknown-pragmas -pthread -I/usr/local/cuda/include -shared -o $@ $<

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log $(OUT) $(CLI) tests/model/*.so
	$(MAKE) -C oracle clean

.PHONY: all oracle model clean
