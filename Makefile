# Builds libmdbg_b200.so (hand-written sm_100a kernels + C ABI) in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC := rust-mdbg_b200/csrc
OUT := rust-mdbg_b200/libmdbg_b200.so
CU := $(wildcard $(CSRC)/*.cu)
CPP := $(wildcard $(CSRC)/*.cpp)
OBJ := $(CU:.cu=.o) $(CPP:.cpp=.o)
HDR := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(OUT) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(CSRC)/%.o: $(CSRC)/%.cpp $(HDR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(OUT): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl -lz -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log $(OUT)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
