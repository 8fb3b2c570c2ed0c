# Builds libmikudance_sm100.so (sm_100a only) in-tree, plus the oracle helpers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC := $(wildcard mikudance_b200/csrc/*.cu)
OBJ := $(patsubst mikudance_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := mikudance_b200/lib/libmikudance_sm100.so

MICRO := build/softmax_loop_bench

all: $(LIB) $(MICRO)

# diagnostic microbenchmark (not product): the attention softmax loop's instruction mix in isolation
build/softmax_loop_bench: tests/micro/softmax_loop_bench.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -o $@ $<

build/%.o: mikudance_b200/csrc/%.cu mikudance_b200/csrc/ptx.cuh mikudance_b200/csrc/attn_common.cuh mikudance_b200/csrc/host_common.h include/mdk.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p mikudance_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB)
.PHONY: all clean
