# Builds libmikudance_sm100.so (sm_100a only) in-tree, plus the oracle helpers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC := $(wildcard mikudance_b200/csrc/*.cu)
OBJ := $(patsubst mikudance_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := mikudance_b200/lib/libmikudance_sm100.so

all: $(LIB)

build/%.o: mikudance_b200/csrc/%.cu mikudance_b200/csrc/ptx.cuh mikudance_b200/csrc/host_common.h include/mdk.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p mikudance_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB)
.PHONY: all clean
