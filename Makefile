# Builds the C-ABI library (sm_100a only) and the oracle's compiled helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard mvoc_b200/csrc/*.cu)
OBJ := $(patsubst mvoc_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := mvoc_b200/lib/libmvoc_b200.so

all: $(LIB)

build/%.o: mvoc_b200/csrc/%.cu mvoc_b200/csrc/common.cuh mvoc_b200/csrc/ptx.cuh include/mvoc_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p mvoc_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)

.PHONY: all clean
