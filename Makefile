# Builds the C-ABI library (sm_100a only) and the oracle's compiled helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard mvoc_b200/csrc/*.cu)
OBJ := $(patsubst mvoc_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := mvoc_b200/lib/libmvoc_b200.so

STAGED_SRC := $(wildcard mvoc_b200/csrc/staged/*.cu)
STAGED_OBJ := $(patsubst mvoc_b200/csrc/staged/%.cu,build/staged_%.o,$(STAGED_SRC))
STAGED_LIB := mvoc_b200/lib/libmvoc_b200_staged.so

all: $(LIB) $(STAGED_LIB)

# kernels staged for the next round: a separate library, not on the product path
build/staged_%.o: mvoc_b200/csrc/staged/%.cu mvoc_b200/csrc/common.cuh mvoc_b200/csrc/ptx.cuh include/mvoc_b200_staged.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/staged_$*.ptxas.log || (cat build/staged_$*.ptxas.log; exit 1)

$(STAGED_LIB): $(STAGED_OBJ) build/api.o
	@mkdir -p mvoc_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(STAGED_OBJ) build/api.o -lcudart

build/%.o: mvoc_b200/csrc/%.cu mvoc_b200/csrc/common.cuh mvoc_b200/csrc/ptx.cuh include/mvoc_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p mvoc_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB) $(STAGED_LIB)

.PHONY: all clean
