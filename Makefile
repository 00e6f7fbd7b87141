# Builds libscb200.so (sm_100a) in-tree.  `make` here or __graft_entry__.build().
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(EXTRA) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-O3,-Wall -Xptxas -v
SRC := $(wildcard springcraft_b200/csrc/*.cu)
OBJ := $(patsubst springcraft_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := springcraft_b200/lib/libscb200.so

all: $(LIB)

build/%.o: springcraft_b200/csrc/%.cu springcraft_b200/csrc/stedc_core.cuh springcraft_b200/csrc/common.cuh springcraft_b200/csrc/subspace.cuh springcraft_b200/csrc/jacobi.cuh springcraft_b200/csrc/paired.cuh springcraft_b200/csrc/resident.cuh include/scb200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p springcraft_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ)

clean:
	rm -rf build $(LIB)
