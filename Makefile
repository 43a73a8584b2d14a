# Build the C-ABI CUDA library (sm_100a only), the C oracle helpers and the GPU self-test.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC      := $(wildcard dsmnet_b200/csrc/*.cu)
HDRS      := $(wildcard dsmnet_b200/csrc/*.cuh) include/dsmnet_b200.h
OBJ       := $(patsubst dsmnet_b200/csrc/%.cu,build/%.o,$(CSRC))
LIB       := dsmnet_b200/libdsmnet_b200.so

all: $(LIB) tests/cuda/conv3d_selftest tests/cuda/mma_bench tests/cuda/mnmajor_test

build/%.o: dsmnet_b200/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

tests/cuda/conv3d_selftest: tests/cuda/conv3d_selftest.cu $(LIB)
	$(NVCC) $(ARCH) -lineinfo -O2 -std=c++17 -o $@ $< -Ldsmnet_b200 -ldsmnet_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../../dsmnet_b200'

tests/cuda/mma_bench: tests/cuda/mma_bench.cu dsmnet_b200/csrc/ptx.cuh
	$(NVCC) $(ARCH) -lineinfo -O2 -std=c++17 -o $@ $<

tests/cuda/mnmajor_test: tests/cuda/mnmajor_test.cu dsmnet_b200/csrc/ptx.cuh dsmnet_b200/csrc/tma_host.cuh
	$(NVCC) $(ARCH) -lineinfo -O2 -std=c++17 -o $@ $<

clean:
	rm -rf build $(LIB) tests/cuda/conv3d_selftest tests/cuda/mma_bench tests/cuda/mnmajor_test

.PHONY: all clean
