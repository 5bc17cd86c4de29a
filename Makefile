# Builds the product library (libvlitefast.so, sm_100a only), the C host
# executables and the test oracles.  Everything lands in-tree so that it
# travels to the GPU box with the snapshot.
NVCC     ?= /usr/local/cuda/bin/nvcc
# the image exports CC=/opt/gcc/bin/gcc, a wrapper without libgomp
HOSTCC   := /usr/bin/gcc
GENCODE  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(GENCODE) -lineinfo -Xcompiler -fPIC -Iinclude -Ivlite-fast_b200/csrc
PKG      := vlite-fast_b200
CSRC     := $(PKG)/csrc
HOST     := $(PKG)/host
LIB      := $(PKG)/libvlitefast.so
TESTLIB  := $(PKG)/libvlitefast_testing.so
GENLIB   := $(PKG)/libvlitegen.so

all: $(LIB) $(TESTLIB) $(GENLIB) host oracle build/vf_fft_hosttest build/vf_fft6250_hosttest scripts/ubench/fp32_rate

build:
	mkdir -p build

build/vf_kernels.o: $(CSRC)/vf_kernels.cu $(CSRC)/vf_kernels.h $(CSRC)/vf_fft12500.cuh $(CSRC)/vf_fft6250.cuh $(CSRC)/vf_fft_consts.h $(CSRC)/vf_pass3_map.h | build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/vf_api.o: $(CSRC)/vf_api.cu $(CSRC)/vf_kernels.h $(CSRC)/vf_fft6250.cuh include/vlitefast.h | build
	$(NVCC) $(NVFLAGS) -c $< -o $@

# -Bsymbolic: calls between the two objects bind inside the library, so that the product and the testing
# build can be loaded into one process (tests) without interposing each other's functions
$(LIB): build/vf_kernels.o build/vf_api.o
	$(NVCC) -shared $(GENCODE) -Xlinker -Bsymbolic -o $@ $^ -ldl

# the same sources with -DVF_TESTING: adds the monolithic channeliser (vf_config.k1_threads) and
# vf_debug_division (csrc/vf_testing.h).  Loaded by tests/ for A/B comparisons only.
build/vf_kernels_t.o: $(CSRC)/vf_kernels.cu $(CSRC)/vf_kernels.h $(CSRC)/vf_fft12500.cuh $(CSRC)/vf_fft6250.cuh $(CSRC)/vf_fft_consts.h $(CSRC)/vf_pass3_map.h | build
	$(NVCC) $(NVFLAGS) -DVF_TESTING -c $< -o $@

build/vf_api_t.o: $(CSRC)/vf_api.cu $(CSRC)/vf_kernels.h $(CSRC)/vf_fft6250.cuh $(CSRC)/vf_testing.h include/vlitefast.h | build
	$(NVCC) $(NVFLAGS) -DVF_TESTING -c $< -o $@

$(TESTLIB): build/vf_kernels_t.o build/vf_api_t.o
	$(NVCC) -shared $(GENCODE) -Xlinker -Bsymbolic -o $@ $^ -ldl

# GPU baseband generator: its own library, the only one that links cuFFT
$(GENLIB): $(CSRC)/vf_genbase_gpu.cu include/vlitegen.h | build
	$(NVCC) $(NVFLAGS) -shared -o $@ $< -lcufft

# issue-rate microbenchmark of the packed fp32 instructions
scripts/ubench/fp32_rate: scripts/ubench/fp32_rate.cu
	$(NVCC) -O3 $(GENCODE) -o $@ $<

build/vf_fft_hosttest: $(CSRC)/vf_fft_hosttest.cu $(CSRC)/vf_fft12500.cuh | build
	$(NVCC) -O2 -std=c++17 -I$(CSRC) -o $@ $<

build/vf_fft6250_hosttest: $(CSRC)/vf_fft6250_hosttest.cu $(CSRC)/vf_fft6250.cuh $(CSRC)/vf_fft12500.cuh $(CSRC)/vf_fft_consts.h | build
	$(NVCC) -O2 -std=c++17 -I$(CSRC) -o $@ $<

host: $(LIB) $(GENLIB)
	@if [ -f $(HOST)/Makefile ]; then $(MAKE) -C $(HOST) HOSTCC=$(HOSTCC); fi

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(TESTLIB) $(GENLIB)
	$(MAKE) -C oracle clean

.PHONY: all host oracle clean
