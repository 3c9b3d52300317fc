# Builds kzg_b200/libkzg_b200.so (sm_100a only) and the CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# A/B builds of the library for experiments: make BUILD=build_x LIB=kzg_b200/libkzg_b200_x.so EXTRA="-DKZG_..." lib
# (select at run time with KZGB_LIB=<path>); the default build takes no EXTRA flags.
EXTRA ?=
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v $(EXTRA)
SRC_DIR := kzg_b200/csrc
BUILD ?= build
SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS := $(patsubst $(SRC_DIR)/%.cu,$(BUILD)/%.o,$(SRCS))
HDRS := $(wildcard $(SRC_DIR)/*.cuh) include/kzg_b200.h
LIB ?= kzg_b200/libkzg_b200.so

all: $(LIB) oracle $(BUILD)/host_mirror_demo $(BUILD)/multi_gpu_check $(BUILD)/threads_check
lib: $(LIB)

# compile check of the header-only C++ host mirror (runs only on a GPU box)
$(BUILD)/host_mirror_demo: tests/cpp/host_mirror_demo.cpp include/kzg_b200.hpp include/kzg_b200.h $(LIB)
	g++ -std=c++17 -O2 -o $@ $< -Lkzg_b200 -lkzg_b200 -Wl,-rpath,'$$ORIGIN/../kzg_b200'

# C++ (no Python) checks of the C ABI: multi-process multi-GPU parity, concurrent host threads (run on a GPU box)
$(BUILD)/multi_gpu_check: tests/cpp/multi_gpu_check.cpp include/kzg_b200.h $(LIB)
	g++ -std=c++17 -O2 -pthread -o $@ $< -Lkzg_b200 -lkzg_b200 -Wl,-rpath,'$$ORIGIN/../kzg_b200'
$(BUILD)/threads_check: tests/cpp/threads_check.cpp include/kzg_b200.h $(LIB)
	g++ -std=c++17 -O2 -pthread -o $@ $< -Lkzg_b200 -lkzg_b200 -Wl,-rpath,'$$ORIGIN/../kzg_b200'

$(BUILD)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -C oracle -s

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all lib oracle clean
