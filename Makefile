# Builds the C-ABI library (libnsd_b200.so) for sm_100a only.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC_DIR := neural_speech_decoder_b200/csrc
BUILD := build
SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS := $(patsubst $(SRC_DIR)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB := neural_speech_decoder_b200/libnsd_b200.so

all: $(LIB)

$(BUILD)/%.o: $(SRC_DIR)/%.cu $(SRC_DIR)/common.cuh $(SRC_DIR)/tc_common.cuh include/nsd_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcuda

clean:
	rm -rf $(BUILD) $(LIB)

.PHONY: all clean
