# Builds the B200-native SpMV library, the CLI and the checker.
#   make            -> s-blas_b200/lib/libsblas_spmv.so  +  test_spmv  (+ oracle/)
# sm_100a only; nvcc cross-compiles without a GPU.
NVCC   ?= /usr/local/cuda/bin/nvcc
HOSTCC ?= /usr/bin/gcc
HOSTCXX ?= /usr/bin/g++
CUDA_HOME ?= /usr/local/cuda
ARCH   := -gencode arch=compute_100a,code=sm_100a
SRC    := s-blas_b200/csrc
OUT    := s-blas_b200/lib
INC    := -Iinclude -I$(SRC) -I$(CUDA_HOME)/include
CFLAGS := -O2 -fPIC -Wall -Wno-unused-function -std=gnu11 $(INC)
NVFLAGS := $(ARCH) -O3 -lineinfo -Xptxas -v -Xcompiler -fPIC $(INC)

OBJS := $(OUT)/sblas_kernels.o $(OUT)/sblas_spmv_tma.o $(OUT)/sblas_spmv_rowtile.o $(OUT)/sblas_synth.o $(OUT)/sblas_partition.o $(OUT)/sblas_plan.o \
        $(OUT)/sblas_api.o $(OUT)/sblas_ingest.o $(OUT)/sblas_shim.o

all: $(OUT)/libsblas_spmv.so test_spmv oracle

$(OUT):
	mkdir -p $(OUT)

$(OUT)/%.o: $(SRC)/%.cu include/sblas_device.h $(SRC)/sblas_dev_common.cuh | $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OUT)/%.o: $(SRC)/%.c include/sblas_device.h $(SRC)/sblas_internal.h include/sblas_spmv.h | $(OUT)
	$(HOSTCC) $(CFLAGS) -c $< -o $@
$(OUT)/%.o: $(SRC)/%.cpp include/sblas_spmv.h | $(OUT)
	$(HOSTCXX) -O2 -fPIC $(INC) -c $< -o $@

$(OUT)/libsblas_spmv.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared -Xlinker -rpath,$(CUDA_HOME)/lib64 -lm

test_spmv: $(SRC)/test_spmv.c $(OUT)/libsblas_spmv.so
	$(HOSTCC) -O2 -Wall -std=gnu11 $(INC) $< -o $@ -L$(OUT) -lsblas_spmv -L$(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/$(OUT)' -Wl,-rpath,$(CUDA_HOME)/lib64 -lm

oracle: $(OUT)/libsblas_spmv.so
	$(MAKE) -C oracle -s
	$(MAKE) -C oracle -s refharness

clean:
	rm -rf $(OUT) test_spmv; $(MAKE) -C oracle clean
.PHONY: all oracle clean
