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
        $(OUT)/sblas_api.o $(OUT)/sblas_ingest.o $(OUT)/sblas_shim.o $(OUT)/sblas_spmm.o $(OUT)/sblas_spmm_plan.o $(OUT)/sblas_sptrans.o $(OUT)/sblas_sptrans_plan.o

all: $(OUT)/libsblas_spmv.so $(OUT)/libsblas_spmv_unsafe.so test_spmv test_spmm oracle

$(OUT):
	mkdir -p $(OUT)

$(OUT)/%.o: $(SRC)/%.cu include/sblas_device.h $(SRC)/sblas_dev_common.cuh | $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OUT)/%.o: $(SRC)/%.c include/sblas_device.h $(SRC)/sblas_internal.h include/sblas_spmv.h include/sblas_spmm.h include/sblas_sptrans.h | $(OUT)
	$(HOSTCC) $(CFLAGS) -c $< -o $@
$(OUT)/%.o: $(SRC)/%.cpp include/sblas_spmv.h | $(OUT)
	$(HOSTCXX) -O2 -fPIC $(INC) -c $< -o $@

$(OUT)/libsblas_spmv.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared -Xlinker -rpath,$(CUDA_HOME)/lib64 -lm

# Test artefact: the same library with the stage ring handed back BEFORE the values read from it are
# known to have arrived (the hazard of DESIGN.md section 4.5, -DSBLAS_UNSAFE_EARLY_RELEASE).  Only the
# regression test loads it (SBLAS_LIB), to show that the stress input does catch the pre-fix code.
UNSAFE_OBJS := $(OUT)/unsafe_sblas_spmv_tma.o $(OUT)/unsafe_sblas_spmv_rowtile.o
$(OUT)/unsafe_%.o: $(SRC)/%.cu include/sblas_device.h $(SRC)/sblas_dev_common.cuh | $(OUT)
	$(NVCC) $(NVFLAGS) -DSBLAS_UNSAFE_EARLY_RELEASE -c $< -o $@ 2>/dev/null
$(OUT)/libsblas_spmv_unsafe.so: $(OBJS) $(UNSAFE_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(filter-out $(OUT)/sblas_spmv_tma.o $(OUT)/sblas_spmv_rowtile.o,$(OBJS)) $(UNSAFE_OBJS) -cudart shared -Xlinker -rpath,$(CUDA_HOME)/lib64 -lm

test_spmv: $(SRC)/test_spmv.c $(OUT)/libsblas_spmv.so
	$(HOSTCC) -O2 -Wall -std=gnu11 $(INC) $< -o $@ -L$(OUT) -lsblas_spmv -L$(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/$(OUT)' -Wl,-rpath,$(CUDA_HOME)/lib64 -lm

test_spmm: $(SRC)/test_spmm.c $(OUT)/libsblas_spmv.so
	$(HOSTCC) -O2 -Wall -std=gnu11 $(INC) $< -o $@ -L$(OUT) -lsblas_spmv -L$(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/$(OUT)' -Wl,-rpath,$(CUDA_HOME)/lib64 -lm

oracle: $(OUT)/libsblas_spmv.so
	$(MAKE) -C oracle -s
	$(MAKE) -C oracle -s refharness

clean:
	rm -rf $(OUT) test_spmv test_spmm; $(MAKE) -C oracle clean
.PHONY: all oracle clean
