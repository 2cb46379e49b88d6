# Builds the CUDA back end (librtat_b200.so: sm_100a kernels + C ABI), the C++ host layer's
# apps/tests, and the oracle.  `python -c "import __graft_entry__ as g; g.build()"` drives this.
NVCC     ?= nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC     := rtatblas_b200/csrc
LIBDIR   := rtatblas_b200/lib
OBJDIR   := build/obj
CUDA_HOME ?= /usr/local/cuda

KERNEL_SRCS := $(wildcard $(CSRC)/*.cu)
KERNEL_OBJS := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(KERNEL_SRCS))

all: $(LIBDIR)/librtat_b200.so host oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/librtat_b200.so: $(KERNEL_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xlinker -rpath,$(CUDA_HOME)/lib64

host: $(LIBDIR)/librtat_b200.so
	@if [ -f rtatblas_b200/host/Makefile ]; then $(MAKE) --no-print-directory -C rtatblas_b200/host; fi

oracle:
	$(MAKE) --no-print-directory -C oracle all

clean:
	rm -rf build $(LIBDIR) && $(MAKE) -C oracle clean
.PHONY: all host oracle clean
