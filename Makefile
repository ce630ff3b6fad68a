# Top-level build: the product (CUDA C-ABI library + host libfiasco) and, separately, the
# oracle (test infrastructure).  `python -c "import __graft_entry__ as g; g.build()"` runs
# `make all`.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference is x86-64 SSE fp32 without contraction; a fused a*b+c would
# round once instead of twice and eventually flip an arg-min.
NVFLAGS  := $(ARCH) -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
            -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -Ifiasco_b200/csrc
LIBDIR   := fiasco_b200/lib
CSRC     := fiasco_b200/csrc
KOBJ     := $(LIBDIR)/tile_kernel.o $(LIBDIR)/ffi.o

.PHONY: all product oracle clean ptxas
all: product oracle

product: $(LIBDIR)/libfiasco_b200.so

$(LIBDIR)/.dir:
	mkdir -p $(LIBDIR) && touch $@

$(LIBDIR)/%.o: $(CSRC)/%.cu $(CSRC)/tile_kernel.cuh include/fiasco_b200.h | $(LIBDIR)/.dir
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(LIBDIR)/$*.ptxas.log || (cat $(LIBDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libfiasco_b200.so: $(KOBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(KOBJ) -cudart static

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIBDIR) && $(MAKE) -C oracle clean
