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
KOBJ     := $(LIBDIR)/tile_kernel.o $(LIBDIR)/ffi.o $(LIBDIR)/motion_kernel.o

.PHONY: all product oracle clean ptxas
all: product oracle

product: $(LIBDIR)/libfiasco_b200.so $(LIBDIR)/libfiasco.so cfiasco

# host side of libfiasco (plain C): options, PNM input, .fco writer, fiasco_coder()
HOST     := fiasco_b200/host
HOSTSRC  := messages host_util c_options pnm_input bitstream fco_writer regenerate coder_api
HOSTOBJ  := $(addprefix $(LIBDIR)/host_,$(addsuffix .o,$(HOSTSRC)))
HCFLAGS  := -O2 -g -std=gnu11 -fPIC -Wall -Wno-unused-function -ffp-contract=off -Iinclude -I$(HOST)

# the PNM reader converts every sample of every frame (u8 -> the coder's 12.4 shorts): vectorised
$(LIBDIR)/host_pnm_input.o: HCFLAGS += -O3
$(LIBDIR)/host_%.o: $(HOST)/%.c $(HOST)/fi_internal.h include/fiasco.h include/fiasco_b200.h | $(LIBDIR)/.dir
	gcc $(HCFLAGS) -c $< -o $@

$(LIBDIR)/libfiasco.so: $(HOSTOBJ) $(LIBDIR)/libfiasco_b200.so
	gcc -shared -o $@ $(HOSTOBJ) -L$(LIBDIR) -lfiasco_b200 -Wl,-rpath,'$$ORIGIN' -lm

# The reference command line front end, compiled UNCHANGED from where it lies, against OUR
# include/fiasco.h and linked against OUR libfiasco.  Only possible where
# /root/reference exists; the binary travels to the GPU box with the snapshot.
REF      ?= /root/reference
CLI_SRC  := cwfa params binerror getopt getopt1
.PHONY: cfiasco
ifneq ($(wildcard $(REF)/bin/cwfa.c),)
cfiasco: $(LIBDIR)/cfiasco
$(LIBDIR)/cfiasco: $(addprefix $(REF)/bin/,$(addsuffix .c,$(CLI_SRC))) $(LIBDIR)/libfiasco.so
	gcc -O2 -w -fcommon -Iinclude -Ioracle/refcfg -I$(REF)/lib -I$(REF)/bin \
	    -o $@ $(addprefix $(REF)/bin/,$(addsuffix .c,$(CLI_SRC))) \
	    -L$(LIBDIR) -lfiasco -lfiasco_b200 -Wl,-rpath,'$$ORIGIN' -lm
else
cfiasco:
	@echo "cfiasco: $(REF) not present -- using prebuilt $(LIBDIR)/cfiasco (if any)"
endif

$(LIBDIR)/.dir:
	mkdir -p $(LIBDIR) && touch $@

$(LIBDIR)/%.o: $(CSRC)/%.cu $(CSRC)/tile_kernel.cuh include/fiasco_b200.h | $(LIBDIR)/.dir
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(LIBDIR)/$*.ptxas.log || (cat $(LIBDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libfiasco_b200.so: $(KOBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(KOBJ) -cudart static

# diagnostics build: thread-0 lap timers per sub-phase (FB200_LIB=gpurun_exp/laps/libfiasco_b200.so)
.PHONY: laps
laps:
	mkdir -p gpurun_exp/laps
	for f in tile_kernel ffi motion_kernel; do \
	  $(NVCC) $(NVFLAGS) -DFB200_LAPS -c $(CSRC)/$$f.cu -o gpurun_exp/laps/$$f.o || exit 1; done
	$(NVCC) $(ARCH) -shared -o gpurun_exp/laps/libfiasco_b200.so gpurun_exp/laps/tile_kernel.o \
	    gpurun_exp/laps/ffi.o gpurun_exp/laps/motion_kernel.o -cudart static

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIBDIR) && $(MAKE) -C oracle clean
