# Build the C-ABI shared library (sm_100a only) in-tree.  `make -j8` compiles the translation units in parallel.
# HS_EXPERIMENT=1 compiles the timing-only experiment switches (getenv-driven phase skips, alternative kernel
# variants) into the library; the product build has none of them.
NVCC ?= nvcc
PKG := hyperscanning_signal_analysis_b200
UNITS := hs_api mvar_kernels frontend_kernels psd_kernels generic_kernels transfer_mma hilbert_kernels gather_kernels criterion_kernels
SRC := $(foreach u,$(UNITS),$(PKG)/csrc/$(u).cu)
OBJDIR := build/obj
OBJ := $(foreach u,$(UNITS),$(OBJDIR)/$(u).o)
HDR := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh include/*.h)
LIB := $(PKG)/libhs_b200.so
DEFS := $(if $(HS_EXPERIMENT),-DHS_EXPERIMENT=1,)
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $(DEFS)

all: $(LIB)

$(OBJDIR)/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJ)

clean:
	rm -rf $(LIB) $(OBJDIR)

.PHONY: all clean
