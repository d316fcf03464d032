# Build the C-ABI shared library (sm_100a only) in-tree.
NVCC ?= nvcc
PKG := hyperscanning_signal_analysis_b200
SRC := $(PKG)/csrc/hs_api.cu $(PKG)/csrc/mvar_kernels.cu $(PKG)/csrc/frontend_kernels.cu $(PKG)/csrc/generic_kernels.cu $(PKG)/csrc/transfer_mma.cu $(PKG)/csrc/hilbert_kernels.cu
HDR := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh include/*.h)
LIB := $(PKG)/libhs_b200.so
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -cudart static

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC)

clean:
	rm -f $(LIB)
