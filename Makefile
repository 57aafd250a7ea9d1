# Builds libhispmv_cuda.so (the C-ABI CUDA engine), the pyhispmv pybind11 module and the oracle.
# sm_100a only -- there is no other target.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := g++
PY        ?= python
ARCH      := -gencode arch=compute_100a,code=sm_100a
# EXPERIMENTAL=1 also builds the three research kernels of experimental.cu (never selected by the planner)
EXPERIMENTAL ?= 0
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Iinclude -Ihispmv_b200/csrc --expt-relaxed-constexpr $(if $(filter 1,$(EXPERIMENTAL)),-DHISPMV_EXPERIMENTAL,)
PKG       := hispmv_b200
CSRC      := $(PKG)/csrc
OBJDIR    := build/obj
OBJS      := $(OBJDIR)/capi.o $(OBJDIR)/spmv.o $(OBJDIR)/adaptive.o $(OBJDIR)/gemv.o $(OBJDIR)/partition.o $(OBJDIR)/synth.o $(OBJDIR)/exchange.o $(OBJDIR)/batch.o $(OBJDIR)/blocked.o $(OBJDIR)/experimental.o
LIB       := $(PKG)/libhispmv_cuda.so
PYEXT     := $(PKG)/pyhispmv$(shell $(PY) -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC     := $(shell $(PY) -c "import sysconfig,pybind11;print('-I'+sysconfig.get_paths()['include'],'-I'+pybind11.get_include())")

CUSPARSE_REF := tools/libcusparse_ref.so

all: $(LIB) $(PYEXT) $(CUSPARSE_REF) oracle

# comparator only (bench.py's vs_cusparse): the reference's gpu/ baseline call; nothing in the package links it
$(CUSPARSE_REF): tools/cusparse_ref.cu
	$(NVCC) -O2 -shared -Xcompiler -fPIC $(ARCH) $< -o $@ -lcusparse

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/internal.h $(CSRC)/device_utils.cuh $(CSRC)/tile_device.cuh include/hispmv.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

$(PYEXT): $(CSRC)/pyhispmv_bindings.cpp include/hispmv.h $(LIB)
	$(CXX) -O2 -std=c++17 -fPIC -shared -fvisibility=hidden $(PYINC) -Iinclude $< -o $@ \
	    -L$(PKG) -lhispmv_cuda -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(PKG)/pyhispmv*.so $(CUSPARSE_REF)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
