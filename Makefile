# Builds the product library skirt_b200/libskirtgpu.so (CUDA, sm_100a) and the test oracles.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: no FMA contraction, so that fp64 geometry matches the reference bit for bit
NVFLAGS := $(EXTRA) $(ARCH) -ccbin $(HOSTCXX) -std=c++17 -O3 -lineinfo -fmad=false --expt-relaxed-constexpr \
           -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-function -Xptxas -v
CSRC := skirt_b200/csrc
# BUILD / LIB can be overridden to build experiment variants side by side (tools/gpu_variants.sh)
BUILD ?= $(CSRC)/build
LIB ?= skirt_b200/libskirtgpu.so
OBJS := $(BUILD)/engine.o $(BUILD)/path_kernels.o $(BUILD)/mc_kernels.o $(BUILD)/comm.o
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/skirtgpu.h
NCCL_INC ?= $(shell python -c "import os,nvidia.nccl as n; print(os.path.join(list(n.__path__)[0],'include'))" 2>/dev/null)
NCCL_LIB ?= $(shell python -c "import os,nvidia.nccl as n; print(os.path.join(list(n.__path__)[0],'lib'))" 2>/dev/null)

all: $(LIB) skirt_b200/libskirthost.so skirt_b200/skirt_b200_run oracle

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -I$(NCCL_INC) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $(OBJS) -ldl

# Host-side set-up library (grid builders, include/skirthost.h).  The Voronoi builder needs the Voro++ library, which the
# reference vendors (Voro/, the voro++ 0.4.x line): it is a third-party dependency compiled from where it lies into
# skirt_b200/host/_voro/libvoro.a (git-ignored, travels to the GPU box like the other built files; never copied into the
# repository).  Without the sources and without a prebuilt archive the library is built without Voronoi support.
REF ?= /root/reference
VORO_A := skirt_b200/host/_voro/libvoro.a
VORO_SRC := $(filter-out %/v_base_wl.cc,$(wildcard $(REF)/Voro/*.cc))
ifneq ($(VORO_SRC),)
$(VORO_A): $(VORO_SRC)
	@mkdir -p skirt_b200/host/_voro/obj
	for f in $(VORO_SRC); do $(HOSTCXX) -std=c++11 -O3 -fPIC -w -Iskirt_b200/host/voro_shim -I$(REF)/Voro -c $$f -o skirt_b200/host/_voro/obj/$$(basename $$f .cc).o || exit 1; done
	mkdir -p skirt_b200/host/_voro/include && cp $(REF)/Voro/*.hh skirt_b200/host/_voro/include/
	ar rcs $@ skirt_b200/host/_voro/obj/*.o
endif
HAVE_VORO := $(if $(or $(VORO_SRC),$(wildcard $(VORO_A))),1,)
HOSTLIB_SRC := skirt_b200/host/GridBuilders.cpp skirt_b200/host/HostAbi.cpp
skirt_b200/libskirthost.so: $(HOSTLIB_SRC) skirt_b200/host/GridBuilders.hpp include/skirthost.h $(if $(HAVE_VORO),$(VORO_A))
	$(HOSTCXX) -std=c++17 -O2 -fPIC -Wall -shared -o $@ $(HOSTLIB_SRC) $(if $(HAVE_VORO),-DSKIRT_WITH_VORO -Iskirt_b200/host/voro_shim -Iskirt_b200/host/_voro/include $(VORO_A))

# C++ host layer (simulation items with the reference's names) + command-line driver, linked against the C ABI only
HOST := skirt_b200/host
skirt_b200/skirt_b200_run: $(HOST)/skirt_b200_run.cpp $(HOST)/SimulationItems.cpp $(HOST)/Output.cpp $(HOST)/SimulationItems.hpp $(HOST)/Output.hpp $(HOST)/GridBuilders.hpp include/skirtgpu.h skirt_b200/libskirtgpu.so skirt_b200/libskirthost.so
	$(HOSTCXX) -std=c++17 -O2 -Wall -o $@ $(HOST)/skirt_b200_run.cpp $(HOST)/SimulationItems.cpp $(HOST)/Output.cpp -Lskirt_b200 -lskirtgpu -lskirthost -Wl,-rpath,'$$ORIGIN' -Wl,-rpath-link,/usr/local/cuda/lib64

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(CSRC)/build skirt_b200/libskirtgpu.so skirt_b200/libskirthost.so skirt_b200/skirt_b200_run
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
