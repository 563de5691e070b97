# Builds the C-ABI shared library of hand-written sm_100a kernels (in-tree, travels with gpurun).
#   make        product library dualvar_b200/lib/libdualvar_b200.so (+ the oracle's reference install, see oracle/)
#   make diag   diagnostics build dualvar_b200/lib/libdualvar_b200_diag.so: the same sources with -DDV_DIAG (per-role
#               cycle counters in the conv kernels) plus csrc/diag/*.cu (microbenchmarks, TMA probes); used by
#               tests/diag/*.py through DV_LIB_PATH, never by the product path
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC      := dualvar_b200/csrc
OBJDIR    := build/obj
DOBJDIR   := build/obj_diag
LIB       := dualvar_b200/lib/libdualvar_b200.so
DLIB      := dualvar_b200/lib/libdualvar_b200_diag.so
SRCS      := $(wildcard $(CSRC)/*.cu)
DSRCS     := $(wildcard $(CSRC)/diag/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
DOBJS     := $(patsubst $(CSRC)/%.cu,$(DOBJDIR)/%.o,$(SRCS)) $(patsubst $(CSRC)/diag/%.cu,$(DOBJDIR)/diag_%.o,$(DSRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h)

all: $(LIB)

diag: $(DLIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(DOBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(DOBJDIR)
	$(NVCC) $(NVCCFLAGS) -DDV_DIAG -c $< -o $@

$(DOBJDIR)/diag_%.o: $(CSRC)/diag/%.cu $(HDRS)
	@mkdir -p $(DOBJDIR)
	$(NVCC) $(NVCCFLAGS) -DDV_DIAG -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p dualvar_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

$(DLIB): $(DOBJS)
	@mkdir -p dualvar_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(DOBJS) -lcudart

clean:
	rm -rf build $(LIB) $(DLIB)

.PHONY: all diag clean
