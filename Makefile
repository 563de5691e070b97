# Builds the C-ABI shared library of hand-written sm_100a kernels (in-tree, travels with gpurun).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC      := dualvar_b200/csrc
OBJDIR    := build/obj
LIB       := dualvar_b200/lib/libdualvar_b200.so
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h)

all: $(LIB) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p dualvar_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

oracle:
	@true

clean:
	rm -rf build $(LIB)

.PHONY: all clean oracle
