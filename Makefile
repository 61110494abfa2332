# Builds libMiniCVNative.so (sm_100a only) in-tree.  `make` needs nvcc; no GPU required.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function \
             -cudart static --expt-relaxed-constexpr
CSRC      := minicv_b200/csrc
OBJDIR    := build/obj
LIB       := minicv_b200/lib/libMiniCVNative.so
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/MiniCVNative.h

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS) -ldl -lpthread

# Debug variants for the sanitizer runs recorded in profiles/ (never loaded by the package): the homography refit with
# its thread-0 helpers force-inlined and the 5-point solver composed into one device function -- the two forms round 1
# saw mis-execute and worked around.
DBGDIR    := build/dbg
DBGLIB    := $(DBGDIR)/libMiniCVNative.so
DBGOBJS   := $(patsubst $(CSRC)/%.cu,$(DBGDIR)/obj/%.o,$(SRCS))
$(DBGDIR)/obj/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(DBGDIR)/obj
	$(NVCC) $(NVFLAGS) -DMCV_DEBUG_INLINE_REFIT -DMCV_DEBUG_COMPOSED_E5 $(DBGEXTRA) -c $< -o $@
$(DBGLIB): $(DBGOBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(DBGOBJS) -ldl -lpthread
debug-variants: $(DBGLIB)

clean:
	rm -rf build $(LIB)

.PHONY: all clean debug-variants
