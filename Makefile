# Top-level build: libb200rs.so (CUDA, sm_100a only), the oracle, and the drop-in unit test.
#   make            -> oclradixsort_b200/libb200rs.so
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref when /root/reference exists)
#   make unittest   -> oracle/_ref/UnitTest64: the reference's UNCHANGED UnitTest/main.cpp compiled
#                      against include/ and linked with libb200rs.so (build container only)
NVCC   ?= /usr/local/cuda/bin/nvcc
CXX    ?= g++
REF    ?= /root/reference
ARCH   := -gencode arch=compute_100a,code=sm_100a
# make EXPERIMENTS=1: also compiles the measurement-only kernel variants and the B200RS_* environment knobs (tools/sweep.py);
# the default (production) library contains neither.
EXPFLAGS := $(if $(EXPERIMENTS),-DB200RS_EXPERIMENTS=$(EXPERIMENTS),)
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v -Iinclude $(EXPFLAGS)
CSRC   := oclradixsort_b200/csrc
SRCS   := $(CSRC)/b200rs_device.cu $(CSRC)/b200rs_scan.cu $(CSRC)/b200rs_sort.cu $(CSRC)/b200rs_host.cu $(CSRC)/b200rs_prims.cu
OBJS   := $(SRCS:.cu=.o)
LIB    := oclradixsort_b200/libb200rs.so

TAHOE_LIB := oclradixsort_b200/libtahoe_pprims.so

all: $(LIB) $(TAHOE_LIB)

# Tahoe::Pprims (C++ drop-in, include/Tahoe/ParallelPrimitives/Pprims.h) for C++ callers; plain g++, links libb200rs.so
$(TAHOE_LIB): $(CSRC)/Pprims.cpp $(LIB) $(wildcard include/Adl/*.h include/Tahoe/*/*.h include/Tahoe/*/*/*.h)
	$(CXX) -std=c++11 -O2 -Wall -fPIC -shared -Iinclude $(CSRC)/Pprims.cpp -Loclradixsort_b200 -lb200rs -Wl,-rpath,'$$ORIGIN' -o $@

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/b200rs_internal.h include/b200rs.h $(wildcard $(CSRC)/*.cuh)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(@:.o=.ptxas.log) || (cat $(@:.o=.ptxas.log); false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

# Experiments build (measurement-only kernel variants + B200RS_* environment knobs), kept apart from the product:
#   make experiments  ->  tools/_build/libb200rs_exp.so ;  B200RS_LIB=tools/_build/libb200rs_exp.so python tools/msd_probe.py ...
#   make experiments EXPERIMENTS=2 also compiles the round-1 scatter-pass sweep lists (B200RS_KEYS_VARIANT / _PAIRS_VARIANT, tools/sweep.py)
EXP_DIR  := tools/_build/exp
EXP_OBJS := $(patsubst $(CSRC)/%.cu,$(EXP_DIR)/%.o,$(SRCS))
$(EXP_DIR)/%.o: $(CSRC)/%.cu $(CSRC)/b200rs_internal.h include/b200rs.h $(wildcard $(CSRC)/*.cuh)
	mkdir -p $(EXP_DIR)
	$(NVCC) $(NVFLAGS) -DB200RS_EXPERIMENTS=$(or $(EXPERIMENTS),1) -c $< -o $@ 2> $(@:.o=.ptxas.log) || (cat $(@:.o=.ptxas.log); false)
experiments: $(EXP_OBJS)
	$(NVCC) $(ARCH) -shared -o tools/_build/libb200rs_exp.so $(EXP_OBJS)
.PHONY: experiments

oracle:
	$(MAKE) -C oracle all
	if [ -d $(REF) ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB) $(TAHOE_LIB)

.PHONY: all oracle clean

# ---- drop-in check: the reference's UNCHANGED UnitTest/main.cpp against include/ + libb200rs.so ----
# Built only where $(REF) exists (the build container); the binary travels to the GPU box in oracle/_ref/.
# The CPU side of the comparison (Tahoe::RadixSort::sort) is the reference's own RadixSort.cpp.
UT_DEFS := -DNDEBUG -DTH_LOG_LEVEL=3 -DTH_UNIT_TEST -DTH_NON_WINDOW_TEST -D__LINUX__ -D_X64   # premake4.lua:23,42,46,61,81
UT_OUT  := oracle/_ref

unittest: $(LIB)
	mkdir -p $(UT_OUT)
	test -f $(UT_OUT)/gtest-all.o || $(CXX) -std=c++11 -O1 -w -I$(REF)/contrib/include -I$(REF)/contrib/src/gtest-1.6.0 -c $(REF)/contrib/src/gtest-1.6.0/gtest-all.cc -o $(UT_OUT)/gtest-all.o
	$(CXX) -std=c++11 -O2 -w $(UT_DEFS) -Iinclude -I$(REF)/contrib/include \
	    $(REF)/UnitTest/main.cpp $(CSRC)/Pprims.cpp $(REF)/Tahoe/Algorithm/Sort/RadixSort.cpp $(UT_OUT)/gtest-all.o \
	    -Loclradixsort_b200 -lb200rs -lpthread -Wl,-rpath,'$$ORIGIN/../../oclradixsort_b200' -o $(UT_OUT)/UnitTest64

# Same program, but with the CPU check provided by the repo's own oracle (buildable without $(REF) sources
# except main.cpp/gtest, kept for symmetry; not used on the GPU box).
.PHONY: unittest

# External comparison only (north_star): CUB DeviceRadixSort / DeviceScan timed on the same sizes.  A stand-alone
# binary under tools/_build/ (git-ignored, travels to the GPU box); the product never links or calls it.
cub_compare: tools/_build/cub_compare
tools/_build/cub_compare: tools/cub_compare.cu
	mkdir -p tools/_build
	$(NVCC) $(ARCH) -O3 -std=c++17 -o $@ $<
.PHONY: cub_compare

# C++ caller of the drop-in headers for the section-8f rows (copy/fill, Stopwatch, profile CSV); needs no reference
# sources.  tools/_build/ is git-ignored and travels to the GPU box; run by tests/test_gpu_prims.py.
prims_test: tools/_build/prims_dropin
tools/_build/prims_dropin: tests/cpp/prims_dropin.cpp $(CSRC)/Pprims.cpp $(LIB) $(wildcard include/Adl/*.h include/Tahoe/*/*.h include/Tahoe/*/*/*.h)
	mkdir -p tools/_build
	$(CXX) -std=c++11 -O2 -Wall -DNDEBUG -Iinclude tests/cpp/prims_dropin.cpp $(CSRC)/Pprims.cpp \
	    -Loclradixsort_b200 -lb200rs -Wl,-rpath,'$$ORIGIN/../../oclradixsort_b200' -o $@
.PHONY: prims_test

# C++ caller of the partitioned sort: one process, one host thread per GPU (needs >= 2 GPUs to do anything); run by tests/test_gpu_dist.py
dist_test: tools/_build/dist_dropin
tools/_build/dist_dropin: tests/cpp/dist_dropin.cpp $(CSRC)/Pprims.cpp $(LIB) $(wildcard include/Adl/*.h include/Tahoe/*/*.h include/Tahoe/*/*/*.h)
	mkdir -p tools/_build
	$(CXX) -std=c++11 -O2 -Wall -DNDEBUG -Iinclude tests/cpp/dist_dropin.cpp $(CSRC)/Pprims.cpp \
	    -Loclradixsort_b200 -lb200rs -lpthread -Wl,-rpath,'$$ORIGIN/../../oclradixsort_b200' -o $@
.PHONY: dist_test
