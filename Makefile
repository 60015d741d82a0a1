# Top-level build: libb200rs.so (CUDA, sm_100a only), the oracle, and the drop-in unit test.
#   make            -> oclradixsort_b200/libb200rs.so
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref when /root/reference exists)
#   make unittest   -> oracle/_ref/UnitTest64: the reference's UNCHANGED UnitTest/main.cpp compiled
#                      against include/ and linked with libb200rs.so (build container only)
NVCC   ?= /usr/local/cuda/bin/nvcc
CXX    ?= g++
REF    ?= /root/reference
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v -Iinclude
CSRC   := oclradixsort_b200/csrc
SRCS   := $(CSRC)/b200rs_device.cu $(CSRC)/b200rs_scan.cu $(CSRC)/b200rs_sort.cu $(CSRC)/b200rs_host.cu
OBJS   := $(SRCS:.cu=.o)
LIB    := oclradixsort_b200/libb200rs.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/b200rs_internal.h include/b200rs.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(@:.o=.ptxas.log) || (cat $(@:.o=.ptxas.log); false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle all
	if [ -d $(REF) ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)

.PHONY: all oracle clean
