# Builds everything in-tree (the .so files travel with the gpurun snapshot; they are git-ignored).
#   make            host library, CUDA library (sm_100a), oracle port (+ oracle/_ref: reference and drop-in binaries, when /root/reference exists)
ROOT   := $(abspath .)
PKG    := $(ROOT)/lfm_public_b200
HOST   := $(PKG)/host
CSRC   := $(PKG)/csrc
INC    := $(ROOT)/include
# /opt/gcc/bin/g++ (the image default $$CXX) links libstdc++ statically into shared objects, which clashes with the
# libstdc++.so.6 other Python extensions load; the system compiler links it dynamically.
CXX    := /usr/bin/g++
NVCC   ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
NCCL_INC ?= /usr/include
# parity build: no FMA contraction (the CPU reference is built -O3 without -march => no FMA), IEEE div/sqrt
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --fmad=false -prec-div=true -prec-sqrt=true \
           -I$(INC) -I$(NCCL_INC) $(EXTRA_NVFLAGS)
REF    ?= /root/reference

HOST_SRC := $(HOST)/foam_io.cpp $(HOST)/flatten.cpp $(HOST)/hostapi.cpp $(HOST)/hpath.cpp
GPU_SRC  := $(CSRC)/lfmgpu.cu
GPU_HDR  := $(wildcard $(CSRC)/*.cuh) $(INC)/lfmgpu.h

.PHONY: all host gpu oracle clean
all: host gpu oracle

host: $(PKG)/liblfmhost.so
$(PKG)/liblfmhost.so: $(HOST_SRC) $(HOST)/foam_io.h $(HOST)/flatten.h $(INC)/lfmhost.h $(INC)/lfmgpu.h
	$(CXX) -std=c++17 -O2 -fPIC -shared -ffp-contract=off -I$(INC) -I$(HOST) -o $@ $(HOST_SRC)

gpu: $(PKG)/liblfmgpu.so
$(PKG)/liblfmgpu.so: $(GPU_SRC) $(GPU_HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(GPU_SRC) -Xlinker -rpath -Xlinker '$$ORIGIN' -lcudart -ldl

oracle:
	$(MAKE) -C oracle port
	@if [ -d $(REF)/src ]; then $(MAKE) -C oracle ref REF=$(REF); else echo "no $(REF): keeping prebuilt oracle/_ref"; fi

clean:
	rm -f $(PKG)/liblfmhost.so $(PKG)/liblfmgpu.so
	$(MAKE) -C oracle clean
