# Builds the sm_100a shared library (C ABI in include/binius_b200.h) in-tree, plus the CPU oracle.
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 1886
CSRC := binius_b200/csrc
LIB := binius_b200/libbinius_b200.so

all: $(LIB) oracle

$(LIB): $(CSRC)/capi.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/binius_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/capi.cu

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(LIB) oracle/liboracle.so

.PHONY: all oracle clean
