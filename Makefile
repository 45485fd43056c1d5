# Builds the sm_100a shared library (C ABI in include/binius_b200.h) in-tree, plus the CPU oracle
# (test infrastructure) and the C++ conformance test of the host mirror.
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 1886,177
CSRC := binius_b200/csrc
LIB := binius_b200/libbinius_b200.so

all: $(LIB) oracle tests/cpp/conformance tools/keccak_replay_cpp

$(LIB): $(CSRC)/capi.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/binius_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/capi.cu

oracle:
	$(MAKE) -s -C oracle

tests/cpp/conformance: tests/cpp/conformance.cpp binius_b200/host/compute_layer.hpp binius_b200/host/computation_backend.hpp $(LIB) oracle
	g++ -O1 -std=c++17 -o $@ tests/cpp/conformance.cpp -Lbinius_b200 -lbinius_b200 -Loracle -loracle \
	    -Wl,-rpath,'$$ORIGIN/../../binius_b200' -Wl,-rpath,'$$ORIGIN/../../oracle'

tools/keccak_replay_cpp: tools/keccak_replay.cpp binius_b200/host/compute_layer.hpp binius_b200/host/computation_backend.hpp $(LIB)
	g++ -O2 -std=c++17 -o $@ tools/keccak_replay.cpp -Lbinius_b200 -lbinius_b200 -Wl,-rpath,'$$ORIGIN/../binius_b200'

clean:
	rm -f $(LIB) oracle/liboracle.so tests/cpp/conformance

.PHONY: all oracle clean
