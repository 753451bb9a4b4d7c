# Builds, in-tree:
#   chain_b200/libchain_b200.so        the product: sm_100a kernels + C++ engine + C ABI (include/chain_b200.h)
#   tests/hostemu/libcb200_hostemu.so  TEST ONLY: the same engine linked against a serial CPU backend (host-logic tests)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -I/usr/include -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -Wno-sign-compare
SRC       := chain_b200/csrc
OBJ       := build

ENGINE_SRCS := $(SRC)/engine.cpp $(SRC)/engine_io.cpp $(SRC)/capi.cpp
HDRS := $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh) include/chain_b200.h

all: chain_b200/libchain_b200.so tests/hostemu/libcb200_hostemu.so

$(OBJ):
	mkdir -p $(OBJ)

$(OBJ)/backend_cuda.o: $(SRC)/backend_cuda.cu $(HDRS) | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/ptxas_backend.log || (cat $(OBJ)/ptxas_backend.log; false)

$(OBJ)/%.o: $(SRC)/%.cpp $(HDRS) | $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

chain_b200/libchain_b200.so: $(OBJ)/backend_cuda.o $(OBJ)/engine.o $(OBJ)/engine_io.o $(OBJ)/capi.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

tests/hostemu/libcb200_hostemu.so: $(ENGINE_SRCS) tests/hostemu/backend_emu.cpp $(HDRS)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(ENGINE_SRCS) tests/hostemu/backend_emu.cpp

clean:
	rm -rf $(OBJ) chain_b200/libchain_b200.so tests/hostemu/libcb200_hostemu.so

.PHONY: all clean
