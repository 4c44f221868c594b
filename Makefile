# swegl_b200 — product build (the test infrastructure has its own recipe: oracle/Makefile).
#
#   make lib                      swegl_b200/libswegl_b200.so: CUDA kernels + C ABI for sm_100a (what `_build.py` runs)
#   make host SWEGL=<checkout>    build/renderer_b200.o: the ONE translation unit that replaces swegl's
#                                 src/render/renderer.cpp, compiled against the user's swegl headers --
#                                 link it with swegl's other objects and -lswegl_b200 instead of renderer.o
#   make host-check SWEGL=<...>   also compiles the header-only C++ hosts (adapter, pipeline_t, sharded_renderer_t) on their own
#   make clean
#
# swegl's own build needs freon (the author's matrix header, not vendored in the checkout) and SDL2 on the include path:
# pass them with HOST_INC, e.g. HOST_INC="-I/path/to/freon -I/usr/include/SDL2".  The tests build the same unit against
# shims (oracle/Makefile, target dropin).
NVCC      ?= $(shell command -v nvcc || echo /usr/local/cuda/bin/nvcc)
CXX       ?= g++
SWEGL     ?= /root/reference
HOST_INC  ?=
NVCCFLAGS  = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
             -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v --shared -Iinclude
CSRC       = swegl_b200/csrc/abi.cu swegl_b200/csrc/geometry.cu swegl_b200/csrc/fragment.cu swegl_b200/csrc/animate.cu \
             swegl_b200/host/image_decode.cpp
HOSTFLAGS  = -O3 -DNDEBUG -msse4 --std=c++2a -fPIC -w -Iinclude -Iswegl_b200/host -I$(SWEGL) -I$(SWEGL)/src $(HOST_INC)

lib: swegl_b200/libswegl_b200.so
swegl_b200/libswegl_b200.so: $(CSRC) $(wildcard swegl_b200/csrc/*.cuh swegl_b200/csrc/*.h) include/swegl_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(CSRC) -lz

host: build/renderer_b200.o
build/renderer_b200.o: swegl_b200/host/renderer_b200.cpp swegl_b200/host/swegl_b200_adapter.hpp include/swegl_b200.h
	mkdir -p build
	$(CXX) $(HOSTFLAGS) -c -o $@ $<

host-check: host
	printf '#include "swegl_b200_host.hpp"\n' | $(CXX) $(HOSTFLAGS) -x c++ -fsyntax-only -

clean:
	rm -rf build swegl_b200/libswegl_b200.so
.PHONY: lib host host-check clean
