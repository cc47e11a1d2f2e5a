# Build: product library (CUDA sm_100a + host C++), CPU oracle (test infrastructure).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CC        := $(shell test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)
ARCH      := -gencode arch=compute_100a,code=sm_100a
EXTRA_NVFLAGS ?=
NVFLAGS   := $(EXTRA_NVFLAGS) -ccbin $(CXX) $(ARCH) -O3 -lineinfo -std=c++17 -Iinclude -Ixmimsim_b200/csrc -Xcompiler -fPIC,-fopenmp,-O3 -Xptxas -v
CXXFLAGS  := -O3 -fPIC -fopenmp -std=c++17 -Iinclude -Ixmimsim_b200/csrc -Wall -Wno-unknown-pragmas
CFLAGS    := -O3 -fPIC -fopenmp -std=gnu99 -Iinclude -Ixmimsim_b200/csrc -Wall
SRC       := xmimsim_b200/csrc
OBJ       := build/obj
LIB       := xmimsim_b200/lib/libxmimsim_b200.so
CU_SRCS   := $(wildcard $(SRC)/*.cu)
CPP_SRCS  := $(wildcard $(SRC)/*.cpp)
C_SRCS    := $(wildcard $(SRC)/*.c)
OBJS      := $(patsubst $(SRC)/%.cu,$(OBJ)/%.cu.o,$(CU_SRCS)) $(patsubst $(SRC)/%.cpp,$(OBJ)/%.cpp.o,$(CPP_SRCS)) $(patsubst $(SRC)/%.c,$(OBJ)/%.c.o,$(C_SRCS))
HDRS      := $(wildcard include/*.h) $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh)

ORC_SRCS  := $(wildcard oracle/*.c)
ORC_LIB   := oracle/_build/liborc.so

PLUGIN    := xmimsim_b200/lib/xmimsim-cl.so

CLI       := bin/xmimsim-b200
INTERPOSE := xmimsim_b200/lib/libxmimsim-b200-interpose.so
HARNESS   := xmimsim_b200/lib/plugin_harness

all: $(LIB) $(PLUGIN) $(INTERPOSE) $(HARNESS) $(ORC_LIB) $(CLI) oracle_ref
lib: $(LIB) $(PLUGIN) $(INTERPOSE) $(HARNESS) $(CLI)
oracle: $(ORC_LIB)

$(OBJ)/%.cu.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)
	@grep -E "registers|spill|error|warning" $(OBJ)/$*.ptxas.log | grep -v "^$$" | sed 's/^/  [ptxas $*] /' | grep -E "Used|spill" | head -40 || true

$(OBJ)/%.cpp.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJ)/%.c.o: $(SRC)/%.c $(HDRS)
	@mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p xmimsim_b200/lib
	$(NVCC) -ccbin $(CXX) $(ARCH) -shared -o $@ $(OBJS) -Xcompiler -fopenmp -lgomp -ldl -cudart static

# Drop-in plugin file name the reference's loader opens (src/xmi_solid_angle.c:121-136: "<dir>/xmimsim-cl.<so>");
# the same symbols are inside $(LIB); this is that library under the expected name (a relative symlink).
$(PLUGIN): $(LIB)
	ln -sf $(notdir $(LIB)) $(PLUGIN)

# The reference-named history entry point xmi_main_msim (include/xmi_main.h:29) lives in its own file: LD_PRELOAD it into
# the reference's host to route its xmi_main_msim calls to the GPU; inside the plugin file the name
# would clash with libxmimsim's own.
$(INTERPOSE): $(SRC)/plugin_shim.cpp $(LIB) $(HDRS)
	$(CXX) $(CXXFLAGS) -DXMB_EXPORT_XMI_MAIN_MSIM -DXMB_INTERPOSE_ONLY -shared -o $@ $(SRC)/plugin_shim.cpp -Lxmimsim_b200/lib -lxmimsim_b200 -ldl -Wl,-rpath,'$$ORIGIN'

# C harness that calls the reference-named symbols the way bin/xmimsim.c does (tests/test_plugin_harness_gpu.py runs it)
$(HARNESS): tests/c/plugin_harness.c $(LIB) $(INTERPOSE) include/xmimsim_b200.h
	$(CC) -O1 -std=gnu99 -Wall -Iinclude -rdynamic -o $@ tests/c/plugin_harness.c -Lxmimsim_b200/lib -lxmimsim_b200 -ldl -Wl,-rpath,'$$ORIGIN'

# Command-line driver with the reference's options (bin/xmimsim.c); finds the library next to the package.
$(CLI): bin/xmimsim_main.cpp $(LIB) include/xmimsim_b200.h
	$(CXX) -O2 -std=c++17 -Iinclude -o $@ bin/xmimsim_main.cpp -Lxmimsim_b200/lib -lxmimsim_b200 -ldl -Wl,-rpath,'$$ORIGIN/../xmimsim_b200/lib'

# The oracle links the surrogate provider object (third-party stand-in), never the engine.
$(ORC_LIB): $(ORC_SRCS) oracle/oracle.h oracle/orc_rng.h include/xmimsim_b200.h $(SRC)/xrl_surrogate.c
	@mkdir -p oracle/_build
	$(CC) $(CFLAGS) -Ioracle -shared -o $@ $(ORC_SRCS) $(SRC)/xrl_surrogate.c -lm

# The reference's own sources that compile with gcc alone (OpenCL solid-angle kernel through a shim, cubic spline):
# built where /root/reference exists, kept as a prebuilt file elsewhere (oracle/ref_shim/README.md).
oracle_ref:
	sh oracle/build_ref.sh

clean:
	rm -rf build xmimsim_b200/lib oracle/_build $(CLI)
.PHONY: all lib oracle oracle_ref clean
