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

all: $(LIB) $(PLUGIN) $(ORC_LIB) $(CLI) oracle_ref
lib: $(LIB) $(PLUGIN) $(CLI)
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
