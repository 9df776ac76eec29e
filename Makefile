# Plain-make build of the native artefacts (the same commands __graft_entry__.build() runs):
#   make            libcloudsc2_b200.so + the three dwarf programs
#   make oracle     the CPU checker (test infrastructure only)
# sm_100a (B200) only; nvcc cross-compiles without a GPU.
NVCC ?= nvcc
CXX  ?= g++
PKG  := dwarf-p-cloudsc2-tl-ad_b200
OBJ  := build/obj
LIB  := $(PKG)/csrc/libcloudsc2_b200.so
BIN  := $(PKG)/bin
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=true \
           -Xcompiler -fPIC,-O2,-fopenmp -I include

CU  := $(wildcard $(PKG)/csrc/*.cu)
CC  := $(PKG)/host/cloudsc2_host.cc $(PKG)/host/cloudsc2_h5.cc $(PKG)/host/cloudsc2_input.cc \
       $(PKG)/host/cloudsc2_h5write.cc
HDR := $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h)
OBJS := $(patsubst %,$(OBJ)/%.o,$(basename $(notdir $(CU) $(CC))))

all: $(LIB) programs

$(OBJ)/%.o: $(PKG)/csrc/%.cu $(HDR) | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ)/%.o: $(PKG)/host/%.cc $(HDR) | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ):
	mkdir -p $@

$(LIB): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $^ -Xcompiler -fopenmp -lgomp

programs: $(BIN)/dwarf-cloudsc2-nl $(BIN)/dwarf-cloudsc2-tl $(BIN)/dwarf-cloudsc2-ad
$(BIN)/dwarf-cloudsc2-nl: $(PKG)/host/dwarf_cloudsc2.cc $(LIB) include/cloudsc2_host.h include/cloudsc2_b200.h
	mkdir -p $(BIN)
	$(CXX) -O2 -std=c++17 -Wall -I include $< -o $@ -L $(PKG)/csrc -lcloudsc2_b200 '-Wl,-rpath,$$ORIGIN/../csrc'
$(BIN)/dwarf-cloudsc2-tl $(BIN)/dwarf-cloudsc2-ad: $(BIN)/dwarf-cloudsc2-nl
	cp $< $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(BIN) oracle/_build

.PHONY: all programs oracle clean
