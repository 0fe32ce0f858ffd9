# Builds libanemoi_b200.so (sm_100a only), the C oracle, and the standalone IMAD microbenchmark.
NVCC      ?= nvcc
CXX       := g++
CC        := gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 -lineinfo $(ARCH) -Iinclude -Xcompiler -fPIC -Xptxas -v $(EXTRA_NVFLAGS)
CSRC      := anemoi_rust_b200/csrc
FIELDS    := bls12_377 bls12_381 bn_254 ed_on_bls12_377 jubjub pallas vesta
OBJS      := $(patsubst %,build/field_%.o,$(FIELDS)) build/api.o build/imad_peak.o build/merkle_aux.o
LIB       := anemoi_rust_b200/libanemoi_b200.so
HDRS      := $(CSRC)/fp.cuh $(CSRC)/anemoi_kernels.cuh $(CSRC)/field_tu.cuh $(CSRC)/launch.h $(CSRC)/kernel_args.h $(CSRC)/merkle_aux.h $(CSRC)/nccl_dyn.h \
             $(CSRC)/generated/fields.cuh include/anemoi_b200.h

all: $(LIB) oracle/libanemoi_oracle.so tools/imad_peak

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $< > build/$*.ptxas.log 2>&1 || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

tools/imad_peak: tools/imad_peak_main.cu $(LIB)
	$(NVCC) -std=c++17 -O3 $(ARCH) -Iinclude -o $@ tools/imad_peak_main.cu -Lanemoi_rust_b200 -lanemoi_b200 -Xlinker -rpath='$$ORIGIN/../anemoi_rust_b200'

oracle/libanemoi_oracle.so: oracle/anemoi_oracle.c oracle/params_gen.h
	$(CC) -O3 -funroll-loops -march=x86-64-v3 -fopenmp -shared -fPIC -o $@ oracle/anemoi_oracle.c

clean:
	rm -rf build $(LIB) oracle/libanemoi_oracle.so tools/imad_peak

.PHONY: all clean
