# gel-b200 build: everything is built IN-TREE (the .so files travel to the GPU box with the snapshot).
#
#   make            -> gel_b200/libgelcu.so (CUDA, sm_100a), gel_b200/libgelhost.so + gel_b200/host/gel (host C),
#                      oracle/libgeloracle.so (test oracle)
#   make ref        -> oracle/_ref/* (unmodified reference, needs /root/reference)
#
# --fmad=false: the reference's fp32 expressions must not be contracted into FMAs (bit-exact contract);
# the kernels additionally spell every operation as an _rn intrinsic.

NVCC     ?= nvcc
CC       ?= gcc
ARCH      = -gencode arch=compute_100a,code=sm_100a
NVFLAGS   = $(ARCH) -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off
CSTRICT   = -std=c99 -O2 -ffp-contract=off -Wall -Wextra -pedantic -fPIC

all: gel_b200/libgelcu.so gel_b200/libgelhost.so gel_b200/host/gel oracle

gel_b200/libgelcu.so: $(wildcard gel_b200/csrc/*) include/gelcu.h
	$(NVCC) $(NVFLAGS) -shared gel_b200/csrc/gelcu.cu -o $@

gel_b200/libgelhost.so: gel_b200/host/gel_host.c gel_b200/host/gel_host.h
	$(CC) $(CSTRICT) -shared -pthread gel_b200/host/gel_host.c -o $@ -lm

gel_b200/host/gel: gel_b200/host/gel.c gel_b200/host/gel_host.c gel_b200/host/gel_host.h include/gelcu.h gel_b200/libgelcu.so
	$(CC) $(CSTRICT) -Iinclude gel_b200/host/gel.c gel_b200/host/gel_host.c -o $@ \
	  -Lgel_b200 -lgelcu -Wl,-rpath,'$$ORIGIN/..' -lm -lpthread

oracle:
	$(MAKE) -C oracle all

ref:
	$(MAKE) -C oracle ref

sass:
	cuobjdump -sass gel_b200/libgelcu.so > build/gelcu.sass

clean:
	rm -f gel_b200/libgelcu.so gel_b200/libgelhost.so gel_b200/host/gel
	$(MAKE) -C oracle clean

.PHONY: all oracle ref clean sass
