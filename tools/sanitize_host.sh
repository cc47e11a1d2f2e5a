#!/bin/bash
# CPU-only sanitizer passes over everything that runs on the host (no GPU needed):
#  1. oracle built with -ftrivial-auto-var-init=pattern / =zero: the history digests of ten inputs must equal the default
#     build's (a read of an uninitialised automatic variable would change them);
#  2. oracle under AddressSanitizer + UBSan on the same inputs;
#  3. the product library's host code (table generator, device layouts up to the first upload, XML I/O, caches, Ebel,
#     plugin shims) under AddressSanitizer: the "no CUDA device" guard of xmb_main_msim_raw is removed in a scratch copy so
#     that build_device_tables runs to its first cudaMalloc; then the whole CPU test suite against that build.
# Scratch files under /tmp/xmb_sanitize; prints one line per pass (a fourth pass, heap perturbation, is described at the
# end of the file).  Result of the round-1 run: all four clean.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/xmb_sanitize
mkdir -p $W/obj $W/src
CF="-O1 -g -fPIC -fopenmp -std=gnu99 -I$ROOT/include -I$ROOT/xmimsim_b200/csrc -I$ROOT/oracle"
SRC="$ROOT/oracle/*.c $ROOT/xmimsim_b200/csrc/xrl_surrogate.c"
gcc $CF -ftrivial-auto-var-init=pattern -shared -o $W/liborc_pattern.so $SRC -lm 2>/dev/null
gcc $CF -ftrivial-auto-var-init=zero -shared -o $W/liborc_zero.so $SRC -lm 2>/dev/null
gcc $CF -fsanitize=address,undefined -fno-omit-frame-pointer -shared -o $W/liborc_asan.so $SRC -lm 2>/dev/null
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so); STDCPP=$(gcc -print-file-name=libstdc++.so.6)
python $ROOT/tools/oracle_digests.py default > $W/o_default.json
python $ROOT/tools/oracle_digests.py $W/liborc_pattern.so > $W/o_pattern.json
python $ROOT/tools/oracle_digests.py $W/liborc_zero.so > $W/o_zero.json
cmp -s $W/o_default.json $W/o_pattern.json && cmp -s $W/o_default.json $W/o_zero.json && echo "pass 1 (auto-var-init pattern / zero): digests identical" || echo "pass 1: DIGESTS DIFFER"
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0 python $ROOT/tools/oracle_digests.py $W/liborc_asan.so > $W/o_asan.json 2> $W/o_asan.err \
  && cmp -s $W/o_default.json $W/o_asan.json && ! grep -q "ERROR: AddressSanitizer\|runtime error" $W/o_asan.err && echo "pass 2 (oracle ASan + UBSan): clean" || echo "pass 2: REPORTS in $W/o_asan.err"
cp $ROOT/xmimsim_b200/csrc/* $W/src/
sed -i 's|if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: xmb_main_msim has no CPU fallback"); return 0; }|/* dry run */|' $W/src/history.cu
NV="/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -I$ROOT/include -I$W/src -Xcompiler -fPIC,-fopenmp,-O1,-g,-fsanitize=address,-fno-omit-frame-pointer"
for f in $W/src/*.cu; do $NV -c $f -o $W/obj/$(basename $f).o & done
for f in $W/src/*.cpp; do g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -fopenmp -std=c++17 -I$ROOT/include -I$W/src -c $f -o $W/obj/$(basename $f).o & done
gcc -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -fopenmp -std=gnu99 -I$ROOT/include -I$W/src -c $W/src/xrl_surrogate.c -o $W/obj/xrl_surrogate.c.o &
wait
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o $W/libxmimsim_b200.so $W/obj/*.o -Xcompiler -fopenmp,-fsanitize=address -lgomp -ldl -cudart static
ln -sf libxmimsim_b200.so $W/xmimsim-cl.so
cd $ROOT
XMIMSIM_B200_LIB=$W/libxmimsim_b200.so LD_PRELOAD="$ASAN $STDCPP" ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/ -q -m "not gpu" -p no:cacheprovider > $W/tests_asan.log 2>&1 \
  && ! grep -q "ERROR: AddressSanitizer" $W/tests_asan.log && echo "pass 3 (product host code under ASan, CPU suite): $(tail -1 $W/tests_asan.log)" || echo "pass 3: REPORTS in $W/tests_asan.log"
#  4. glibc heap perturbation (MALLOC_PERTURB_: malloc'ed and freed memory filled with a byte pattern): oracle digests must
#     not move and the CPU suite must pass -- a read of uninitialised heap memory on the host would change them.
python $ROOT/tools/oracle_digests.py default > $W/o_plain.json
MALLOC_PERTURB_=165 python $ROOT/tools/oracle_digests.py default > $W/o_perturb.json
cmp -s $W/o_plain.json $W/o_perturb.json && MALLOC_PERTURB_=165 python -m pytest tests/ -q -m "not gpu" -p no:cacheprovider > $W/tests_perturb.log 2>&1 \
  && echo "pass 4 (MALLOC_PERTURB_): digests identical, $(tail -1 $W/tests_perturb.log)" || echo "pass 4: DIFFERENCE, see $W"
