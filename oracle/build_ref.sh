#!/bin/sh
# oracle/build_ref.sh -- TEST INFRASTRUCTURE: compiles the parts of the reference that build from their own sources with
# gcc alone (oracle/ref_shim/README.md) into oracle/_ref/libxmi_ref.so.  Reads /root/reference, writes only oracle/_ref/.
# The Fortran path (history loop, detector response, CPU solid angle) cannot be built here: no Fortran compiler, no
# xraylib / HDF5 / GLib.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${XMB_REFERENCE_DIR:-/root/reference}
OUT="$HERE/_ref"
[ -f "$REF/src/xmi_kernels.cl" ] || { echo "build_ref.sh: $REF not present, keeping the prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
sed -e 's/(const float3) *(/make_float3(/g' "$REF/src/xmi_kernels.cl" > "$OUT/xmi_kernels_cl.inc"
CC=$(test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)
CXX=$(test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$REF/include" -c "$REF/src/xmi_spline.c" -o "$OUT/xmi_spline.o"
$CXX -O2 -fPIC -fopenmp -std=c++14 -Wno-narrowing -I"$HERE/ref_shim" -I"$OUT" -I"$REF/src/Random123" -I"$REF/include" \
     -shared -o "$OUT/libxmi_ref.so" "$HERE/ref_driver.cpp" "$OUT/xmi_spline.o" -lm
rm -f "$OUT/xmi_spline.o"
echo "built $OUT/libxmi_ref.so"
