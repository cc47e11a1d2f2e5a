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
# xmi_output_raw2struct (the history mapping behind the XMSO writer) with its two array macros: lines 1368-1519 of
# src/xmi_data_structs.c, located by their first and last line so that a shifted file still extracts the same function
awk '/^#define ARRAY2D_FORTRAN/ {on = 1} on {print} on && /^#endif/ {exit}' "$REF/src/xmi_data_structs.c" > "$OUT/raw2struct.inc"
grep -q "xmi_output_raw2struct" "$OUT/raw2struct.inc"
{ echo '#define XMI_LINES_NO_CONFIG'; sed -e 's/#include "config.h"//' "$REF/src/xmi_lines.c"; } > "$OUT/xmi_lines_noconfig.c"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$REF/include" -I"$REF/src" -c "$OUT/xmi_lines_noconfig.c" -o "$OUT/xmi_lines.o"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$OUT" -I"$REF/include" -I"$REF/src" -c "$HERE/ref_raw2struct.c" -o "$OUT/ref_raw2struct.o"
$CXX -O2 -fPIC -fopenmp -std=c++14 -Wno-narrowing -I"$HERE/ref_shim" -I"$OUT" -I"$REF/src/Random123" -I"$REF/include" \
     -shared -o "$OUT/libxmi_ref.so" "$HERE/ref_driver.cpp" "$OUT/xmi_spline.o" "$OUT/xmi_lines.o" "$OUT/ref_raw2struct.o" -lm
rm -f "$OUT/xmi_spline.o" "$OUT/xmi_lines.o" "$OUT/ref_raw2struct.o" "$OUT/xmi_lines_noconfig.c"
echo "built $OUT/libxmi_ref.so"
