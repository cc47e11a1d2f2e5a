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
# xmi_input_validate (src/xmi_data_structs.c:899-1255): what the reference's XMSI reader rejects
awk '/^XmiInputFlags xmi_input_validate/ {on = 1} on {print} on && /^}/ {exit}' "$REF/src/xmi_data_structs.c" > "$OUT/input_validate.inc"
grep -q "after_detector" "$OUT/input_validate.inc"
{ echo '#define XMI_LINES_NO_CONFIG'; sed -e 's/#include "config.h"//' "$REF/src/xmi_lines.c"; } > "$OUT/xmi_lines_noconfig.c"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$REF/include" -I"$REF/src" -c "$OUT/xmi_lines_noconfig.c" -o "$OUT/xmi_lines.o"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$OUT" -I"$REF/include" -I"$REF/src" -c "$HERE/ref_shim/ref_raw2struct.c" -o "$OUT/ref_raw2struct.o"
# the cache match rules: xmi_check_solid_angle_match (src/xmi_solid_angle.c:420-673) and xmi_check_escape_ratios_match
# (src/xmi_detector.c:143-172), each from its first line to the closing brace in column 0; CS_Total_Kissel comes from the
# surrogate provider object (third-party stand-in), as in the oracle
awk '/^int xmi_check_solid_angle_match/ {on = 1} on {print} on && /^}/ {exit}' "$REF/src/xmi_solid_angle.c" > "$OUT/solid_angle_match.inc"
awk '/^int xmi_check_escape_ratios_match/ {on = 1} on {print} on && /^}/ {exit}' "$REF/src/xmi_detector.c" > "$OUT/escape_ratios_match.inc"
awk '/^xmi_escape_ratios_options xmi_get_default_escape_ratios_options/ {on = 1} on {print} on && /^}/ {exit}' "$REF/src/xmi_detector.c" > "$OUT/default_escape_options.inc"
awk '/^static const xmi_main_options __default_main_options/ {on = 1} on {print} on && /^};/ {exit}' "$REF/src/xmi_data_structs.c" > "$OUT/default_main_options.inc"
grep -q "use_variance_reduction" "$OUT/default_main_options.inc" && grep -q "1990" "$OUT/default_escape_options.inc"
grep -q "XMI_IF_COMPARE_GEOMETRY2" "$OUT/solid_angle_match.inc" && grep -q "crystal_layers" "$OUT/escape_ratios_match.inc"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/ref_shim" -I"$OUT" -I"$REF/include" -I"$HERE/../include" -c "$HERE/ref_shim/ref_match.c" -o "$OUT/ref_match.o"
$CC -O2 -fPIC -std=gnu99 -I"$HERE/../include" -I"$HERE/../xmimsim_b200/csrc" -c "$HERE/../xmimsim_b200/csrc/xrl_surrogate.c" -o "$OUT/xrl_surrogate.o"
# struct layouts of include/xmimsim_b200.h against the reference's headers: _Static_asserts, a mismatch fails this build
$CC -O2 -fPIC -std=gnu11 -I"$HERE/ref_shim" -I"$REF/include" -I"$HERE/../include" -c "$HERE/ref_shim/ref_layout.c" -o "$OUT/ref_layout.o"
$CXX -O2 -fPIC -fopenmp -std=c++14 -Wno-narrowing -I"$HERE/ref_shim" -I"$OUT" -I"$REF/src/Random123" -I"$REF/include" \
     -shared -o "$OUT/libxmi_ref.so" "$HERE/ref_driver.cpp" "$OUT/xmi_spline.o" "$OUT/xmi_lines.o" "$OUT/ref_raw2struct.o" "$OUT/ref_match.o" "$OUT/xrl_surrogate.o" "$OUT/ref_layout.o" -lm
rm -f "$OUT/ref_layout.o" "$OUT/xmi_spline.o" "$OUT/xmi_lines.o" "$OUT/ref_raw2struct.o" "$OUT/ref_match.o" "$OUT/xrl_surrogate.o" "$OUT/xmi_lines_noconfig.c"
echo "built $OUT/libxmi_ref.so"
