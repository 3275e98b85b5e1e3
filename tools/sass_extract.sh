#!/bin/bash
# tools/sass_extract.sh -- SASS of the two hot kernels out of the built library, for profiles/: the full listing of
# block_sweep_kernel (tcgen05 / TMEM / TMA mnemonics visible), an instruction histogram of both kernels, and the
# tile sweep's single-gate variant (MODE 0) in full.  Run after `python rocquantum_b200/build.py`.
set -e
cd "$(dirname "$0")/.."
LIB=rocquantum_b200/lib/libhipStateVec.so
OUT=profiles
BS=$(cuobjdump -sass $LIB | grep -E "Function :.*block_sweep_kernel" | sed 's/.*Function : //')
TS=$(cuobjdump -sass $LIB | grep -E "Function :.*tile_sweep_small_cu.*Li0ELb0" | sed 's/.*Function : //')
TP=$(cuobjdump -sass $LIB | grep -E "Function :.*tile_sweep_large_cu.*Li2ELb0" | sed 's/.*Function : //')
cuobjdump -sass -fun "$BS" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' > $OUT/r02_sass_block_sweep_kernel.txt
cuobjdump -sass -fun "$TS" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' > $OUT/r02_sass_tile_sweep_kernel_mode0.txt
{
  echo "# SASS instruction histograms (cuobjdump -sass, sm_100a), round 2"
  for pair in "block_sweep_kernel:$BS" "tile_sweep_kernel<small, MODE 0> (one-gate sweeps):$TS" "tile_sweep_kernel<large, MODE 2> (window phases, complex64):$TP"; do
    name=${pair%%:*}; fun=${pair#*:}
    echo; echo "## $name"; echo; echo '```'
    cuobjdump -sass -fun "$fun" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#^\s+/\*[0-9a-f]+\*/\s+##; s#^@!?U?P[0-9T]+ ##' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn | head -45
    echo '```'
  done
} > $OUT/r02_sass_histograms.md
wc -l $OUT/r02_sass_*.txt $OUT/r02_sass_histograms.md
