#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma ->
# UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG, legacy mma.sync -> HMMA, cp.async -> LDGSTS.
#   bash tests/gpu_checks/sass_opcounts.sh > profiles/r02_sass_opcounts.txt      (no GPU needed)
SO=${1:-univl_b200/libunivl_b200.so}
echo "# cuobjdump -sass $SO  (sm_100a) — mnemonic counts per kernel"
echo "# columns: UTCHMMA(all forms) UTMALDG UTMASTG UTMAREDG LDTM UTCBAR HMMA LDGSTS MUFU kernel"
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
  { if ($0 ~ /UTCHMMA/) c["mma"]++; if ($0 ~ /UTMALDG/) c["ldg"]++; if ($0 ~ /UTMASTG/) c["stg"]++;
    if ($0 ~ /UTMAREDG/) c["red"]++; if ($0 ~ /LDTM/) c["ldtm"]++; if ($0 ~ /UTCBAR/) c["bar"]++;
    if ($0 ~ / HMMA/) c["hmma"]++; if ($0 ~ /LDGSTS/) c["ldgsts"]++; if ($0 ~ /MUFU/) c["mufu"]++ }
  function emit() { printf "%4d %4d %4d %4d %4d %4d %4d %4d %4d  %s\n", c["mma"], c["ldg"], c["stg"], c["red"], c["ldtm"], c["bar"], c["hmma"], c["ldgsts"], c["mufu"], name }
  END { if (name != "") emit() }' | while read a b c d e f g h i n; do echo "$a $b $c $d $e $f $g $h $i $(echo $n | c++filt | cut -c1-110)"; done | sort -k10
