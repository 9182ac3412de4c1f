#!/bin/bash
# Per-kernel SASS evidence of the Blackwell-native path: tcgen05.mma (UTCHMMA / UTCQMMA...), TMA loads / stores
# (UTMALDG / UTMASTG), TMEM loads / stores (LDTM / STTM), 256-bit stores, vector reductions.
# usage: tools/sass_counts.sh [lib] > profiles/r02_sass_counts.txt
LIB=${1:-gcbfplus_b200/lib/libgcbf_b200.so}
echo "# cuobjdump -sass $LIB | per-function counts of UTCHMMA UTMALDG UTMASTG LDTM STTM STG.E.ENL2.256 REDG.*F32x4 UCGABAR"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
/UTCHMMA/ { a[fn]++ } /UTMALDG/ { b[fn]++ } /UTMASTG/ { c[fn]++ } /LDTM/ { d[fn]++ } /STTM/ { e[fn]++ }
/STG.E.ENL2.256/ { f[fn]++ } /REDG.*F32x4/ { g[fn]++ } /UCGABAR/ { h[fn]++ }
{ seen[fn]=1 }
END { for (k in seen) if (a[k]+b[k]+c[k]+d[k]+e[k]+f[k]+g[k]+h[k] > 0)
        printf "%6d %6d %6d %5d %5d %6d %6d %5d  %s\n", a[k], b[k], c[k], d[k], e[k], f[k], g[k], h[k], k }' | sort -k9 | \
  (echo "UTCHMMA UTMALDG UTMASTG  LDTM  STTM STG256 REDGx4 UCGABAR kernel"; cat) | c++filt
echo "# totals"
for m in UTCHMMA UTMALDG UTMASTG LDTM STTM "STG.E.ENL2.256" "REDG.*F32x4" UCGABAR; do
  printf "%-16s %d\n" "$m" "$(cuobjdump -sass "$LIB" | grep -cE "$m")"
done
