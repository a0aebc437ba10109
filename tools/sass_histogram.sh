#!/bin/bash
# SASS evidence that the shipped library is Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# per-kernel counts of tcgen05 (UTC*MMA), TMEM (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP) and legacy (HMMA) opcodes.
#   bash tools/sass_histogram.sh > profiles/round2/sass_opcodes.txt
lib=${1:-millieye_b200/lib/libmillieye_b200.so}
echo "# cuobjdump -sass $lib ($(sha256sum $lib | cut -c1-16)), $(date -u +%FT%TZ)"
cuobjdump -sass "$lib" > /tmp/sass.txt
echo "## whole library"
for op in UTCHMMA UTCHMMA.2CTA LDTM STTM UTMALDG UTMALDG.2D UTMALDG.4D.IM2COL UTMASTG UTMAPF UTCBAR UBLKCP LDGSTS HGMMA SYNCS.ARRIVE; do
  printf "%-22s %6d\n" $op $(grep -c "$op" /tmp/sass.txt)
done
printf "%-22s %6d   (mma.sync / wmma: none expected)\n" "HMMA (legacy)" $(grep -E "(^|[^C])HMMA" /tmp/sass.txt | grep -vc UTCHMMA)
echo "## per kernel (tensor-core / TMEM / TMA opcodes)"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} /HMMA/ && !/UTCHMMA/ {e[name]++}
     END {for (k in a) printf "%s\n    UTCHMMA %d  LDTM %d  UTMALDG %d  UTMASTG %d  HMMA(legacy) %d\n", k, a[k], b[k], c[k], d[k], e[k]}' /tmp/sass.txt | c++filt | cut -c1-220
