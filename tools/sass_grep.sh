# tools/sass_grep.sh -- per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA in the shipped library
# (runs on the build container: cuobjdump needs no GPU).  Output: profiles/r2_sass_grep.txt
OUT=${1:-profiles/r2_sass_grep.txt}
LIB=yolo_quantization_b200/libyq_b200.so
{
echo "# cuobjdump -sass $LIB (sm_100a only: $(cuobjdump -lelf $LIB | grep -c sm_100a) cubins, $(cuobjdump -lelf $LIB | grep -vc sm_100a) others)"
echo "# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit"
echo "# kernel | UTCIMMA | LDTM | UTMALDG | UTMASTG | UBLKCP | UTCBAR"
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") print name, a, b, c, d, e, f; name=$3; a=b=c=d=e=f=0 }
/UTCIMMA/ {a++} /LDTM/ {b++} /UTMALDG/ {c++} /UTMASTG/ {d++} /UBLKCP/ {e++} /UTCBAR/ {f++}
END { print name, a, b, c, d, e, f }' | while read n a b c d e f; do
  if [ "$a$b$c$d$e" != "00000" ]; then echo "$(echo $n | c++filt | sed 's/(anonymous namespace):://; s/(CUtensorMap_st.*//; s/void //') | $a | $b | $c | $d | $e | $f"; fi
done | sort
echo "# totals: UTCIMMA $(cuobjdump -sass $LIB | grep -c UTCIMMA), LDTM $(cuobjdump -sass $LIB | grep -c LDTM), UTMALDG $(cuobjdump -sass $LIB | grep -c UTMALDG), UTMASTG $(cuobjdump -sass $LIB | grep -c UTMASTG), UBLKCP $(cuobjdump -sass $LIB | grep -c UBLKCP)"
} > $OUT
wc -l $OUT; tail -3 $OUT
