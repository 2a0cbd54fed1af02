#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): run here, no GPU needed.
#   bash profiles/sass_summary.sh > profiles/r2_sass_summary.txt
so=${1:-continual-skeletons_b200/csrc/libcosk.so}
echo "# cuobjdump -sass $so : per kernel, count of UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG (TMA loads), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), HMMA (legacy mma.sync: must be 0)"
printf "%8s %13s %8s %6s %6s %7s %5s  %s\n" UTCHMMA UTCHMMA.2CTA UTMALDG LDTM STTM UTCBAR HMMA kernel
cuobjdump -sass "$so" 2>/dev/null | awk '
/Function : /{ if (name != "") out(); name=$3; a=b=c=d=e=f=g=0; next }
/UTCHMMA\.2CTA/{b++; next} /UTCHMMA/{a++} /UTMALDG/{c++} /LDTM/{d++} /STTM/{e++} /UTCBAR/{f++} / HMMA/{g++}
function out(){ if (a+b+c+d+e+f+g > 0) printf "%s %8d %13d %8d %6d %6d %7d %5d\n", name, a, b, c, d, e, f, g }
END{ out() }' | while read -r line; do n=$(echo "$line" | awk '{print $1}'); rest=$(echo "$line" | cut -d" " -f2-); printf "%s  %s\n" "$rest" "$(echo $n | c++filt | sed 's/cosk:://g; s/(int)//g; s/(bool)//g')"; done
