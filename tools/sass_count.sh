#!/bin/bash
# Development aid: SASS instruction mix per kernel of libpbd_b200.so (FFMA/FMUL/FADD counts show how much FMA contraction the
# explicit fmaf() calls recover under -fmad=false; FCHK/CALL = IEEE division slow paths).
SO=${1:-$(dirname "$0")/../positionbaseddynamics_b200/libpbd_b200.so}
cuobjdump -sass "$SO" | awk '
/Function :/ {name=$3}
/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ {
  op=$2; if (op ~ /^@/) op=$3;
  cnt[name]++;
  if (op ~ /^FFMA/) ffma[name]++; if (op ~ /^FMUL/) fmul[name]++; if (op ~ /^FADD/) fadd[name]++;
  if (op ~ /^CALL|^FCHK/) call[name]++; if (op ~ /^MUFU/) mufu[name]++;
}
END {for (n in cnt) printf "%6d ffma=%4d fmul=%4d fadd=%4d mufu=%3d call/fchk=%3d %s\n", cnt[n], ffma[n], fmul[n], fadd[n], mufu[n], call[n], n}' \
 | while read -r a b c d e f name; do echo "$a $b $c $d $e $f $(echo "$name" | c++filt | sed 's/pbdk:://g' | cut -c1-110)"; done | sort -k7
