#!/bin/bash
# SASS evidence per kernel of the shipped library: opcode histogram (top 40) + the counts of the instructions that prove
# the tcgen05 / TMEM / bulk-copy path and of local-memory traffic.  Usage: tools/sass_histogram.sh [round tag]
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r02}
LIB=sp-nerf_b200/lib/libspnerf_sm100a.so
OUT=profiles/sass_${TAG}.txt
TMP=$(mktemp)
cuobjdump -sass $LIB > $TMP
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%F)), per kernel: instruction count, key opcodes, top opcodes"
  awk '
    /Function :/ { fn=$3; next }
    /^[ \t]+\/\*[0-9a-f]+\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); n[fn]++; c[fn,op]++; ops[op]=1 }
    END {
      for (f in n) {
        printf "\n== %s: %d instructions\n   key:", f, n[f]
        nk = split("UTCHMMA UTCBAR UTCATOMSWS LDTM STTM UBLKCP UTMALDG SYNCS ELECT MUFU.SIN MUFU.SQRT MUFU.EX2 LDL STL LDG STG ATOMG RED", k, " ")
        for (i = 1; i <= nk; i++) { t=0; for (o in ops) if (index(o, k[i]) == 1) t += c[f,o]; if (t) printf " %s=%d", k[i], t }
        printf "\n   top:"
        m=0; for (o in ops) if (c[f,o]) { m++; cnt[m]=c[f,o]; nam[m]=o }
        for (i = 1; i <= m; i++) for (j = i + 1; j <= m; j++) if (cnt[j] > cnt[i]) { t=cnt[i]; cnt[i]=cnt[j]; cnt[j]=t; s=nam[i]; nam[i]=nam[j]; nam[j]=s }
        for (i = 1; i <= m && i <= 40; i++) printf " %s:%d", nam[i], cnt[i]
        printf "\n"
      }
    }' $TMP | c++filt
} > $OUT
rm -f $TMP
echo wrote $OUT
