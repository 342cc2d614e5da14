#!/bin/bash
# SASS evidence for the tcgen05 / TMEM / TMA kernels: per kernel, the count of every tensor-core / tensor-memory / bulk-copy mnemonic and the
# first lines that carry them (cuobjdump -sass of the in-tree library; CPU only).
LIB=mpres-blas_b200/libmpres_b200.so
OUT=${1:-profiles/r02_sass_tcgen05_excerpt.txt}
cuobjdump -sass $LIB > /tmp/mpres_sass.txt
{
  echo "# cuobjdump -sass $LIB ($(date -u +%F)); sm_100a.  Mnemonics: UTCIMMA = tcgen05.mma kind::i8, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA),"
  echo "# UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk (1-D TMA copy), SYNCS = mbarrier ops, IMMA = legacy mma.sync int8, IDP = dp4a"
  awk '/Function : /{fn=$3} /UTCIMMA|LDTM|STTM|UTMALDG|UTCBAR|UTCATOM|UBLKCP|IMMA\.|SYNCS|IDP\./{split($0,a," "); for(i in a){ if (a[i] ~ /^(UTCIMMA|LDTM|STTM|UTMALDG|UTCBAR|UBLKCP|IMMA|SYNCS|IDP)/) {split(a[i],b,"."); cnt[fn" "b[1]]++}}} END{for(k in cnt) print cnt[k], k}' /tmp/mpres_sass.txt | sort -k2,2 -k3,3 | awk '{printf "%-8s %-10s %s\n", $1, $3, $2}'
  echo
  for k in k_small_umma_p k_limb_umma; do
    echo "## first tensor-core / TMEM / TMA instructions of the first instantiation of $k"
    awk -v k="$k" '/Function : /{on = index($3, k) > 0 ? on + 1 : 0} on == 1 && /UTCIMMA|LDTM|STTM|UTMALDG|UTCBAR/ {print}' /tmp/mpres_sass.txt | head -24
    echo
  done
} > $OUT
wc -l $OUT
