#!/bin/bash
# Instruction evidence for the hand-written sm_100a kernels: counts of the Blackwell-specific SASS mnemonics per kernel
# (all template instantiations of a kernel summed), from the objects `make` builds (cuobjdump -sass).
#   bash profiles/sass_histogram.sh > profiles/r02_sass_histogram.txt        (build container, after `make`)
set -e
cd "$(dirname "$0")/.."
printf "%-34s %-28s %s\n" "kernel (instantiations summed)" "SASS mnemonic" "count"
for obj in build/gemm_tc.o build/gru_ts.o build/gru_stream.o build/frontend.o build/ctc.o; do
  [ -f "$obj" ] || continue
  cuobjdump -sass "$obj" | c++filt | awk -v OBJ="$obj" '
    /Function :/ { fn=$0; sub(/.*Function : /,"",fn); sub(/^void /,"",fn); sub(/<.*/,"",fn); sub(/\(.*/,"",fn); sub(/.*::/,"",fn) }
    { for (i=1;i<=NF;i++) { t=$i; sub(/;$/,"",t);
        if (t ~ /^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCCP|HMMA|LDSM|MEMBAR|UCGABAR_ARV|REDG|ATOMG)(\.|$)/) c[fn" "t]++ } }
    END { for (k in c) print c[k], k }' | sort -k2,2 -k1,1nr | awk '{printf "%-34s %-28s %6d\n", $2, $3, $1}'
done
