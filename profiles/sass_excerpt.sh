#!/bin/bash
# SASS evidence of the Blackwell-native instructions in the shipped library (run here, no GPU needed):
#   UBLKCP = cp.async.bulk, SYNCS = mbarrier, IDP.4A = dp4a, UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, USETMAXREG = setmaxnreg
LIB=${1:-fast-llama_b200/libfastllama_b200.so}
cuobjdump -sass "$LIB" | awk '
/Function : / { fn = $3 }
{ for (i = 1; i <= NF; ++i) { t = $i; if (t ~ /^(UBLKCP|SYNCS|IDP\.4A|UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|USETMAXREG|UTCBAR|HMMA|IMMA)/) { split(t, a, "."); key = a[1]; if (key == "IDP") key = "IDP.4A"; n[fn "\t" key]++ } } }
END { for (k in n) print n[k] "\t" k }' | sort -k2,2 -k3,3 | awk -F'\t' '{ printf "%-120s %-12s %6d\n", $2, $3, $1 }'
