#!/bin/bash
# Per-kernel counts of the Blackwell-only SASS mnemonics in the shipped library (tcgen05 MMA = UTCHMMA, TMEM load = LDTM,
# bulk TMA = UBLKCP, tensor-map TMA = UTMALDG, tcgen05.commit = UTCBAR, setmaxnreg = USETMAXREG) and of legacy HMMA (mma.sync).
#   usage: tools/dev/sass_proof.sh > profiles/rNN_sass_tc.txt
cd "$(dirname "$0")/../.."
echo "# cuobjdump -sass hvpr_b200/libhvpr_b200.so | mnemonic counts per kernel ($(date -u +%FT%TZ), nvcc $(nvcc --version | grep -o 'V[0-9.]*'))"
cuobjdump -sass hvpr_b200/libhvpr_b200.so | awk '
/Function :/ {fn=$3; next}
{ for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|UTCBAR|UTCCP|USETMAXREG|HMMA|SYNCS|REDUX|ATOMS|RED\.|ATOMG)/) { sub(/;$/,"",$i); c[fn"\t"$i]++ } }
END { for (k in c) print k"\t"c[k] }' | sort | c++filt
