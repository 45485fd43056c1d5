#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove Blackwell-native paths (B200_PROFILING.md):
# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier ops.
# usage: tools/sass_grep.sh > profiles/r2_sass_grep.txt
so=binius_b200/libbinius_b200.so
echo "# cuobjdump -sass $so ($(date -u +%F)): kernel, UTCIMMA, LDTM, UBLKCP, SYNCS, LDS, total instructions"
cuobjdump -sass $so | awk '
/Function :/ { if (name != "") print name, u, l, b, s, lds, n; name=$3; u=l=b=s=lds=n=0; next }
/^[ \t]+\/\*[0-9a-f]+\*\// { n++; if ($0 ~ /UTCIMMA/) u++; if ($0 ~ /LDTM/) l++; if ($0 ~ /UBLKCP/) b++; if ($0 ~ /SYNCS/) s++; if ($0 ~ / LDS/) lds++ }
END { print name, u, l, b, s, lds, n }' | while read name rest; do echo "$(echo $name | c++filt | sed 's/(.*//') $rest"; done | sort
