#!/bin/bash
# Instruction distance between consecutive UTCHMMA (tcgen05.mma) in each kernel of libvknet.so: a lean issue loop is ~5-8.
SO=${1:-video-k-net_b200/vknet/libvknet.so}
cuobjdump -sass "$SO" | awk '
/Function :/ {fn=$3; n=0; last=0}
/^\s+\/\*[0-9a-f]+\*\/\s/ { n++; if ($0 ~ /UTCHMMA/) { if (last>0) {gaps[fn]=gaps[fn] " " (n-last)}; last=n; cnt[fn]++ } }
END { for (f in cnt) print f, "MMAs:", cnt[f], "gaps:", gaps[f] }'
