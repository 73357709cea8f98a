#!/bin/bash
# static SASS instruction count per kernel of libqugar_b200.so (I-cache footprint proxy)
cuobjdump -sass ${1:-qugar_b200/libqugar_b200.so} | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -n | tail -${2:-14}
