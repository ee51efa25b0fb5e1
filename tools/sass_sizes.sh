#!/bin/bash
# instruction count per kernel in the built library (code size matters for the single-CTA epilogues)
cuobjdump -sass "${1:-wt-pse-code_b200/libwtpse_b200.so}" 2>/dev/null | awk '
/Function :/ {name=$3}
/^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {cnt[name]++}
END {for (n in cnt) print cnt[n], n}' | sort -n | sed -E 's/_ZN5wtpse[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+//' | cut -c1-90
