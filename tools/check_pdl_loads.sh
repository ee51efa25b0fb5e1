#!/bin/bash
# For every kernel that contains a programmatic-dependent-launch wait (griddepcontrol.wait -> SASS ACQBULK), count the
# global loads that ptxas placed in front of the first ACQBULK.  Loads of data the programmatic primary writes must not
# be there: an invariant (__ldg / ld.global.nc) load may legally be hoisted above the wait, and was once
# (DESIGN.md 5a).  Streams that are complete before the primary starts (z, the ReLU gradient) are expected to show up.
# Usage: tools/check_pdl_loads.sh   (no GPU needed)
cd "$(dirname "$0")/../wt-pse-code_b200/csrc" || exit 1
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o /tmp/_pdl_check.cubin "$f" 2>/dev/null || continue
  cuobjdump -sass /tmp/_pdl_check.cubin | awk -v F="$f" '
    /Function :/ { fn = $3; n = 0; nc = 0; has = 0 }
    /LDG/ { if (!has) { n++; if ($0 ~ /CONSTANT/) nc++ } }
    /ACQBULK/ { if (!has) { has = 1; printf "%-28s %-60s loads before wait: %3d (invariant: %d)\n", F, substr(fn, length(fn) - 59), n, nc } }'
done
rm -f /tmp/_pdl_check.cubin
