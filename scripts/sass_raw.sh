#!/bin/bash
# dump the SASS of the raw-gather kernel of a library: scratch/sass_raw.sh lib.so out.sass
cuobjdump -sass "$1" | awk '/Function : .*extract_tiledILb0ELb0ELi8ELi4ELi7ELb1E/{p=1} p&&/Function : /&&!/ELi7ELb1E/{p=0} p' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s+\/\*.*$//' > "$2"
wc -l "$2"
