#!/bin/bash
# usage: sass_arith.sh <object/.so> <kernel-name-substring> : float/convert/control SASS of one kernel
obj=$1; pat=$2
cuobjdump -sass "$obj" | awk -v pat="$pat" '
/Function :/ { on = (index($0, pat) > 0); if (on) print $0 }
on && /\/\*[0-9a-f]{4}\*\// { line=$0; sub(/\/\* 0x[0-9a-f]+ \*\//, "", line); print line }' \
 | grep -E "Function|F[A-Z0-9]+\.|FADD|FMUL|FFMA|MUFU|I2F|F2I|FSET|FSEL|FMNMX|FCHK|BRA|ATOM|RED|STG|LDG|BAR|EXIT|I2FP|F2FP" 
