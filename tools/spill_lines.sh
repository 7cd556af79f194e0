#!/bin/bash
# Attribute the local-memory instructions (LDL / STL) of one kernel to source lines (-lineinfo).
# Usage: tools/spill_lines.sh <object basename: mlp_fwd|mlp_bwd|...> <substring of the mangled kernel name>
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
cuobjdump -xelf $1 sp-nerf_b200/lib/libspnerf_sm100a.so > /dev/null 2>&1 || (cd $T && cuobjdump -xelf all $OLDPWD/sp-nerf_b200/lib/libspnerf_sm100a.so > /dev/null)
[ -f $1.sm_100a.cubin ] && mv $1.sm_100a.cubin $T/
nvdisasm --print-line-info $T/$1.sm_100a.cubin > $T/all.sass
python3 - "$T/all.sass" "$2" <<'PY'
import re, sys, collections
fn = None; line = None; cnt = collections.Counter()
for l in open(sys.argv[1]):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if fn and sys.argv[2] in fn and re.search(r'\b(LDL|STL)\b', l):
        cnt[(line, 'LDL' if 'LDL' in l else 'STL')] += 1
for k, v in sorted(cnt.items()): print(k[0][0], k[0][1], k[1], v)
PY
rm -rf $T
