#!/bin/bash
# usage: make_profile_md.sh <file.ncu-rep> <title> > profiles/<name>.md
rep=$1; shift
echo "# ncu summary: $*"
echo '```'
python tools/ncu_summary.py $rep 2>/dev/null
echo '```'
echo
echo "## executed warp-instructions by opcode"
echo '```'
python tools/ncu_sass_hist.py $rep 2>/dev/null | head -32
echo '```'
echo
echo "## by source line"
echo '```'
python tools/ncu_line_hist.py $rep 36 2>/dev/null | cut -c1-160
echo '```'
echo
echo "## top stall lines"
echo '```'
for r in stall_wait stall_long_sb stall_short_sb stall_barrier; do python tools/ncu_stall_lines.py $rep $r 6 2>/dev/null | cut -c1-150; done
echo '```'
