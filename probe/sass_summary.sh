#!/bin/bash
# SASS opcode summary per translation unit of librtat_b200.so (what proves TMA / tcgen05 / DMMA are in the shipped code).
# usage: probe/sass_summary.sh > profiles/r02_sass_opcodes.txt
echo "# cuobjdump -sass build/obj/<unit>.o | opcode histogram of the instructions that matter (sm_100a), $(date -u +%F)"
for o in build/obj/*.o; do
  u=$(basename $o .o)
  cuobjdump -sass $o 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4,}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' > /tmp/ops_$u.txt
  total=$(wc -l < /tmp/ops_$u.txt)
  [ "$total" -eq 0 ] && continue
  echo "== $u.cu ($total SASS instructions)"
  grep -E '^(DMMA|UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTMACCTL|UBLKCP|SYNCS|LDGSTS|USETMAXREG|ELECT|UCGABAR|CCTL|LDG\.E\.(128|256|ENL2\.256)|STG\.E\.(128|256|ENL2\.256)|LDS\.128|STS\.128|MEMBAR|FENCE|REDUX|MUFU|F2FP|F2F|CCTLL|ACQBULK|UTCCP)' /tmp/ops_$u.txt | sort | uniq -c | sort -k1,1nr | awk '{printf "  %8d  %s\n", $1, $2}'
done
