#!/bin/bash
# usage: tools/ncu_csv.sh <out.csv> <kernel regex> <count> <cmd...>   (GPU box; keeps only the raw-page CSV, the .ncu-rep stays in /tmp)
out=$1; k=$2; c=$3; shift 3
rm -f /tmp/ncu_tmp.ncu-rep
ncu --set full --clock-control none -k regex:$k -c $c -o /tmp/ncu_tmp "$@" > /tmp/ncu_tmp.log 2>&1
ncu -i /tmp/ncu_tmp.ncu-rep --page raw --csv > $out 2>/dev/null
wc -c $out
