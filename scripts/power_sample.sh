#!/bin/bash
# usage: scripts/power_sample.sh <label> <command...> : runs the command while sampling board power / SM clock every 100 ms
label=$1; shift
nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader,nounits -lms 100 > /tmp/power_$label.csv &
pid=$!
"$@"
kill $pid
awk -F, -v l=$label 'NR>5 {p+=$1; c+=$2; n++} END {printf "%s: mean power %.0f W, mean SM clock %.0f MHz over %d samples\n", l, p/n, c/n, n}' /tmp/power_$label.csv
