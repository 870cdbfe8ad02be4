#!/bin/bash
# ncu captures behind profiles/r02_*: the launch list of the default bench command and one
# `--set full` capture of the blend kernel per workload (DRAM bytes per launch -> profiles/r02_roofline_traffic.json)
#   gpurun -- bash tools/ncu_round2.sh
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > $O/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $B > $O/r02_ncu_launches.log 2>&1
cap() {   # name, kernel regex, bench args...
  local name=$1 k=$2; shift 2
  local cmd="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras $*"
  $cmd > $O/r02_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 2 -f -o $O/r02_$name $cmd > $O/r02_ncu_$name.log 2>&1
  ncu -i $O/r02_$name.ncu-rep --page raw --csv > $O/r02_$name.raw.csv 2>/dev/null
  # the reports themselves are big (the merge back is limited to 64 MiB): only config 3's is kept
  [ "$name" = cfg3 ] || rm -f $O/r02_$name.ncu-rep
}
cap cfg3 ttmlblend_group_kernel --config 3
cap cfg3_distinct ttmlblend_group_kernel --config 3 --distinct-cues
cap cfg5 ttmlblend_group_kernel --config 5
cap cfg2 ttmlblend_group_kernel --config 2
cap cfg1 ttmlblend_group_kernel --config 1
cap cfg4_rgba ttmlblend_group_kernel --config 4 --format RGBA
cap cfg4_ayuv ttmlblend_group_kernel --config 4 --format AYUV
ls -la $O/r02_*.ncu-rep | wc -l
