#!/usr/bin/env bash
# Round-2 ncu captures (run under gpurun, one GPU).  Every target first runs plain (must exit 0), then under ncu.
set -u
O=gpurun_out
cap() {  # name, kernel regex, skip, target args...
  local name=$1 rx=$2 skip=$3; shift 3
  python tools/ncu_targets.py "$@" > $O/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r2_$name python tools/ncu_targets.py "$@" > $O/ncu_$name.log 2>&1 &&
  ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > $O/r2_$name.raw.csv 2>/dev/null &&
  ncu -i /tmp/r2_$name.ncu-rep --page source --csv > $O/r2_$name.source.csv 2>/dev/null
  gzip -f $O/r2_$name.source.csv
}
# launch list of one sweep step with DRAM bytes (3 steps run: skip the first two steps' launches = 2 x 23)
python tools/ncu_targets.py sweep > $O/plain_sweep.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 46 -c 23 --csv --log-file $O/r2_sweep_launches.csv python tools/ncu_targets.py sweep > $O/ncu_sweep.log 2>&1
cap grp5_pw26 k_synth_group 2 win 9 26
cap grp7_pw26 k_synth_group 2 win 10 26
cap grp4_pw26 k_synth_group 2 win 6 26
cap build31 k_table_build_u 2 win 10 26
cap bank7_dw24 k_synth_bank 2 bank7_dw24
cap bank7_dds48 k_synth_bank 2 bank7_dds48
cap build_inq32 k_table_build_inq_u 2 bank7_dds48
cap taylor k_direct_taylor 2 cfg4
cap direct_window k_direct_window 2 direct 10 20 40
ls -la $O/r2_*
