#!/bin/bash
# Re-capture selected kernels after a change: bash tools/ncu_recap.sh name:regex:driver ...   (driver = mvar | frontend)
set -u
OUT=/tmp/hsncu; TXT=gpurun_out; mkdir -p $OUT $TXT
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; drv=${rest#*:}
  if [ "$drv" = "frontend" ]; then cmd="python tools/bench_frontend.py 600"; else cmd="python tools/prof_mvar.py 599 1"; fi
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $OUT/r01_$name $cmd > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$?"
  { echo "# ncu --set full --clock-control none -k regex:$rx -c 1 $cmd"; python tools/ncu_summary.py $OUT/r01_$name.ncu-rep 30;
    echo "-- segments between barriers"; python tools/ncu_segments.py /tmp/ncu_source.csv; } > $TXT/r01_${name}_ncu.txt 2>&1
done
