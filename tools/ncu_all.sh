#!/bin/bash
# One ncu --set full capture per hot kernel (run on the GPU box, one GPU).  The reports (~100 MB together) stay in /tmp on the
# box; tools/ncu_summary.py + tools/ncu_segments.py turn each into the text summary gpurun_out/r01_<kernel>_ncu.txt that
# is committed under profiles/.
set -u
OUT=/tmp/hsncu
TXT=gpurun_out
mkdir -p $OUT $TXT
cap() {  # name regex driver...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $OUT/r01_$name "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$? $(grep -c '==PROF== Profiling' $OUT/ncu_$name.log) profiled"
  { echo "# ncu --set full --clock-control none -k regex:$rx -c 1 $*"; python tools/ncu_summary.py $OUT/r01_$name.ncu-rep 30;
    echo "-- segments between barriers"; python tools/ncu_segments.py /tmp/ncu_source.csv; } > $TXT/r01_${name}_ncu.txt 2>&1
}
cap k3_lagcov   'lagcov_kernel'          python tools/prof_mvar.py 599 1
cap k4_lwr1     'lwr1_kernel'            python tools/prof_mvar.py 599 1
cap k5_mma      'transfer_mma_kernel'    python tools/prof_mvar.py 599 1
cap fin         'dtf_finalize_kernel'    python tools/prof_mvar.py 599 1
cap k1_local    'iir_tile_local_kernel'  python tools/bench_frontend.py 600
cap k1_apply    'iir_tile_apply_kernel'  python tools/bench_frontend.py 600
cap k2_decimate 'fir_decimate_kernel'    python tools/bench_frontend.py 600
cap k6_psd      'mt_psd_kernel'          python tools/bench_frontend.py 600
ls -la $OUT/r01_*.ncu-rep $TXT/r01_*_ncu.txt
