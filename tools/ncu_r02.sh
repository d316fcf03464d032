#!/bin/bash
# Round-2 captures: one ncu --set full per changed / judged kernel (run on the GPU box, one GPU); text summaries go to
# gpurun_out/r02_<kernel>_ncu.txt (committed under profiles/), the launch list of bench.py to gpurun_out/r02_launches.csv.
set -u
OUT=/tmp/hsncu
TXT=gpurun_out
mkdir -p $OUT $TXT
cap() {  # name regex driver...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $OUT/r02_$name "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$? $(grep -c '==PROF== Profiling' $OUT/ncu_$name.log) profiled"
  { echo "# ncu --set full --clock-control none -k regex:$rx -c 1 $*"; python tools/ncu_summary.py $OUT/r02_$name.ncu-rep 30;
    echo "-- segments between barriers"; python tools/ncu_segments.py /tmp/ncu_source.csv; } > $TXT/r02_${name}_ncu.txt 2>&1
}
for k in "$@"; do
  case $k in
    k3)   cap k3_lagcov_tma 'lagcov_mma_kernel'     python tools/prof_mvar.py 599 1 ;;
    k3g)  cap k3_gemm_cfg5  'lagcov_gemm_kernel'    python tools/prof_cfg5.py ;;
    k4)   cap k4_lwr2       'lwr2_kernel'           python tools/prof_mvar.py 599 1 ;;
    k5)   cap k5_mma        'transfer_mma_kernel'   python tools/prof_mvar.py 599 1 ;;
    k5w)  cap k5_ws         'transfer_ws_kernel'    python tools/prof_mvar.py 599 1 ;;
    fin)  cap fin           'dtf_finalize_kernel'   python tools/prof_mvar.py 599 1 ;;
    k1)   cap k1_fused      'iir_tile_fused_kernel' python tools/bench_frontend.py 600 ;;
    k2)   cap k2_fir_mma    'fir_mma_kernel'   python tools/bench_frontend.py 1800 ;;
    k6)   cap k6_r16        'mt_psd_r16_kernel'     python tools/bench_frontend.py 600 ;;
    list) python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $TXT/r02_list_plain.log 2>&1 &&
          ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $TXT/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $TXT/r02_list_ncu.log 2>&1
          echo "list rc=$?" ;;
  esac
done
ls -la $TXT/r02_*_ncu.txt $TXT/r02_launches.csv 2>/dev/null
