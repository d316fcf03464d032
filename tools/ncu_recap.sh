set -u
OUT=/tmp/hsncu; TXT=gpurun_out; mkdir -p $OUT $TXT
cap() { local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $OUT/r01_$name "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$?"
  { echo "# ncu --set full --clock-control none -k regex:$rx -c 1 $*"; python tools/ncu_summary.py $OUT/r01_$name.ncu-rep 30;
    echo "-- segments between barriers"; python tools/ncu_segments.py /tmp/ncu_source.csv; } > $TXT/r01_${name}_ncu.txt 2>&1
}
cap k3_lagcov 'lagcov_mma_kernel' python tools/prof_mvar.py 599 1
cap k5_mma 'transfer_mma_kernel' python tools/prof_mvar.py 599 1
python bench.py --no-cpu --steps 2 --warmup 3 > /tmp/b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v2.csv python bench.py --no-cpu --steps 2 --warmup 3 > /tmp/ncu_b.log 2>&1
ncu -i $OUT/r01_k5_mma.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; d=dict(zip(h,r[2]))
for k in h:
    if any(t in k for t in ['dmma_cycles_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__cycles_elapsed.max','pipe_fp64_cycles_active.avg.pct']) and 'per_second' not in k and 'pct_of_peak_sustained_elapsed' not in k: print(k, d[k])
" > gpurun_out/r01_k5_raw.txt
cat gpurun_out/r01_k5_raw.txt
