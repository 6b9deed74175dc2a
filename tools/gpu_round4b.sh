#!/bin/bash
# Second short gpurun call of the round: stage times of the shipped library and of the tuning builds beside it
# (libzqb200_<variant>.so, see profiles/README.md r04b), then one ncu --set full capture with source counters of the
# three large kernels of the -m2 step (read back here with tools/ncu_lines.py).
# usage: gpurun --timeout 420 -- 'bash tools/gpu_round4b.sh r04b [noncu]'
export TAG=${1:-r04b}
O=gpurun_out; mkdir -p $O
P=$PWD/zpaqfranz_b200
t00=$(date +%s)
for f in $P/libzqb200.so $P/libzqb200_*.so; do
  v=$(basename $f .so); v=${v#libzqb200}
  ZQ_LIB=$f timeout 120 python tools/quick_bench.py --units 10000 --steps 3 > $O/${TAG}_quick$v.log 2>&1
  echo "quick$v rc=$? $(tail -2 $O/${TAG}_quick$v.log | tr '\n' ' ' | cut -c1-330)"
done
echo "[quick] $(( $(date +%s) - t00 )) s"
[ "$2" = "noncu" ] && { echo "[all] $(( $(date +%s) - t00 )) s"; exit 0; }
t0=$(date +%s)
timeout 240 ncu --clock-control none --set full --import-source on -k regex:'k_suffix_sort16|k_lz_scan' -c 3 -f -o $O/${TAG}_m2 \
  python tools/prof_step.py --units 10000 --steps 1 > $O/${TAG}_ncu_m2.log 2>&1
echo "[ncu] rc=$? $(( $(date +%s) - t0 )) s  $(ls -la $O/${TAG}_m2.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
echo "[all] $(( $(date +%s) - t00 )) s"
