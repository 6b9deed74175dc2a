#!/bin/bash
# One gpurun call for the last round-2 measurements, most important first (the box time is short):
#   1. stage times of the -m2 step with the suffix-sort variants built beside the shipped library (tuning builds,
#      ZQ_LIB): shipped (dense alphabet + bin-limited sort network + tie runs dealt to lanes), _r03a (round-2 code
#      as measured in r03a), _raw (raw bytes + dealt ties), _densefull (dense, full network), _tie512
#   2. the -m gpu parity suite on the shipped library
#   3. bench lines of both arms (and of the fastest variant if that is not the shipped one)
#   4. smoke(), ncu launch list
# usage: gpurun --timeout 900 -- 'bash tools/gpu_round4.sh r04a'
export TAG=${1:-r04a}
O=gpurun_out; mkdir -p $O
P=$PWD/zpaqfranz_b200
t00=$(date +%s)
for v in "" _r03a _raw _densefull _tie512; do
  [ -f $P/libzqb200$v.so ] || continue
  ZQ_LIB=$P/libzqb200$v.so timeout 150 python tools/quick_bench.py --units 10000 --steps 3 > $O/${TAG}_quick$v.log 2>&1
  echo "quick$v rc=$? $(tail -2 $O/${TAG}_quick$v.log | tr '\n' ' ' | cut -c1-400)"
done
echo "[quick] $(( $(date +%s) - t00 )) s"
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
echo "[tests] $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json
echo "[bench] rc=$? $(( $(date +%s) - t0 )) s"
# the fastest variant, if it is not the shipped library: its bench line too (parity of all 10 000 blocks)
best=$(python - <<'EOF'
import glob, re, os
best, bt = "", 1e9
for f in glob.glob("gpurun_out/%s_quick*.log" % os.environ.get("TAG", "r04a")):
    s = open(f).read()
    m = re.search(r"'total': ([0-9.]+)", s)
    if m and "parity ok" in s and float(m.group(1)) < bt:
        bt, best = float(m.group(1)), re.search(r"_quick(.*)\.log", f).group(1)
print(best)
EOF
)
echo "fastest variant: '${best}'"
if [ -n "$best" ]; then
  t0=$(date +%s)
  ZQ_LIB=$P/libzqb200$best.so timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench$best.json 2> $O/${TAG}_bench$best.err
  echo "[bench$best] rc=$? $(( $(date +%s) - t0 )) s"
fi
t0=$(date +%s)
timeout 250 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
echo "[bench reference] rc=$? $(( $(date +%s) - t0 )) s"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
t0=$(date +%s)
timeout 200 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches_bench_m2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/${TAG}_launches_bench.log 2>&1
echo "[launches] rc=$? $(( $(date +%s) - t0 )) s"
echo "[all] $(( $(date +%s) - t00 )) s"
