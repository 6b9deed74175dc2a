#!/bin/bash
# One gpurun call: bench lines (both arms, every config), GPU parity tests, ncu launch list + full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [steps...]'
TAG=${1:-r02}; shift
STEPS=${*:-bench tests launches ncu_m2}
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
for s in $STEPS; do
  t0=$(date +%s)
  case $s in
    bench)
      timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
      timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err ;;
    bench_c3)
      timeout 400 python bench.py --config c3 --steps 3 --warmup 2 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
      timeout 300 python bench.py --impl reference --config c3 --steps 2 --warmup 1 > $O/${TAG}_bench_c3_reference.json 2> $O/${TAG}_bench_c3_reference.err ;;
    bench_c4)
      timeout 500 python bench.py --config c4 --steps 3 --warmup 2 --c4-gb ${C4GB:-4} > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err
      timeout 300 python bench.py --impl reference --config c4 --steps 2 --warmup 1 --c4-gb ${C4GB:-4} > $O/${TAG}_bench_c4_reference.json 2> $O/${TAG}_bench_c4_reference.err ;;
    bench_c5)
      timeout 500 python bench.py --config c5 --steps 2 --warmup 1 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
      timeout 300 python bench.py --impl reference --config c5 --steps 1 --warmup 1 > $O/${TAG}_bench_c5_reference.json 2> $O/${TAG}_bench_c5_reference.err ;;
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log ;;
    launches)
      timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches_bench_m2.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/${TAG}_launches_bench.log 2>&1 ;;
    ncu_m2)
      timeout 400 $NCU --set full --import-source on -k regex:'k_lz_scan|k_lz_walk|k_lz_emit|k_suffix_sort' -c 7 -f -o $O/${TAG}_m2 \
        python tools/prof_step.py --units 10000 --steps 1 > $O/${TAG}_ncu_m2.log 2>&1 ;;
    ncu_cm)
      timeout 300 $NCU --set full --import-source on -k regex:k_cm_encode -c 1 -f -o $O/${TAG}_cm_m3 \
        python tools/prof_step.py --method 3 --units 1525 --steps 1 > $O/${TAG}_ncu_cm_m3.log 2>&1
      timeout 300 $NCU --set full --import-source on -k regex:k_cm_encode -c 1 -f -o $O/${TAG}_cm_m5 \
        python tools/prof_step.py --method 5 --units 256 --steps 1 > $O/${TAG}_ncu_cm_m5.log 2>&1 ;;
    smoke)
      timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log ;;
    *) eval "$s" ;;
  esac
  echo "[$s] rc=$? $(( $(date +%s) - t0 )) s"
done
