#!/bin/bash
# Final gpurun call of round 2: the -m gpu parity suite on the shipped library, bench lines of both arms, smoke(),
# the ncu launch list of the bench command and one ncu --set full capture of the three large kernels.
# usage: gpurun --timeout 480 -- 'bash tools/gpu_round4f.sh r04f'
export TAG=${1:-r04f}
O=gpurun_out; mkdir -p $O
t00=$(date +%s)
timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
echo "[tests] $(( $(date +%s) - t00 )) s"
t0=$(date +%s)
timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "[bench] rc=$? $(( $(date +%s) - t0 )) s"; cut -c1-200 $O/${TAG}_bench.json
t0=$(date +%s)
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "[bench reference] rc=$? $(( $(date +%s) - t0 )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
t0=$(date +%s)
timeout 150 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches_bench_m2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/${TAG}_launches_bench.log 2>&1
echo "[launches] rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/${TAG}_bench_20steps.json 2> /dev/null; echo "[bench 20 steps] rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 200 ncu --clock-control none --set full -k regex:'k_suffix_sort16|k_lz_scan' -c 3 -f -o $O/${TAG}_m2 \
  python tools/prof_step.py --units 10000 --steps 1 > $O/${TAG}_ncu_m2.log 2>&1
echo "[ncu] rc=$? $(( $(date +%s) - t0 )) s"
echo "[all] $(( $(date +%s) - t00 )) s"
