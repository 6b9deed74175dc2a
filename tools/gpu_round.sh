#!/bin/bash
# One gpurun call: bench lines (both arms), GPU parity tests, ncu launch list + full captures.
# usage: gpurun --timeout 1000 -- 'bash tools/gpu_round.sh <tag> [steps...]'   (steps: bench tests launches ncu_m2 ncu_cm)
TAG=${1:-r01h}; shift
STEPS=${*:-bench tests launches ncu_m2 ncu_cm}
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
for s in $STEPS; do
  t0=$(date +%s)
  case $s in
    bench)
      timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
      timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err ;;
    tests)
      timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log ;;
    launches)
      timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches_bench_m2.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1 ;;
    ncu_m2)
      timeout 300 $NCU --set full --import-source on -k regex:'k_lz77_sa|k_suffix_sort' -c 2 -f -o $O/${TAG}_m2 \
        python tools/prof_step.py --units 2000 --steps 1 > $O/${TAG}_ncu_m2.log 2>&1 ;;
    ncu_cm)
      timeout 300 $NCU --set full --import-source on -k regex:k_cm_encode -c 1 -f -o $O/${TAG}_cm_bwt \
        python tools/prof_step.py --method 36,200,1 --units 296 --unit 16384 --steps 1 > $O/${TAG}_ncu_cm_bwt.log 2>&1
      timeout 300 $NCU --set full --import-source on -k regex:k_cm_encode -c 1 -f -o $O/${TAG}_cm_m5 \
        python tools/prof_step.py --method 5 --units 64 --unit 4096 --steps 1 > $O/${TAG}_ncu_cm_m5.log 2>&1 ;;
    ncu_bwt)
      timeout 300 $NCU --set full --import-source on -k regex:k_cm_encode -c 1 -f -o $O/${TAG}_cm_bwt \
        python tools/prof_step.py --method 36,200,1 --units 296 --unit 16384 --steps 1 > $O/${TAG}_ncu_cm_bwt.log 2>&1 ;;
    configs)
      timeout 500 python tools/bench_configs.py --c5-units 1200 > $O/${TAG}_configs_c3_c4_c5.json 2> $O/${TAG}_configs.err ;;
    tests_vm1)
      ZQ_CM_VM=1 timeout 700 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu_vm1.log 2>&1; tail -4 $O/${TAG}_pytest_gpu_vm1.log ;;
    tests_cm_vm0)
      ZQ_CM_VM=0 timeout 500 python -m pytest tests/test_gpu_compress.py tests/test_gpu_decode.py tests/test_gpu_segments.py -m gpu -q > $O/${TAG}_pytest_gpu_cm_vm0.log 2>&1; tail -4 $O/${TAG}_pytest_gpu_cm_vm0.log ;;
    tests_cm_vm2)
      ZQ_CM_VM=2 timeout 500 python -m pytest tests/test_gpu_compress.py tests/test_gpu_decode.py tests/test_gpu_segments.py -m gpu -q > $O/${TAG}_pytest_gpu_cm_vm2.log 2>&1; tail -4 $O/${TAG}_pytest_gpu_cm_vm2.log ;;
    cmtime_vm2)
      ZQ_CM_VM=2 timeout 400 python tools/bench_configs.py --skip c4 --c5-units 1200 > $O/${TAG}_configs_c3_c5_vm2.json 2> $O/${TAG}_configs_vm2.err ;;
    smoke)
      timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log ;;
    cmtime_vm1)
      ZQ_CM_VM=1 timeout 400 python tools/bench_configs.py --skip c4 --c5-units 1200 > $O/${TAG}_configs_c3_c5_vm1.json 2> $O/${TAG}_configs_vm1.err ;;
    cmtime)
      timeout 400 python tools/bench_configs.py --skip c4 --c5-units 1200 > $O/${TAG}_configs_c3_c5.json 2> $O/${TAG}_configs.err ;;
    cmtime_nopf)
      ZQ_CM_PREFETCH=0 timeout 400 python tools/bench_configs.py --skip c4 --c5-units 1200 > $O/${TAG}_configs_c3_c5_noprefetch.json 2> $O/${TAG}_configs_nopf.err ;;
    tests_all)
      timeout 700 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -15 $O/${TAG}_pytest_gpu.log ;;
    *) eval "$s" ;;
  esac
  echo "[$s] rc=$? $(( $(date +%s) - t0 )) s"
done
