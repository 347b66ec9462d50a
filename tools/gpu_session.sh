#!/bin/bash
# One gpurun session: smoke, GPU tests, CUDA golden vectors, bench, sweep, ncu.  Everything lands
# in gpurun_out/.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_session.sh [steps...]'
set -u
OUT=gpurun_out
mkdir -p $OUT
STEPS=${*:-smoke tests golden bench sweep launches ncu}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for s in $STEPS; do
  echo "=== $s $(date +%T)"
  case $s in
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log ;;
    tests)   timeout 1500 python -m pytest tests -q -m gpu --maxfail=60 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log ;;
    golden)  timeout 600 python tests/golden/gen_golden_cuda.py $OUT/ref_cuda.npz > $OUT/golden.log 2>&1; echo "golden rc=$?"; tail -2 $OUT/golden.log ;;
    bench)   timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "benchref rc=$?"; cat $OUT/bench_ref.json ;;
    sweep)   timeout 1200 python tools/op_sweep.py --json $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log ;;
    sweepq)  timeout 600 python tools/op_sweep.py --quick --ops attn --json $OUT/sweepq.json > $OUT/sweepq.log 2>&1; echo "sweepq rc=$?"; cat $OUT/sweepq.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline > $OUT/launches.log 2>&1; echo "launches rc=$?"; tail -2 $OUT/launches.log ;;
    ncusmall) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reshape_and_cache|rms_vec|act_vec|rope_neox|reduce_kernel" -s 40 -c 10 -f -o $OUT/small_prof python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline > $OUT/ncusmall.log 2>&1; echo "ncusmall rc=$?"; tail -2 $OUT/ncusmall.log ;;
    ncu)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 8 -c 2 -f -o $OUT/attn_prof python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log ;;
  esac
done
echo "=== done $(date +%T)"
