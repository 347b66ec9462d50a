#!/bin/bash
# One gpurun session: smoke, GPU tests, CUDA golden vectors, bench, sweep, ncu.  Everything lands
# in gpurun_out/.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_session.sh [steps...]'
set -u
OUT=gpurun_out
mkdir -p $OUT
STEPS=${*:-smoke tests golden bench sweep launches ncu}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
# ncu reports with imported source run to tens of MB and gpurun_out/ is capped at 64 MiB: keep the
# CSV pages (what profiles/ is written from) and drop reports above 14 MB
export_rep() {
  local rep=$OUT/$1.ncu-rep
  [ -f $rep ] || return 0
  ncu -i $rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  ncu -i $rep --page details --csv > $OUT/$1.details.csv 2>/dev/null
  ncu -i $rep --page source --csv > $OUT/$1.source.csv 2>/dev/null
  local sz=$(stat -c %s $rep)
  if [ $sz -gt 14000000 ]; then rm -f $rep; echo "$1.ncu-rep: $sz bytes, dropped (CSV pages kept)"; fi
  if [ -f $OUT/$1.source.csv ] && [ $(stat -c %s $OUT/$1.source.csv) -gt 8000000 ]; then gzip -f $OUT/$1.source.csv; fi
}
for s in $STEPS; do
  echo "=== $s $(date +%T)"
  case $s in
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log ;;
    tests)   timeout 1500 python -m pytest tests -q -m gpu --maxfail=60 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log ;;
    predtest) timeout 300 python -m pytest tests/test_gpu_predictor.py -x -q -k "dense_layer" -p no:cacheprovider > $OUT/pred_gemm.log 2>&1; echo "pred gemm rc=$?"; tail -12 $OUT/pred_gemm.log
              timeout 600 python -m pytest tests/test_gpu_predictor.py -x -q -k "not dense_layer" -p no:cacheprovider > $OUT/pred_fwd.log 2>&1; echo "pred fwd rc=$?"; tail -12 $OUT/pred_fwd.log ;;
    newtests) timeout 900 python -m pytest tests/test_gpu_ltr.py tests/test_gpu_golden.py tests/test_gpu_cache.py -q -p no:cacheprovider > $OUT/newtests.log 2>&1; echo "newtests rc=$?"; tail -8 $OUT/newtests.log ;;
    sweepcache) timeout 600 python tools/op_sweep.py --ops cache --json $OUT/sweep_cache.json > $OUT/sweep_cache.log 2>&1; echo "sweepcache rc=$?"; cat $OUT/sweep_cache.log ;;
    attntests) timeout 1200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_golden.py -q --maxfail=30 -p no:cacheprovider > $OUT/attntests.log 2>&1; echo "attntests rc=$?"; tail -40 $OUT/attntests.log ;;
    sweepattn) timeout 900 python tools/op_sweep.py --ops attn --json $OUT/sweep_attn.json > $OUT/sweep_attn.log 2>&1; echo "sweepattn rc=$?"; cat $OUT/sweep_attn.log ;;
    ncumma)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 3 -c 1 -f -o $OUT/attn_mma_prof python tools/op_sweep.py --ops attn --only 8b-gqa-bf16:256:trace:v1 > $OUT/ncumma.log 2>&1; echo "ncumma rc=$?"; tail -3 $OUT/ncumma.log
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 3 -c 1 -f -o $OUT/attn_tp8_prof python tools/op_sweep.py --ops attn --only 70b-tp8-bf16:256:512:v1 > $OUT/ncutp8.log 2>&1; echo "ncutp8 rc=$?"; tail -3 $OUT/ncutp8.log ;;
    sweepsampler) timeout 600 python tools/op_sweep.py --ops sampler --json $OUT/sweep_sampler.json > $OUT/sweep_sampler.log 2>&1; echo "sweepsampler rc=$?"; cat $OUT/sweep_sampler.log ;;
    swaptest) timeout 300 python -m pytest tests/test_gpu_cache.py -q -k "swap" -p no:cacheprovider > $OUT/swaptest.log 2>&1; echo "swaptest rc=$?"; tail -6 $OUT/swaptest.log
              timeout 300 python tools/op_sweep.py --ops swap > $OUT/sweep_swap.log 2>&1; echo "sweepswap rc=$?"; cat $OUT/sweep_swap.log ;;
    sweepsmall) timeout 600 python tools/op_sweep.py --ops rms,silu,rope,cache --json $OUT/sweep_small.json > $OUT/sweep_small.log 2>&1; echo "sweepsmall rc=$?"; cat $OUT/sweep_small.log ;;
    sweeppred) timeout 600 python tools/op_sweep.py --ops gemm,predictor --json $OUT/sweep_pred.json > $OUT/sweep_pred.log 2>&1; echo "sweeppred rc=$?"; cat $OUT/sweep_pred.log ;;
    ncugemm) timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 20 -c 3 -f -o $OUT/gemm_prof python tools/op_sweep.py --ops gemm > $OUT/ncugemm.log 2>&1; echo "ncugemm rc=$?"; tail -3 $OUT/ncugemm.log ;;
    prefilltests) timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_properties.py -q --maxfail=20 -p no:cacheprovider > $OUT/prefilltests.log 2>&1; echo "prefilltests rc=$?"; tail -25 $OUT/prefilltests.log ;;
    hotpathtests) timeout 600 python -m pytest tests/test_gpu_hotpath.py -q -p no:cacheprovider > $OUT/hotpathtests.log 2>&1; echo "hotpathtests rc=$?"; tail -15 $OUT/hotpathtests.log ;;
    tune)    timeout 900 python tools/attn_tune.py --grid wide --trace --ref --json $OUT/attn_tune.json > $OUT/attn_tune.log 2>&1; echo "tune rc=$?"; cat $OUT/attn_tune.log ;;
    chain)   timeout 600 python tools/chain_bench.py --json $OUT/chain.json > $OUT/chain.log 2>&1; echo "chain rc=$?"; cat $OUT/chain.log ;;
    tune3)   timeout 900 python tools/attn_tune.py --grid r3 --trace --cases 70b-tp8-bf16:256:trace,8b-gqa-bf16:64:trace,8b-gqa-bf16:256:trace,7b-mha-bf16:64:trace --json $OUT/attn_tune3.json > $OUT/attn_tune3.log 2>&1; echo "tune3 rc=$?"; cat $OUT/attn_tune3.log ;;
    samplertests) timeout 600 python -m pytest tests/test_gpu_sampler.py -q -p no:cacheprovider > $OUT/samplertests.log 2>&1; echo "samplertests rc=$?"; tail -25 $OUT/samplertests.log ;;
    tp2)     timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/tp_check.py > $OUT/tp_check_n2.log 2>&1; echo "tp_check rc=$?"; tail -22 $OUT/tp_check_n2.log
             timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; cat $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
             timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --no-cpu-baseline --launch eager > $OUT/bench_n2_eager.json 2> $OUT/bench_n2_eager.err; echo "bench n2 eager rc=$?"; cat $OUT/bench_n2_eager.json
             timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --model llama-2-70b --layers 20 --no-cpu-baseline > $OUT/bench_n2_70b20.json 2> $OUT/bench_n2_70b20.err; echo "bench n2 70b rc=$?"; cat $OUT/bench_n2_70b20.json ;;
    tune2)   timeout 900 python tools/attn_tune.py --grid r3 --trace --json $OUT/attn_tune2.json > $OUT/attn_tune2.log 2>&1; echo "tune2 rc=$?"; cat $OUT/attn_tune2.log ;;
    tuneq)   timeout 600 python tools/attn_tune.py --trace --json $OUT/attn_tuneq.json > $OUT/attn_tuneq.log 2>&1; echo "tuneq rc=$?"; cat $OUT/attn_tuneq.log ;;
    benchab) timeout 600 python bench.py --no-pdl --launch eager --no-cpu-baseline > $OUT/bench_nopdl_eager.json 2> $OUT/bench_ab.err; echo "benchab rc=$?"; cat $OUT/bench_nopdl_eager.json ;;
    benchmodels) for m in llama-3-8b llama-3-8b-fp8kv; do timeout 600 python bench.py --model $m --no-cpu-baseline > $OUT/bench_$m.json 2> $OUT/bench_$m.err; echo "bench $m rc=$?"; cat $OUT/bench_$m.json; done ;;
    ncusampler) timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 3 -c 2 -f -o $OUT/sampler_prof python tools/op_sweep.py --ops sampler > $OUT/ncusampler.log 2>&1; echo "ncusampler rc=$?"; tail -3 $OUT/ncusampler.log; export_rep sampler_prof ;;
    ncugqa)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 8 -c 2 -f -o $OUT/attn_gqa_prof python bench.py --model llama-3-8b --steps 1 --warmup 3 --layers 4 --no-cpu-baseline --launch eager > $OUT/ncugqa.log 2>&1; echo "ncugqa rc=$?"; tail -2 $OUT/ncugqa.log; export_rep attn_gqa_prof
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 8 -c 2 -f -o $OUT/attn_fp8_prof python bench.py --model llama-3-8b-fp8kv --steps 1 --warmup 3 --layers 4 --no-cpu-baseline --launch eager > $OUT/ncufp8.log 2>&1; echo "ncufp8 rc=$?"; tail -2 $OUT/ncufp8.log; export_rep attn_fp8_prof ;;
    golden)  timeout 600 python tests/golden/gen_golden_cuda.py $OUT/ref_cuda.npz > $OUT/golden.log 2>&1; echo "golden rc=$?"; tail -2 $OUT/golden.log ;;
    bench)   timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "benchref rc=$?"; cat $OUT/bench_ref.json ;;
    sweep)   timeout 1200 python tools/op_sweep.py --json $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/sweep.log ;;
    sweepq)  timeout 600 python tools/op_sweep.py --quick --ops attn --json $OUT/sweepq.json > $OUT/sweepq.log 2>&1; echo "sweepq rc=$?"; cat $OUT/sweepq.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline --launch eager > $OUT/launches.log 2>&1; echo "launches rc=$?"; tail -2 $OUT/launches.log ;;
    ncusmall) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reshape_and_cache|rms_vec|act_vec|rope_neox|reduce_kernel" -s 40 -c 10 -f -o $OUT/small_prof python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline --launch eager > $OUT/ncusmall.log 2>&1; echo "ncusmall rc=$?"; tail -2 $OUT/ncusmall.log ;;
    ncu)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 8 -c 2 -f -o $OUT/attn_prof python bench.py --steps 1 --warmup 3 --layers 4 --no-cpu-baseline --launch eager > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log; export_rep attn_prof ;;
  esac
done
echo "=== done $(date +%T)"
