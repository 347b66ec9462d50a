#!/bin/bash
# round-2 GPU session: full GPU test-suite, prefill bench, predictor sweep, default bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
stage() { echo "=== $1 ($(date +%T))" | tee -a $O/session2.log; }
for s in ${STAGES:-tests prefill bench}; do
case $s in
tests) stage "pytest -m gpu"
  timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  tail -15 $O/pytest_gpu.log;;
prefill) stage "prefill bench"
  timeout 600 python tools/prefill_bench.py --json $O/prefill_bench.json > $O/prefill_bench.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  cat $O/prefill_bench.log | tail -20;;
bench) stage "bench.py default"
  timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" | tee -a $O/session2.log
  head -c 6000 $O/bench_default.json; tail -5 $O/bench_default.err;;
benchref) stage "bench.py --impl reference"
  timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" | tee -a $O/session2.log
  head -c 1500 $O/bench_reference.json;;
trace7b) for arm in b200; do stage "7B trace $arm"
  timeout 900 python tools/engine_trace.py --arm $arm --model llama-2-7b --num-prompts 10000 --stop-after 1000 --json $O/e_7b_$arm.json > $O/e_7b_$arm.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  grep TRACE $O/e_7b_$arm.log | tail -1 | head -c 1500; done;;
predbench) stage "predictor bench"
  for v in "" "--no-pdl" "--gemm 1" "--gemm 1 --no-pdl"; do
  echo "--- predictor_bench $v" >> $O/predictor_bench.log
  timeout 600 python tools/predictor_bench.py $v >> $O/predictor_bench.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  done
  grep -v Warn $O/predictor_bench.log | tail -80;;
predtests) stage "predictor tests"
  timeout 900 python -m pytest tests/test_gpu_predictor.py -x -q -m gpu -p no:cacheprovider > $O/pytest_pred.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  tail -5 $O/pytest_pred.log;;
profile) for arm in b200 ref_cuda; do stage "profile $arm"
  timeout 900 python tools/engine_trace.py --arm $arm --model llama-2-7b --num-prompts 10000 --stop-after 1000 --max-steps 260 --profile 200:40:$O/engine_profile_$arm.txt --json $O/e_prof_$arm.json > $O/e_prof_$arm.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  head -50 $O/engine_profile_$arm.txt; done;;
attntests) stage "attention tests"
  timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_hotpath.py tests/test_gpu_properties.py -x -q -m gpu -p no:cacheprovider > $O/pytest_attn.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  tail -8 $O/pytest_attn.log;;
attntune) stage "attention split tuning"
  timeout 900 python tools/attn_tune.py --grid split --cases ${ATTN_CASES:-70b-tp8-bf16:256:trace,7b-tp8-bf16:256:trace,8b-tp8-bf16:256:trace,8b-gqa-bf16:64:trace,7b-tp2-bf16:256:trace,7b-mha-bf16:256:trace} --json $O/attn_tune_split.json > $O/attn_tune_split.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  cat $O/attn_tune_split.log;;
attntrace) stage "attention timeline (trace build)"
  VLTR_LIB=$PWD/vllm_ltr_b200/libvllm_ltr_b200_trace.so timeout 600 python tools/attn_tune.py --trace --cases ${ATTN_CASES:-70b-tp8-bf16:256:trace,7b-tp8-bf16:256:trace} > $O/attn_trace.log 2>&1; echo "rc=$?" | tee -a $O/session2.log
  VLTR_PDL=0 VLTR_LIB=$PWD/vllm_ltr_b200/libvllm_ltr_b200_trace.so timeout 600 python tools/attn_tune.py --trace --cases 70b-tp8-bf16:256:trace >> $O/attn_trace.log 2>&1
  cat $O/attn_trace.log;;
ipc) stage "ipc probe"; timeout 120 python tools/ipc_probe.py > $O/ipc_probe.log 2>&1; cat $O/ipc_probe.log | tail -5;;
esac
done
stage done
