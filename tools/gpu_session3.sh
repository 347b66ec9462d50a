#!/bin/bash
# round-2 (second half) GPU session.  STAGES selects what runs; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
stage() { echo "=== $1 ($(date +%T))" | tee -a $O/session3.log; }
export_rep() {
  local rep=$O/$1.ncu-rep
  [ -f $rep ] || return 0
  ncu -i $rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  local sz=$(stat -c %s $rep)
  if [ $sz -gt 14000000 ]; then rm -f $rep; fi
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
for s in ${STAGES:-quick attn cache}; do
case $s in
quick) stage "targeted tests"
  timeout 900 python -m pytest tests/test_gpu_cache.py tests/test_gpu_logits.py tests/test_gpu_golden.py -x -q -m gpu -p no:cacheprovider > $O/pytest_quick.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -6 $O/pytest_quick.log
  timeout 900 python -m pytest tests/test_gpu_predictor.py -x -q -m gpu -k dense_layer -p no:cacheprovider > $O/pytest_dense.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -4 $O/pytest_dense.log;;
attntests) stage "attention tests"
  timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_hotpath.py tests/test_gpu_properties.py -x -q -m gpu -p no:cacheprovider > $O/pytest_attn.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -8 $O/pytest_attn.log;;
attn) stage "attention timing"
  timeout 600 python tools/attn_tune.py --ref --cases ${ATTN_CASES:-8b-gqa-fp8:256:trace,8b-gqa-fp8:256:2048,8b-gqa-bf16:256:trace,7b-mha-bf16:256:trace,70b-tp8-bf16:256:trace} --json $O/attn_tune.json > $O/attn_tune.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  cat $O/attn_tune.log;;
mb4) stage "e5m2 attention: 128-register build (4 CTAs/SM) vs default"
  MB4=$PWD/vllm_ltr_b200/libvllm_ltr_b200_mb4.so
  C=8b-gqa-fp8:256:trace,8b-gqa-fp8:256:2048,8b-gqa-fp8:64:trace
  echo "--- default build" > $O/attn_mb4.log
  timeout 300 python tools/attn_tune.py --grid fp8 --cases $C >> $O/attn_mb4.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  echo "--- VLTR_ATTN_FP8_MINB=4 build" >> $O/attn_mb4.log
  VLTR_LIB=$MB4 timeout 300 python tools/attn_tune.py --grid fp8 --cases $C >> $O/attn_mb4.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  cat $O/attn_mb4.log
  VLTR_LIB=$MB4 timeout 900 python -m pytest tests/test_gpu_attention.py -x -q -m gpu -k "fp8 or e5m2" -p no:cacheprovider > $O/pytest_attn_mb4.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -4 $O/pytest_attn_mb4.log;;
timeline) stage "attention per-warp timeline (trace build) + split grid"
  C=${ATTN_CASES:-8b-gqa-fp8:256:trace,8b-gqa-bf16:256:trace}
  VLTR_LIB=$PWD/vllm_ltr_b200/libvllm_ltr_b200_trace.so timeout 600 python tools/attn_tune.py --trace --cases $C > $O/attn_timeline.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  [ -n "${ATTN_GRID:-}" ] && timeout 600 python tools/attn_tune.py --grid $ATTN_GRID --cases $C >> $O/attn_timeline.log 2>&1
  cat $O/attn_timeline.log;;
pair) stage "two KV blocks per iteration: default build (e5m2 only) vs PAIR=2 build (2-byte caches too)"
  P2=$PWD/vllm_ltr_b200/libvllm_ltr_b200_p2.so
  C=8b-gqa-fp8:256:trace,8b-gqa-fp8:256:2048,8b-gqa-bf16:256:trace,70b-tp8-bf16:256:trace,8b-gqa-fp8:64:trace,8b-gqa-bf16:64:trace
  echo "--- default build (VLTR_ATTN_PAIR=1)" > $O/attn_pair.log
  timeout 300 python tools/attn_tune.py --cases $C >> $O/attn_pair.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  echo "--- VLTR_ATTN_PAIR=2 build" >> $O/attn_pair.log
  VLTR_LIB=$P2 timeout 300 python tools/attn_tune.py --grid ctas --cases $C >> $O/attn_pair.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  cat $O/attn_pair.log
  timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_hotpath.py tests/test_gpu_properties.py -x -q -m gpu -p no:cacheprovider > $O/pytest_attn.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -4 $O/pytest_attn.log
  VLTR_LIB=$P2 timeout 900 python -m pytest tests/test_gpu_attention.py -x -q -m gpu -p no:cacheprovider > $O/pytest_attn_p2.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -4 $O/pytest_attn_p2.log;;
ropetests) stage "fused rope + cache tests, hot-path tests"
  timeout 900 python -m pytest tests/test_gpu_rope_cache.py tests/test_gpu_hotpath.py tests/test_gpu_cache.py -x -q -m gpu -p no:cacheprovider > $O/pytest_rope.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -6 $O/pytest_rope.log;;
chain) stage "chain bench"
  timeout 600 python tools/chain_bench.py --model llama-2-7b > $O/chain_bench.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  timeout 600 python tools/chain_bench.py --model llama-3-8b >> $O/chain_bench.log 2>&1
  timeout 600 python tools/chain_bench.py --model llama-3-8b-fp8kv >> $O/chain_bench.log 2>&1
  cat $O/chain_bench.log;;
cache) stage "reshape_and_cache bench"
  timeout 600 python tools/cache_bench.py --json $O/cache_bench.json > $O/cache_bench.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  cat $O/cache_bench.log;;
ncufinal) stage "ncu --set full: e5m2 and bf16 GQA attention, fused rope+cache"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 4 -c 1 -f -o $O/attn_fp8_final python tools/attn_tune.py --cases 8b-gqa-fp8:256:trace > $O/ncu_attn_fp8.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  export_rep attn_fp8_final
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 4 -c 1 -f -o $O/attn_gqa_final python tools/attn_tune.py --cases 8b-gqa-bf16:256:trace > $O/ncu_attn_gqa.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  export_rep attn_gqa_final
  timeout 600 ncu --set full --clock-control none -k regex:rope_and_cache -s 8 -c 2 -f -o $O/rope_cache_prof python tools/chain_bench.py --model llama-2-7b --n 8 > $O/ncu_rope_cache.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  export_rep rope_cache_prof;;
ncu) stage "ncu: e5m2 attention, reshape_and_cache"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attention_kernel -s 4 -c 1 -f -o $O/attn_fp8_prof python tools/attn_tune.py --cases 8b-gqa-fp8:256:trace > $O/ncu_attn_fp8.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  export_rep attn_fp8_prof
  timeout 600 ncu --set full --clock-control none -k regex:reshape_and_cache -s 40 -c 3 -f -o $O/cache_prof python tools/cache_bench.py --n 8 > $O/ncu_cache.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  export_rep cache_prof;;
tests) stage "pytest -m gpu (full)"
  timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -15 $O/pytest_gpu.log;;
smoke) stage "smoke"
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" | tee -a $O/session3.log; tail -3 $O/smoke.log;;
bench) stage "bench.py default"
  timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" | tee -a $O/session3.log
  head -c 7000 $O/bench_default.json; tail -5 $O/bench_default.err;;
benchref) stage "bench.py --impl reference"
  timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" | tee -a $O/session3.log
  head -c 1500 $O/bench_reference.json;;
launches) stage "ncu launch list of bench.py"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"paged_attention|merge_partials|reshape_and_cache|rms_|rope_|act_|rank_|gemm_tn|layernorm_kernel|embed_kernel|score_kernel|attention_kernel" -c 4000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --launch eager --trace off --no-extras --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "rc=$?" | tee -a $O/session3.log;;
fixed) stage "attention fixed cost (one tile per item)"
  timeout 300 python tools/attn_tune.py --cases 8b-gqa-bf16:256:16,7b-mha-bf16:256:16,8b-gqa-fp8:256:16,70b-tp8-bf16:256:16,8b-gqa-bf16:256:128 > $O/attn_fixed.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  cat $O/attn_fixed.log;;
gemmtests) stage "dense layer / logits / predictor tests"
  timeout 900 python -m pytest tests/test_gpu_predictor.py tests/test_gpu_logits.py -x -q -m gpu -p no:cacheprovider > $O/pytest_gemm.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -4 $O/pytest_gemm.log
  timeout 300 python - > $O/lm_head_time.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from vllm_ltr_b200.logits import lm_head_logits
dev = "cuda:0"
for rows, vocab, hidden, dt in ((256, 32000, 4096, torch.float16), (256, 128256, 4096, torch.bfloat16), (64, 32000, 4096, torch.float16), (256, 4000, 8192, torch.bfloat16)):
    h = torch.randn(rows, hidden, device=dev).to(dt)
    ws = [(torch.randn(vocab, hidden, device=dev) * 0.02).to(dt) for _ in range(3)]   # rotate: > L2
    def run(fn, n=30):
        for w in ws: fn(h, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(h, ws[i % 3])
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    ours = run(lambda a, w: lm_head_logits(a, w))
    cub = run(lambda a, w: torch.matmul(a, w.t()))
    gb = vocab * hidden * 2 / 1e3
    print(f"lm_head rows={rows} vocab={vocab} hidden={hidden} {dt}: ours {ours:7.1f} us ({gb/ours:6.0f} GB/s of weights)  cuBLAS {cub:7.1f} us")
PY
  cat $O/lm_head_time.log;;
enginetest) stage "engine test"
  timeout 1800 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -p no:cacheprovider > $O/pytest_engine.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  tail -8 $O/pytest_engine.log;;
profile) stage "engine profile b200"
  timeout 900 python tools/engine_trace.py --arm b200 --model llama-2-7b --num-prompts 10000 --stop-after 1000 --max-steps 260 --profile 200:40:$O/engine_profile_b200.txt --json $O/e_prof_b200.json > $O/e_prof_b200.log 2>&1; echo "rc=$?" | tee -a $O/session3.log
  head -40 $O/engine_profile_b200.txt;;
esac
done
stage done
