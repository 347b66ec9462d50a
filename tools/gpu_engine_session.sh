#!/bin/bash
# GPU session: the reference engine with both arms (tools/engine_trace.py). Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
stage() { echo "=== $1 ($(date +%T))" | tee -a $O/engine_session.log; }
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $O/gpu.txt 2>&1

if [ "${STAGE1:-1}" = 1 ]; then
stage "tiny ref_cuda"
timeout 600 python tools/engine_trace.py --arm ref_cuda --model llama-tiny --num-prompts 64 --stop-after 0 \
  --max-output 32 --record --logprobs 5 --gpu-memory-utilization 0.3 --swap-space 1 \
  --json $O/e_tiny_ref.json > $O/e_tiny_ref.log 2>&1; echo "rc=$?" | tee -a $O/engine_session.log
tail -5 $O/e_tiny_ref.log
stage "tiny b200"
timeout 600 python tools/engine_trace.py --arm b200 --model llama-tiny --num-prompts 64 --stop-after 0 \
  --max-output 32 --record --gpu-memory-utilization 0.3 --swap-space 1 \
  --json $O/e_tiny_b200.json > $O/e_tiny_b200.log 2>&1; echo "rc=$?" | tee -a $O/engine_session.log
tail -5 $O/e_tiny_b200.log
python tools/engine_compare.py $O/e_tiny_ref.json $O/e_tiny_b200.json > $O/e_tiny_compare.json 2>&1
echo "compare rc=$?" | tee -a $O/engine_session.log
head -c 1500 $O/e_tiny_compare.json
fi

if [ "${STAGE2:-1}" = 1 ]; then
stage "pytest test_gpu_engine"
timeout 1800 python -m pytest tests/test_gpu_engine.py -x -q -m gpu > $O/pytest_engine.log 2>&1
echo "rc=$?" | tee -a $O/engine_session.log
tail -30 $O/pytest_engine.log
fi

if [ "${STAGE3:-1}" = 1 ]; then
NP=${NP:-10000}; STOP=${STOP:-1000}
for arm in b200 ref_cuda; do
stage "7B trace $arm NP=$NP STOP=$STOP"
timeout ${ARM_TIMEOUT:-900} python tools/engine_trace.py --arm $arm --model llama-2-7b --num-prompts $NP --stop-after $STOP \
  --json $O/e_7b_$arm.json > $O/e_7b_$arm.log 2>&1; echo "rc=$?" | tee -a $O/engine_session.log
grep TRACE $O/e_7b_$arm.log | tail -1 | head -c 1800; tail -3 $O/e_7b_$arm.log
done
fi
stage done
