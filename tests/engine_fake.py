"""CPU stand-in for the model executor of the reference engine: the REAL ``LLMEngine`` (scheduler,
block manager, output processor, detokenizer, stop checker - /root/reference/vllm/engine/
llm_engine.py) runs unmodified, only ``execute_model`` is replaced by a deterministic token
function.  Lets the scheduler plugin be compared with the reference's own scheduler THROUGH THE
ENGINE on a machine without a GPU (tests/test_engine_cpu.py)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")


def engine_trace():
    if TOOLS not in sys.path:
        sys.path.insert(0, TOOLS)
    import engine_trace as E
    return E


def make_fake_executor_class():
    from vllm.executor.executor_base import ExecutorBase
    from vllm.sequence import Logprob, SamplerOutput, SequenceGroupOutput, SequenceOutput

    class FakeExecutor(ExecutorBase):
        NUM_GPU_BLOCKS = 2048
        NUM_CPU_BLOCKS = 1024

        def _init_executor(self):
            self.driver_worker = None
            self.calls = 0

        _init_aux_executor = _init_executor

        def determine_num_available_blocks(self):
            return self.NUM_GPU_BLOCKS, self.NUM_CPU_BLOCKS

        def initialize_cache(self, num_gpu_blocks, num_cpu_blocks):
            pass

        def execute_model(self, seq_group_metadata_list, blocks_to_swap_in, blocks_to_swap_out,
                          blocks_to_copy, num_lookahead_slots=0):
            self.calls += 1
            outs = []
            for md in seq_group_metadata_list:
                samples = []
                for seq_id, data in md.seq_data.items():
                    tok = 1000 + (seq_id * 7919 + data.get_len() * 104729) % 30000
                    samples.append(SequenceOutput(seq_id, tok, {tok: Logprob(0.0)}))
                outs.append(SequenceGroupOutput(samples, None))
            return [SamplerOutput(outputs=outs)]

        def add_lora(self, lora_request):
            return False

        def remove_lora(self, lora_id):
            return False

        def list_loras(self):
            return set()

        def check_health(self):
            return

    return FakeExecutor


def build_fake_llm(assets, schedule_type, gpu_blocks=2048, cpu_blocks=1024):
    """An object with ``.llm_engine`` and ``.request_counter`` (what tools/engine_trace.run_trace
    uses of ``vllm.LLM``) around a real LLMEngine with the fake executor."""
    from vllm.engine.arg_utils import EngineArgs
    from vllm.engine.llm_engine import LLMEngine
    from vllm.utils import Counter
    ex = make_fake_executor_class()
    ex.NUM_GPU_BLOCKS, ex.NUM_CPU_BLOCKS = gpu_blocks, cpu_blocks
    ea = EngineArgs(model=assets["model"], tokenizer=assets["model"], load_format="dummy",
                    enforce_eager=True, enable_chunked_prefill=True, max_num_batched_tokens=768,
                    max_num_seqs=256, block_size=16, schedule_type=schedule_type, swap_space=1,
                    disable_log_stats=True, device="cpu")
    cfg = ea.create_engine_config()
    eng = LLMEngine(**cfg.to_dict(), executor_class=ex, log_stats=False)
    return types.SimpleNamespace(llm_engine=eng, request_counter=Counter())
