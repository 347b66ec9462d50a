"""CPU: vllm_ltr_b200.decode_inputs.prepare_decode against the reference's own
ModelRunner._prepare_decode (vllm/worker/model_runner.py:424-548), field by field, over ragged
batches, sliding windows and the CUDA-graph padding mode.  Needs /root/reference."""
import random
import types

import numpy as np
import pytest
import torch

from tests import sched_replay as R

pytestmark = pytest.mark.skipif(not R.have_reference(),
                                reason="needs the reference checkout at /root/reference")


def _metas(n, seed, block_size, multi_seq=False):
    from vllm import SamplingParams
    from vllm.sequence import SequenceData, SequenceGroupMetadata
    rnd = random.Random(seed)
    metas, next_block, seq_id = [], 0, 0
    for i in range(n):
        seq_data, tables = {}, {}
        for _ in range(rnd.choice([1, 1, 2, 3]) if multi_seq else 1):
            length = rnd.choice([1, 2, block_size - 1, block_size, block_size + 1, 77, 513, 2047])
            sd = SequenceData([rnd.randrange(1000) for _ in range(min(length, 5))])
            for _ in range(length - min(length, 5)):
                sd.append_token_id(rnd.randrange(1000), 0.0)
            nblk = (length + block_size - 1) // block_size
            blocks = [next_block + rnd.randrange(10_000) for _ in range(nblk)]
            next_block += 10_000
            seq_data[seq_id], tables[seq_id] = sd, blocks
            seq_id += 1
        metas.append(SequenceGroupMetadata(str(i), False, seq_data, SamplingParams(), tables,
                                           token_chunk_size=1))
    return metas


def _runner(block_size, sliding_window, eager, max_capture=8192):
    nb = (max_capture + block_size - 1) // block_size
    return types.SimpleNamespace(
        sliding_window=sliding_window, block_size=block_size, device="cpu",
        model_config=types.SimpleNamespace(enforce_eager=eager),
        max_context_len_to_capture=max_capture,
        graph_block_tables=np.full((256, nb), 7, dtype=np.int32),      # stale content on purpose
        attn_backend=types.SimpleNamespace(make_metadata=lambda **kw: types.SimpleNamespace(**kw)))


@pytest.mark.parametrize("n", [1, 7, 64, 256])
@pytest.mark.parametrize("block_size", [8, 16])
@pytest.mark.parametrize("sliding_window", [None, 256])
@pytest.mark.parametrize("eager", [True, False])
def test_prepare_decode_matches_reference(n, block_size, sliding_window, eager):
    R.import_reference()
    from vllm.worker.model_runner import ModelRunner

    from vllm_ltr_b200 import decode_inputs as ours
    metas = _metas(n, n * 31 + block_size, block_size, multi_seq=(n == 7))
    if sum(len(m.seq_data) for m in metas) > 256:
        pytest.skip("beyond the engine's 256-sequence budget")
    a, b = _runner(block_size, sliding_window, eager), _runner(block_size, sliding_window, eager)
    want = ModelRunner._prepare_decode(a, metas)
    got = ours.prepare_decode(b, metas)
    for f in ("input_tokens", "input_positions", "slot_mapping", "lora_index_mapping",
              "lora_prompt_mapping", "lora_requests"):
        assert getattr(got, f) == getattr(want, f), f
    ga, wa = got.attn_metadata, want.attn_metadata
    assert ga.max_context_len == wa.max_context_len and ga.use_cuda_graph == wa.use_cuda_graph
    assert ga.context_lens.dtype == wa.context_lens.dtype and torch.equal(ga.context_lens, wa.context_lens)
    assert ga.block_tables.dtype == wa.block_tables.dtype and ga.block_tables.shape == wa.block_tables.shape
    assert torch.equal(ga.block_tables, wa.block_tables)
    assert np.array_equal(a.graph_block_tables, b.graph_block_tables)     # same persistent state


def test_empty_batch_and_install():
    R.import_reference()
    from vllm.worker.model_runner import PrepareDecodeMetadata

    from vllm_ltr_b200 import decode_inputs as ours
    r = _runner(16, None, True)
    assert ours.prepare_decode(r, []) == PrepareDecodeMetadata.empty()
    ours.install(r)
    assert r._prepare_decode([]) == PrepareDecodeMetadata.empty()


def _prompt_metas(seed, block_size, first_only=False):
    """Prefill chunks the way the scheduler hands them over under chunked prefill: block tables
    allocated for the whole prompt, some chunks first (0 computed tokens), some continued."""
    from vllm import SamplingParams
    from vllm.sequence import SequenceData, SequenceGroupMetadata
    rnd = random.Random(seed)
    metas, next_block = [], 0
    for i in range(rnd.choice([1, 3, 12])):
        plen = rnd.choice([1, 5, block_size, block_size + 1, 300, 2048])
        sd = SequenceData([rnd.randrange(1000) for _ in range(plen)])
        done = 0 if first_only else rnd.choice([0, 0, min(plen - 1, block_size), plen // 2])
        sd.update_num_computed_tokens(done)
        chunk = rnd.choice([1, 64, 768])
        nblk = (plen + block_size - 1) // block_size
        blocks = [next_block + rnd.randrange(10_000) for _ in range(nblk)]
        next_block += 10_000
        sp = SamplingParams(prompt_logprobs=(2 if i == 1 else None))
        metas.append(SequenceGroupMetadata(str(i), True, {i: sd}, sp, {i: blocks},
                                           token_chunk_size=chunk))
    return metas


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("block_size", [8, 16])
def test_prepare_prompt_matches_reference(seed, block_size):
    R.import_reference()
    from vllm.worker.model_runner import ModelRunner

    from vllm_ltr_b200 import decode_inputs as ours

    def runner():
        r = _runner(block_size, None, True)
        r.scheduler_config = types.SimpleNamespace(chunked_prefill_enabled=True)
        r.vision_language_config = None
        r._prepare_prompt = types.MethodType(ModelRunner._prepare_prompt, r)
        return r
    metas = _prompt_metas(seed, block_size, first_only=(seed % 4 == 0))
    a, b = runner(), ours.install(runner())
    want = ModelRunner._prepare_prompt(a, metas)
    got = b._prepare_prompt(metas)
    for f in ("input_tokens", "input_positions", "slot_mapping", "prompt_lens", "subquery_lens",
              "lora_index_mapping", "lora_prompt_mapping", "lora_requests", "multi_modal_input"):
        assert getattr(got, f) == getattr(want, f), f
    ga, wa = got.attn_metadata, want.attn_metadata
    for f in ("is_prompt", "prompt_lens", "max_subquery_len", "max_context_len", "max_prompt_len",
              "use_cuda_graph"):
        assert getattr(ga, f) == getattr(wa, f), f
    for f in ("prompt_lens_tensor", "subquery_start_loc", "seq_start_loc", "context_lens",
              "block_tables"):
        x, y = getattr(ga, f), getattr(wa, f)
        assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), f


def test_prepare_prompt_delegates_what_it_does_not_cover():
    """No block tables (the memory-profiling run) / chunked prefill off: the reference method."""
    R.import_reference()
    from vllm import SamplingParams
    from vllm.sequence import SequenceData, SequenceGroupMetadata

    from vllm_ltr_b200 import decode_inputs as ours
    r = _runner(16, None, True)
    r.scheduler_config = types.SimpleNamespace(chunked_prefill_enabled=True)
    calls = []
    r._prepare_prompt = lambda mds: calls.append(len(mds)) or "REFERENCE"
    ours.install(r)
    md = SequenceGroupMetadata("0", True, {0: SequenceData([1, 2, 3])}, SamplingParams(), None,
                               token_chunk_size=3)
    assert r._prepare_prompt([md]) == "REFERENCE" and calls == [1]
