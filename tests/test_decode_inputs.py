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
