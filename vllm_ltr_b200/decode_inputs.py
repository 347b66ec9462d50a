"""Host-side step preparation for decode batches (first piece).

``prepare_decode`` is a drop-in for ``ModelRunner._prepare_decode``
(vllm/worker/model_runner.py:424-548): same arguments, same ``PrepareDecodeMetadata`` (token /
position / slot lists, ``context_lens`` and padded ``block_tables`` tensors inside the attention
metadata), element for element.  The reference pads every sequence's block table as a Python
list (``make_tensor_with_pad``, vllm/utils.py) and lets ``torch.tensor`` walk the nested lists -
4-5 ms at 256 running sequences of the trace, as long as the whole GPU step of Llama-3-8B.  Here
the ragged tables are converted once (one ``np.fromiter`` over all block ids) and scattered into
the padded array with a mask; the per-sequence Python loop only reads what it must from the
``SequenceData`` objects.

``prepare_prompt`` is the same for the prefill chunks of a step (``ModelRunner._prepare_prompt``,
model_runner.py:224-422) in the configuration the LTR engine runs - chunked prefill, no LoRA, no
multi-modal input, no prefix caching, no sliding window; anything else is handed to the
reference method.  The reference walks every chunk token in Python to compute its slot and
creates six small device tensors one blocking copy at a time plus two cumsum launches; here
positions and slots of a chunk are two numpy expressions, the prefix sums are computed on the
host, and all int32 metadata of the step crosses to the device in ONE copy (block tables,
context lengths and both start-offset vectors are views of it).

``install(model_runner)`` re-binds both on a reference ``ModelRunner``.
"""
import types
from itertools import chain
from typing import List

import numpy as np
import torch


def _ref():
    import vllm.worker.model_runner as M
    return M


def _padded_tables(block_tables: List[List[int]], out: np.ndarray) -> None:
    """out[i, :len(block_tables[i])] = block_tables[i] for every row, in one conversion."""
    n = len(block_tables)
    lens = np.fromiter(map(len, block_tables), dtype=np.int64, count=n)
    total = int(lens.sum())
    if total == 0:
        return
    flat = np.fromiter(chain.from_iterable(block_tables), dtype=np.int32, count=total)
    mask = np.arange(out.shape[1])[None, :] < lens[:, None]
    out[:n][mask] = flat          # row-major: the ragged order


def prepare_decode(self, seq_group_metadata_list):
    M = _ref()
    input_tokens: List[int] = []
    input_positions: List[int] = []
    slot_mapping: List[int] = []
    context_lens: List[int] = []
    block_tables: List[List[int]] = []
    lora_index_mapping: List[int] = []
    lora_prompt_mapping: List[int] = []
    lora_requests = set()

    if len(seq_group_metadata_list) == 0:
        return M.PrepareDecodeMetadata.empty()

    bs, sw = self.block_size, self.sliding_window
    sw_blocks = None if sw is None else sw // bs
    for md in seq_group_metadata_list:
        assert not md.is_prompt
        assert md.token_chunk_size == 1
        lora_id = md.lora_int_id
        if lora_id > 0:
            lora_requests.add(md.lora_request)
        tables = md.block_tables
        for seq_id, seq_data in md.seq_data.items():
            input_tokens.append(seq_data.get_last_token_id())
            seq_len = seq_data.get_len()
            position = seq_len - 1
            input_positions.append(position)
            context_lens.append(seq_len if sw is None else min(seq_len, sw))
            table = tables[seq_id]
            slot_mapping.append(table[position // bs] * bs + position % bs)
            lora_index_mapping.append(lora_id)
            lora_prompt_mapping.append(lora_id)
            block_tables.append(table if sw_blocks is None else table[-sw_blocks:])

    batch_size = len(input_tokens)
    max_context_len = max(context_lens)
    use_captured_graph = (not self.model_config.enforce_eager
                          and batch_size <= M._BATCH_SIZES_TO_CAPTURE[-1]
                          and max_context_len <= self.max_context_len_to_capture)
    if use_captured_graph:
        graph_batch_size = M._get_graph_batch_size(batch_size)
        assert graph_batch_size >= batch_size
        pad = graph_batch_size - batch_size
        input_tokens.extend([0] * pad)
        input_positions.extend([0] * pad)
        slot_mapping.extend([M._PAD_SLOT_ID] * pad)
        context_lens.extend([1] * pad)
        block_tables.extend([] for _ in range(pad))
        lora_index_mapping.extend([0] * pad)
        batch_size = graph_batch_size

    context_lens_tensor = torch.tensor(context_lens, dtype=torch.int, device=self.device)

    if use_captured_graph:
        # the persistent [max batch, max blocks] array: rows are overwritten up to their length
        # only, exactly like the reference's row-by-row assignment (:511-514)
        input_block_tables = self.graph_block_tables[:batch_size]
        _padded_tables(block_tables, input_block_tables)
        block_tables_t = torch.tensor(input_block_tables, device=self.device)
    else:
        width = max(len(t) for t in block_tables)
        padded = np.zeros((batch_size, width), dtype=np.int32)
        _padded_tables(block_tables, padded)
        block_tables_t = torch.from_numpy(padded).to(self.device)

    attn_metadata = self.attn_backend.make_metadata(
        is_prompt=False, prompt_lens=None, prompt_lens_tensor=None, max_subquery_len=None,
        max_context_len=max_context_len, max_prompt_len=None, subquery_start_loc=None,
        seq_start_loc=None, context_lens=context_lens_tensor, block_tables=block_tables_t,
        use_cuda_graph=use_captured_graph)
    return M.PrepareDecodeMetadata(
        input_tokens=input_tokens, input_positions=input_positions, attn_metadata=attn_metadata,
        lora_index_mapping=lora_index_mapping, lora_prompt_mapping=lora_prompt_mapping,
        lora_requests=lora_requests, slot_mapping=slot_mapping)


def prepare_prompt(self, seq_group_metadata_list):
    M = _ref()
    mds = seq_group_metadata_list
    if len(mds) == 0:
        return M.PreparePromptMetadata.empty()
    chunked = self.scheduler_config is not None and self.scheduler_config.chunked_prefill_enabled
    plain = chunked and self.sliding_window is None and all(
        md.lora_int_id == 0 and not md.multi_modal_data and not md.computed_block_nums
        and md.block_tables is not None and len(md.seq_data) == 1 for md in mds)
    if not plain:        # profiling run, LoRA, images, prefix caching, sliding window: reference
        return self._vltr_ref_prepare_prompt(mds)
    bs = self.block_size
    n = len(mds)
    tokens: List[int] = []
    pos_parts, slot_parts, tables = [], [], []
    ctx = np.empty(n, dtype=np.int64)          # tokens already computed (the prefix)
    end = np.empty(n, dtype=np.int64)          # prefix + this chunk
    for i, md in enumerate(mds):
        assert md.is_prompt
        (seq_id, data), = md.seq_data.items()
        c = data.get_num_computed_tokens()
        e = min(data.get_len(), c + md.token_chunk_size)
        ctx[i], end[i] = c, e
        tokens.extend(data.get_token_ids()[c:e])
        table = md.block_tables[seq_id]
        tables.append(table)
        p = np.arange(c, e, dtype=np.int64)
        pos_parts.append(p)
        # slot of chunk token at position p: its block's id * block size + offset in the block
        slot_parts.append(np.asarray(table, dtype=np.int64)[p // bs] * bs + p % bs)
    sub = end - ctx
    assert int(sub.max()) > 0
    width = max(len(t) for t in tables)
    # one int32 staging array: [block tables | context lens | subquery starts | sequence starts]
    nbt = n * width
    stage = np.zeros(nbt + n + 2 * (n + 1), dtype=np.int32)
    if width:
        _padded_tables(tables, stage[:nbt].reshape(n, width))
    stage[nbt:nbt + n] = ctx
    np.cumsum(sub, out=stage[nbt + n + 1:nbt + 2 * n + 1])
    np.cumsum(end, out=stage[nbt + 2 * n + 2:])
    dev = torch.from_numpy(stage).to(self.device)
    prompt_lens = end.tolist()
    attn_metadata = self.attn_backend.make_metadata(
        is_prompt=True, prompt_lens=prompt_lens,
        prompt_lens_tensor=torch.from_numpy(end).to(self.device), max_subquery_len=int(sub.max()),
        max_context_len=None, max_prompt_len=int(end.max()),
        subquery_start_loc=dev[nbt + n:nbt + 2 * n + 1], seq_start_loc=dev[nbt + 2 * n + 1:],
        context_lens=dev[nbt:nbt + n], block_tables=dev[:nbt].view(n, width),
        use_cuda_graph=False)
    ntok = len(tokens)
    return M.PreparePromptMetadata(
        input_tokens=tokens, input_positions=np.concatenate(pos_parts).tolist(),
        attn_metadata=attn_metadata, prompt_lens=prompt_lens, subquery_lens=sub.tolist(),
        lora_index_mapping=[0] * ntok,
        lora_prompt_mapping=[0] * sum(int(sub[i]) if md.sampling_params.prompt_logprobs else 1
                                      for i, md in enumerate(mds)),
        lora_requests=set(),
        multi_modal_input=None, slot_mapping=np.concatenate(slot_parts).tolist())


def install(model_runner):
    """Bind the fast ``_prepare_decode`` / ``_prepare_prompt`` on a reference ``ModelRunner``."""
    model_runner._prepare_decode = types.MethodType(prepare_decode, model_runner)
    ref_prompt = getattr(model_runner, "_prepare_prompt", None)
    if ref_prompt is not None:
        if not hasattr(model_runner, "_vltr_ref_prepare_prompt"):
            model_runner._vltr_ref_prepare_prompt = ref_prompt
        model_runner._prepare_prompt = types.MethodType(prepare_prompt, model_runner)
    return model_runner
