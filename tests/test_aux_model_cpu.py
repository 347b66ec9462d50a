"""Host logic of ``B200AuxModel.obtain_aux_scores`` (the scheduler.aux_model contract,
aux_llm_engine.py:332-412) without a GPU: chunked tokenisation on the helper thread must hand the
predictor the same token ids, in the same order, as the one-by-one path, and every group gets its
own score.  The forward itself is replaced by a recording stand-in (the CUDA forward is tested in
tests/test_gpu_predictor.py)."""
import os
import sys
import threading

import pytest
import torch

from vllm_ltr_b200 import predictor as P

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


class Seq:
    def __init__(self, text):
        self.prompt = text


class Group:
    def __init__(self, text):
        self.seqs_dict = {7: Seq(text)}
        self.aux_model_score = None

    def need_aux_model_score(self):
        return self.aux_model_score is None

    def set_aux_model_score(self, s):
        self.aux_model_score = s


class RecordingPredictor:
    """score = f(token ids) so that a prompt routed to the wrong group shows up."""

    def __init__(self):
        self.calls = []
        self.threads = set()

    def score_column(self, prompts):
        self.threads.add(threading.get_ident())
        self.calls.append([list(p) for p in prompts])
        return torch.tensor([float(sum(p) % 997) + 0.5 * len(p) for p in prompts],
                            dtype=torch.float32)


def _tok(text):
    return [2] + [3 + (ord(c) % 200) for c in text]


def _want(texts, L):
    return [float(sum(_tok(t)[:L]) % 997) + 0.5 * len(_tok(t)[:L]) for t in texts]


@pytest.mark.parametrize("n,chunk", [(0, 4), (1, 4), (4, 4), (5, 4), (23, 4), (23, 1), (9, 100)])
@pytest.mark.parametrize("batched", [False, True])
def test_chunked_scores_match_one_by_one(n, chunk, batched):
    texts = ["w" * (1 + (5 * i) % 13) + chr(97 + i % 26) * (i % 7) for i in range(n)]
    pred = RecordingPredictor()
    batch_calls = []

    def batch_tok(ts):
        batch_calls.append(len(ts))
        return [_tok(t) for t in ts]

    aux = P.B200AuxModel(pred, _tok, max_length=9, batch_tokenizer=batch_tok if batched else None,
                         chunk_prompts=chunk)
    groups = [Group(t) for t in texts]
    ret = aux.obtain_aux_scores(groups)
    assert ret == _want(texts, 9)
    assert [g.aux_model_score for g in groups] == ret
    # the predictor saw every prompt once, in request order, truncated to max_length
    seen = [p for call in pred.calls for p in call]
    assert seen == [_tok(t)[:9] for t in texts]
    assert all(len(c) <= chunk for c in pred.calls)
    # every forward is issued from the calling thread (CUDA calls never move to the helper)
    assert pred.threads <= {threading.get_ident()}
    if batched:
        # a chunk of one text takes the single-text tokenizer (same ids, no batch call)
        sizes = [min(chunk, n - c) for c in range(0, n, chunk)]
        assert batch_calls == [k for k in sizes if k > 1]


def test_scored_group_is_rejected():
    aux = P.B200AuxModel(RecordingPredictor(), _tok)
    g = Group("abc")
    g.set_aux_model_score(1.0)
    with pytest.raises(AssertionError):
        aux.obtain_aux_scores([g])


def test_tokenizer_error_propagates_from_helper_thread():
    def bad(ts):
        raise ValueError("tokenizer failed")

    aux = P.B200AuxModel(RecordingPredictor(), _tok, batch_tokenizer=bad, chunk_prompts=2)
    with pytest.raises(ValueError, match="tokenizer failed"):
        aux.obtain_aux_scores([Group("ab"), Group("cd"), Group("ef"), Group("gh"), Group("ij")])


def test_batch_encoder_equals_encode_one_by_one(tmp_path):
    """The raw ``encode_batch`` route gives ``hf.encode(text)`` for every text; a backend
    tokenizer that carries truncation / padding settings goes through ``hf(texts)`` instead."""
    transformers = pytest.importorskip("transformers")
    pytest.importorskip("tokenizers")
    import numpy as np

    import engine_trace as ET
    d = str(tmp_path / "tok")
    os.makedirs(d)
    ET._write_tokenizer(d, 5000, 2, 2, prefix_bos=True)
    hf = transformers.AutoTokenizer.from_pretrained(d)
    rng = np.random.default_rng(0)
    texts = [" ".join(f"t{t}" for t in rng.integers(3, 5000, int(n)))
             for n in rng.integers(1, 300, 64)] + ["", "t5 unknownword t7"]
    enc = P.B200AuxModel._batch_encoder(hf)
    assert enc is not None
    want = [hf.encode(t) for t in texts]
    assert enc(texts) == want
    assert all(isinstance(p, list) for p in enc(texts))
    hf.backend_tokenizer.enable_truncation(max_length=8)
    enc2 = P.B200AuxModel._batch_encoder(hf)
    got = enc2(texts)
    assert got == [hf.encode(t) for t in texts]
