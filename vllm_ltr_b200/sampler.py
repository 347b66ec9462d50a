"""Next-token sampling for one engine step on the B200 (csrc/sampler.cu, ``vltr_sample``).

Mirror of ``vllm.model_executor.layers.sampler.Sampler`` (sampler.py:17-90) for the sampling
parameters the reference's throughput benchmarks issue (``benchmark_throughput_original.py:251-259``:
``n=1, temperature=1.0 | 0.0, top_p=1.0, ignore_eos``): greedy, or random with a temperature.  The
reference turns the logits into two fp32 [rows, vocab] matrices (softmax, log_softmax), draws a
third (exponential noise) and reduces them with a dozen launches, after building padded
prompt/output token matrices on the host for penalties nobody asked for
(sampling_metadata.py:95-200, :207-330).  Here one kernel reads each logits row once.

``sample_rows``   tensor-level entry (tests, bench): tokens, logprobs, ranks for [rows, vocab] logits.
``B200Sampler``   drop-in for the engine: ``model.sampler = B200Sampler()``; same ``forward``
                  signature and ``SamplerOutput``; requests outside the subset raise
                  ``NotImplementedError`` (there is no silent fallback).
"""
import ctypes
from typing import List, Optional, Tuple

import torch

from ._lib import check, declare, lib

_P = ctypes.c_void_p
declare("vltr_sample", ctypes.c_int,
        [_P, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P, _P, ctypes.c_uint64,
         ctypes.c_uint64, _P, _P, _P, _P, _P, _P])

_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_SAMPLING_EPS = 1e-5      # vllm/model_executor/sampling_metadata.py:12


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def sample_rows(logits: torch.Tensor, temperatures: Optional[torch.Tensor] = None,
                random: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                seed: int = 0, offset: int = 0, want_probs: bool = False,
                want_logprobs: bool = False):
    """logits [R, V] (f32/f16/bf16, unit column stride).  temperatures [R] (any float dtype; used
    after rounding to the logits dtype, as the reference's temperature tensor is), random [R]
    bool/uint8 (None = greedy everywhere), noise [R, V] fp32 Exp(1) samples (None = Philox in the
    kernel, keyed by seed/offset).  Returns (tokens int64 [R], logprobs f32 [R], ranks int64 [R],
    probs f32 [R, V] | None, logprobs_full f32 [R, V] | None)."""
    if not logits.is_cuda:
        raise RuntimeError("vllm_ltr_b200 sampler: logits must be a CUDA tensor (no CPU path)")
    if logits.dim() != 2 or logits.stride(1) != 1 or logits.dtype not in _DTYPE:
        raise RuntimeError("vllm_ltr_b200 sampler: logits must be [rows, vocab] f32/f16/bf16 with "
                           "unit column stride")
    R, V = logits.shape
    dev = logits.device
    t = None
    if temperatures is not None:
        t = temperatures.to(device=dev, dtype=logits.dtype).to(torch.float32).contiguous()
        assert t.numel() == R
    r = None
    if random is not None:
        r = random.to(device=dev, dtype=torch.uint8).contiguous()
        assert r.numel() == R
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        assert noise.shape == (R, V)
    tokens = torch.empty(R, dtype=torch.int64, device=dev)
    lp = torch.empty(R, dtype=torch.float32, device=dev)
    rank = torch.empty(R, dtype=torch.int64, device=dev)
    probs = torch.empty(R, V, dtype=torch.float32, device=dev) if want_probs else None
    lpf = torch.empty(R, V, dtype=torch.float32, device=dev) if want_logprobs else None
    with torch.cuda.device(dev):
        check(lib.vltr_sample(logits.data_ptr(), logits.stride(0), R, V, _DTYPE[logits.dtype],
                              _ptr(t), _ptr(r), _ptr(noise), seed & (2 ** 64 - 1),
                              offset & (2 ** 64 - 1), tokens.data_ptr(), lp.data_ptr(),
                              rank.data_ptr(), _ptr(probs), _ptr(lpf),
                              torch._C._cuda_getCurrentRawStream(dev.index)))
    return tokens, lp, rank, probs, lpf


class B200Sampler(torch.nn.Module):
    """``Sampler.forward(logits, sampling_metadata, pred_scores, aux_model_scores)`` on the fused
    kernel.  The Philox offset advances by one per call, so steps draw independent noise; pass
    ``seed`` for reproducible runs (the reference uses torch's global generator)."""

    def __init__(self, seed: int = 0):
        super().__init__()
        self.seed = seed
        self.calls = 0

    @staticmethod
    def _check(i, sp, sampling_metadata):
        from vllm.sampling_params import SamplingType
        bad = None
        if sp.sampling_type not in (SamplingType.GREEDY, SamplingType.RANDOM):
            bad = f"sampling_type {sp.sampling_type!r}"
        elif sp.prompt_logprobs is not None:
            bad = "prompt_logprobs"
        elif sp.logprobs is not None and sp.logprobs > 0:
            bad = "logprobs > 0"
        elif sp.top_p < 1.0 - _SAMPLING_EPS or sp.top_k != -1 or sp.min_p > _SAMPLING_EPS:
            bad = "top_p / top_k / min_p"
        elif (abs(sp.presence_penalty) >= _SAMPLING_EPS or abs(sp.frequency_penalty) >= _SAMPLING_EPS
              or abs(sp.repetition_penalty - 1.0) >= _SAMPLING_EPS):
            bad = "presence / frequency / repetition penalty"
        elif getattr(sp, "min_tokens", 0) > 0:
            bad = "min_tokens"
        elif i < sampling_metadata.num_prompts and sp.best_of > 1:
            bad = "best_of > 1"
        if bad is not None:
            raise NotImplementedError(
                f"vllm_ltr_b200.B200Sampler handles greedy / temperature sampling only; request "
                f"uses {bad} - keep the reference Sampler for it")

    def forward(self, logits: torch.Tensor, sampling_metadata, pred_scores=None,
                aux_model_scores=None):
        from vllm.model_executor.layers.sampler import _build_sampler_output
        from vllm.sampling_params import SamplingType
        from vllm.sequence import Logprob
        assert logits is not None
        temps: List[float] = []
        rnd: List[int] = []
        for i, (seq_ids, sp) in enumerate(sampling_metadata.seq_groups):
            self._check(i, sp, sampling_metadata)
            t = sp.temperature if sp.temperature >= _SAMPLING_EPS else 1.0   # sampling_metadata.py:133-137
            temps += [t] * len(seq_ids)
            rnd += [int(sp.sampling_type == SamplingType.RANDOM)] * len(seq_ids)
        assert len(temps) == logits.shape[0]
        all_unit = all(t == 1.0 for t in temps)
        tokens, lps, ranks, _, _ = sample_rows(
            logits, None if all_unit else torch.tensor(temps, dtype=torch.float32),
            torch.tensor(rnd, dtype=torch.uint8) if any(rnd) else None,
            seed=self.seed, offset=self.calls)
        self.calls += 1
        packed = torch.stack((tokens.to(torch.float64), lps.to(torch.float64),
                              ranks.to(torch.float64))).cpu()          # one D2H for the step
        toks = packed[0].to(torch.int64).tolist()
        lpl, rkl = packed[1].tolist(), packed[2].to(torch.int64).tolist()
        results: List[Tuple[List[int], List[int]]] = []
        sample_logprobs = []
        row = 0
        for seq_ids, sp in sampling_metadata.seq_groups:
            n = len(seq_ids)
            results.append((toks[row:row + n], list(range(n))))
            sample_logprobs.append([{toks[row + j]: Logprob(lpl[row + j], rkl[row + j])}
                                    for j in range(n)])
            row += n
        return _build_sampler_output(results, sampling_metadata,
                                     [None] * len(sampling_metadata.seq_groups), sample_logprobs,
                                     pred_scores, aux_model_scores)
