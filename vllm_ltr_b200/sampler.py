"""Next-token sampling for one engine step on the B200 (csrc/sampler.cu, ``vltr_sample``).

Mirror of ``vllm.model_executor.layers.sampler.Sampler`` (sampler.py:17-90) for the sampling
parameters the reference's throughput benchmarks issue (``benchmark_throughput_original.py:251-259``:
``n=1, temperature=1.0 | 0.0, top_p=1.0, ignore_eos``): greedy, or random with a temperature.  The
reference turns the logits into two fp32 [rows, vocab] matrices (softmax, log_softmax), draws a
third (exponential noise) and reduces them with a dozen launches, after building padded
prompt/output token matrices on the host for penalties nobody asked for
(sampling_metadata.py:95-200, :207-330).  Here one kernel reads each logits row once.

``sample_rows``   tensor-level entry (tests, bench): tokens, logprobs, ranks for [rows, vocab] logits.
``B200Sampler``   drop-in for the engine: ``model.sampler = B200Sampler(reference=old_sampler)``;
                  same ``forward`` signature and ``SamplerOutput``.  Greedy / temperature steps
                  are one kernel.  Steps with penalties, top-k / top-p / min-p, min_tokens or
                  top-N logprobs first apply the reference's own logits edits, then the same
                  kernel (``_forward_general``).  Beam search, best_of > 1, per-request seeds and
                  prompt_logprobs hand the step, whole, to the reference ``Sampler`` the engine
                  was built with (counted in ``.reference_steps``); without one it raises
                  ``NotImplementedError``.
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
    kernel.  The Philox offset advances by one per call, so steps draw independent noise.  The
    default seed is derived from torch's global seed (``torch.initial_seed()``), so
    ``torch.manual_seed`` / the engine's ``seed`` argument make runs reproducible like they do for
    the reference (which draws from torch's global generator); pass ``seed`` to pin it."""

    def __init__(self, seed: Optional[int] = None, reference: Optional[torch.nn.Module] = None):
        super().__init__()
        self.seed = (torch.initial_seed() & (2 ** 63 - 1)) if seed is None else seed
        self.calls = 0
        self.reference = reference
        self.reference_steps = 0
        self._stage = None           # pinned fp32: temperatures | random flags of one step

    @staticmethod
    def _unsupported(i, sp, sampling_metadata) -> Optional[str]:
        """Requests only the engine's own Sampler handles (the step is handed to it whole)."""
        from vllm.sampling_params import SamplingType
        if sp.sampling_type not in (SamplingType.GREEDY, SamplingType.RANDOM):
            return f"sampling_type {sp.sampling_type!r}"     # beam search, per-request seeds
        if sp.prompt_logprobs is not None:
            return "prompt_logprobs"
        if i < sampling_metadata.num_prompts and sp.best_of > 1:
            return "best_of > 1"
        return None

    @staticmethod
    def _needs_tensor_program(sp) -> bool:
        """Penalties, top-k / top-p / min-p, min_tokens or top-N logprobs: the logits are first
        edited by the reference's own tensor functions (sampler.py:104-224), then sampled by the
        fused kernel; logprobs > 0 additionally materialises the fp32 log-prob matrix."""
        return (sp.top_p < 1.0 - _SAMPLING_EPS or sp.top_k != -1 or sp.min_p > _SAMPLING_EPS
                or abs(sp.presence_penalty) >= _SAMPLING_EPS
                or abs(sp.frequency_penalty) >= _SAMPLING_EPS
                or abs(sp.repetition_penalty - 1.0) >= _SAMPLING_EPS
                or getattr(sp, "min_tokens", 0) > 0
                or (sp.logprobs is not None and sp.logprobs > 0))

    def _staged(self, temps: List[float], rnd: List[int], device):
        """temperatures / random flags through one pinned buffer and one non-blocking copy."""
        n = len(temps)
        if self._stage is None or self._stage.numel() < 2 * n:
            self._stage = torch.empty(2 * max(n, 1024), dtype=torch.float32)
            if torch.cuda.is_available():
                self._stage = self._stage.pin_memory()
        self._stage[:n] = torch.tensor(temps, dtype=torch.float32)
        self._stage[n:2 * n] = torch.tensor(rnd, dtype=torch.float32)
        d = self._stage[:2 * n].to(device, non_blocking=True)   # forward() syncs before returning
        return d[:n], d[n:].to(torch.uint8)

    def forward(self, logits: torch.Tensor, sampling_metadata, pred_scores=None,
                aux_model_scores=None):
        from vllm.model_executor.layers.sampler import _build_sampler_output
        from vllm.sampling_params import SamplingType
        from vllm.sequence import Logprob
        assert logits is not None
        temps: List[float] = []
        rnd: List[int] = []
        general = False
        for i, (seq_ids, sp) in enumerate(sampling_metadata.seq_groups):
            bad = self._unsupported(i, sp, sampling_metadata)
            if bad is not None:
                if self.reference is None:
                    raise NotImplementedError(
                        f"vllm_ltr_b200.B200Sampler does not sample requests with {bad}; "
                        f"construct it with reference=<the engine's Sampler>")
                self.reference_steps += 1
                return self.reference(logits, sampling_metadata, pred_scores=pred_scores,
                                      aux_model_scores=aux_model_scores)
            general = general or self._needs_tensor_program(sp)
            t = sp.temperature if sp.temperature >= _SAMPLING_EPS else 1.0   # sampling_metadata.py:133-137
            temps += [t] * len(seq_ids)
            rnd += [int(sp.sampling_type == SamplingType.RANDOM)] * len(seq_ids)
        assert len(temps) == logits.shape[0]
        if general:
            return self._forward_general(logits, sampling_metadata, rnd, pred_scores,
                                         aux_model_scores)
        all_unit = all(t == 1.0 for t in temps)
        td = rd = None
        if not all_unit or any(rnd):
            td, rd = self._staged(temps, rnd, logits.device)
        tokens, lps, ranks, _, _ = sample_rows(
            logits, None if all_unit else td, rd if any(rnd) else None,
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

    def _forward_general(self, logits, sampling_metadata, rnd, pred_scores, aux_model_scores):
        """Steps with penalties / top-k / top-p / min-p / min_tokens / top-N logprobs.  The logits
        edits are the reference's own functions, called in the reference's order (sampler.py:
        47-75: min-tokens penalty, penalties, temperature, top-k/top-p, min-p); the softmax /
        log-softmax / draw / rank that follow (sampler.py:77-90) are the fused kernel, and the
        top-N dictionaries come from the reference's _get_logprobs over the kernel's fp32
        log-prob matrix (only when some request asks for logprobs > 0)."""
        from vllm.model_executor.layers import sampler as ref
        from vllm.model_executor.sampling_metadata import SamplingTensors
        from vllm.sequence import Logprob
        self.general_steps = getattr(self, "general_steps", 0) + 1
        _, vocab = logits.shape
        logits = ref._apply_min_tokens_penalty(logits, sampling_metadata)
        st, do_pen, do_topk, do_minp = SamplingTensors.from_sampling_metadata(
            sampling_metadata, vocab, logits.device, logits.dtype)
        if do_pen:
            logits = ref._apply_penalties(logits, st.prompt_tokens, st.output_tokens,
                                          st.presence_penalties, st.frequency_penalties,
                                          st.repetition_penalties)
        logits.div_(st.temperatures.unsqueeze_(dim=1))
        if do_topk:
            logits = ref._apply_top_k_top_p(logits, st.top_ps, st.top_ks)
        if do_minp:
            logits = ref._apply_min_p(logits, st.min_ps)
        want_lp = any(sp.logprobs is not None and sp.logprobs > 0
                      for _, sp in sampling_metadata.seq_groups)
        rd = None
        if any(rnd):
            rd = torch.tensor(rnd, dtype=torch.uint8).to(logits.device, non_blocking=True)
        if not logits.is_contiguous():
            logits = logits.contiguous()
        tokens, lps, ranks, _, lpf = sample_rows(logits, None, rd, seed=self.seed,
                                                 offset=self.calls, want_logprobs=want_lp)
        self.calls += 1
        packed = torch.stack((tokens.to(torch.float64), lps.to(torch.float64),
                              ranks.to(torch.float64))).cpu()
        toks = packed[0].to(torch.int64).tolist()
        lpl, rkl = packed[1].tolist(), packed[2].to(torch.int64).tolist()
        results, row = [], 0
        for seq_ids, _ in sampling_metadata.seq_groups:
            n = len(seq_ids)
            results.append((toks[row:row + n], list(range(n))))
            row += n
        if want_lp:
            prompt_lp, sample_lp = ref._get_logprobs(lpf, sampling_metadata, results)
        else:
            prompt_lp = [None] * len(results)
            sample_lp, row = [], 0
            for seq_ids, _ in sampling_metadata.seq_groups:
                n = len(seq_ids)
                sample_lp.append([{toks[row + j]: Logprob(lpl[row + j], rkl[row + j])}
                                  for j in range(n)])
                row += n
        return ref._build_sampler_output(results, sampling_metadata, prompt_lp, sample_lp,
                                         pred_scores, aux_model_scores)
