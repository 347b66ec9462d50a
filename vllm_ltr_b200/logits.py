"""lm_head logits of one engine step on the tcgen05 dense layer (``vltr_lm_head_logits``).

Mirror of ``LogitsProcessor._get_logits`` (vllm/model_executor/layers/logits_processor.py:61-71):

    logits = torch.matmul(hidden_states, embedding.t())      # [rows, vocab shard]
    if embedding_bias is not None: logits += embedding_bias
    logits = tensor_model_parallel_gather(logits)             # rank 0 gets [rows, padded vocab]
    if logits is not None: logits = logits[:, :self.org_vocab_size]

``lm_head_logits``         tensor-level entry (tests, bench).
``get_logits``             the replacement method: same arguments, same return value (None on
                           the ranks the gather does not deliver to); the matmul and the bias go
                           through the C ABI, the gather is the engine's own collective.
``install(model)``         binds it on ``model.logits_processor`` (every reference model builds
                           one, e.g. llama.py:349-350).

Rounding differs from the reference in one place: the bias is added in fp32 before the single
rounding of the result (the reference rounds the product, then the sum).  Without a bias (Llama,
OPT lm_head) the two agree to the summation order of the fp32 accumulation.
"""
import ctypes
import types
from typing import Optional

import torch

from ._lib import check, declare, lib

_P = ctypes.c_void_p
declare("vltr_lm_head_logits", ctypes.c_int,
        [_P, _P, _P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])

_DTYPE = {torch.float16: 1, torch.bfloat16: 2}


def lm_head_logits(hidden_states: torch.Tensor, embedding: torch.Tensor,
                   embedding_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hidden_states [rows, hidden] x embedding [vocab_rows, hidden] -> [rows, vocab_rows]."""
    if not hidden_states.is_cuda:
        raise RuntimeError("vllm_ltr_b200 logits: tensors must be on a CUDA device (no CPU path)")
    code = _DTYPE.get(hidden_states.dtype)
    if code is None or embedding.dtype != hidden_states.dtype:
        raise RuntimeError(f"Unsupported data type: {hidden_states.dtype} x {embedding.dtype}")
    if hidden_states.dim() != 2 or embedding.dim() != 2 or \
            hidden_states.shape[1] != embedding.shape[1]:
        raise RuntimeError("vllm_ltr_b200 logits: expected [rows, hidden] x [vocab, hidden]")
    h = hidden_states if hidden_states.is_contiguous() else hidden_states.contiguous()
    w = embedding if embedding.is_contiguous() else embedding.contiguous()
    b = None
    if embedding_bias is not None:
        b = embedding_bias.to(dtype=h.dtype).contiguous()
    rows, hidden = h.shape
    vocab = w.shape[0]
    out = torch.empty(rows, vocab, dtype=h.dtype, device=h.device)
    with torch.cuda.device(h.device):
        check(lib.vltr_lm_head_logits(h.data_ptr(), w.data_ptr(),
                                      None if b is None else b.data_ptr(), out.data_ptr(), rows,
                                      vocab, hidden, code,
                                      torch._C._cuda_getCurrentRawStream(h.device.index)))
    return out


def get_logits(self, hidden_states: torch.Tensor, embedding: torch.Tensor,
               embedding_bias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Drop-in for ``LogitsProcessor._get_logits`` (bound on the instance by ``install``)."""
    from vllm.distributed import tensor_model_parallel_gather
    logits = lm_head_logits(hidden_states, embedding, embedding_bias)
    logits = tensor_model_parallel_gather(logits)
    if logits is not None:
        logits = logits[:, :self.org_vocab_size]
    return logits


def install(model) -> bool:
    """Bind ``get_logits`` on ``model.logits_processor``.  False when the model has none or its
    lm_head is not f16 / bf16 (the dense layer's operand types)."""
    lp = getattr(model, "logits_processor", None)
    if lp is None or not hasattr(lp, "_get_logits"):
        return False
    head = getattr(model, "lm_head", None)
    weight = getattr(head, "weight", None)
    if weight is not None and weight.dtype not in _DTYPE:
        return False
    lp._get_logits = types.MethodType(get_logits, lp)
    return True
