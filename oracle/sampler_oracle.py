"""CPU restatement of the reference sampler's tensor program (TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU legs - never by the product).

Follows vllm/model_executor/layers/sampler.py for requests without penalties / top-k / top-p /
min-p:
  :67       logits.div_(temperatures.unsqueeze(1))     temperatures is a tensor of the logits dtype
  :77, :80  probs = softmax(logits, fp32); logprobs = log_softmax(logits, fp32)
  :395      greedy: argmax(logprobs)
  :333-357  random: q ~ Exp(1) per element; argmax(probs / q)
  :528-544  rank = (logprobs > logprobs[tok]).sum() + 1
  :594-600  logprob = logprobs[row, tok]
Pinned by tests/golden/sampler_ref.npz = outputs of the reference's own Sampler.forward run on CPU
with its exponential noise replaced by a recorded array (tests/golden/gen_golden_sampler.py).
"""
import torch


def sample_rows(logits: torch.Tensor, temperatures=None, random=None, noise=None):
    """logits [R, V] (any float dtype, CPU).  Returns tokens [R] int64, logprobs [R] f32,
    ranks [R] int64, probs [R, V] f32, logprobs_full [R, V] f32."""
    x = logits.clone()
    R, V = x.shape
    if temperatures is not None:
        x.div_(torch.as_tensor(temperatures).to(x.dtype).unsqueeze(1))
    probs = torch.softmax(x, dim=-1, dtype=torch.float)
    logprobs = torch.log_softmax(x, dim=-1, dtype=torch.float)
    tokens = torch.argmax(logprobs, dim=-1)
    if random is not None and bool(torch.as_tensor(random).any()):
        assert noise is not None
        rnd = torch.as_tensor(random).bool()
        drawn = (probs / noise.to(torch.float)).argmax(dim=1)
        tokens = torch.where(rnd, drawn, tokens)
    lp = logprobs[torch.arange(R), tokens]
    ranks = (logprobs > lp[:, None]).long().sum(1) + 1
    return tokens, lp, ranks, probs, logprobs
