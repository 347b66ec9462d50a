"""CPU: install_engine() wires every seam onto a reference-shaped engine object and reports what
it did; preconditions that do not hold are reported as skipped (never silently substituted)."""
import types

import pytest

from tests import sched_replay as R

pytestmark = pytest.mark.skipif(not R.have_reference(),
                                reason="needs the reference checkout at /root/reference")


def test_install_engine_on_reference_shaped_objects(monkeypatch):
    R.import_reference()
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler

    from vllm_ltr_b200 import decode_inputs, integration, ltr_scheduler
    sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
    cc = CacheConfig(16, 1.0, 1, "auto")
    cc.num_gpu_blocks, cc.num_cpu_blocks = 64, 64
    sched = Scheduler(sc, cc, None)
    # no GPU here: the ranker constructor is the only piece that needs one
    monkeypatch.setattr(ltr_scheduler, "GpuRanker", lambda: R.PythonRanker())
    runner = types.SimpleNamespace(model=types.SimpleNamespace(sampler=object()))
    ce = types.SimpleNamespace(cpu_cache=[], gpu_cache=[], swap_in=None, swap_out=None)
    engine = types.SimpleNamespace(
        scheduler=sched, model_config=types.SimpleNamespace(prefill_predictor_model_config=None),
        model_executor=types.SimpleNamespace(
            driver_worker=types.SimpleNamespace(model_runner=runner, cache_engine=ce)))
    done = integration.install_engine(engine, prefill=False, attention_backend=False)
    assert done["scheduler"].startswith("LTR tick")
    assert sched._schedule.__func__ is ltr_scheduler.general_schedule
    assert done["predictor"].startswith("skipped")
    assert callable(ce.swap_in) and callable(ce.swap_out)
    assert runner._prepare_decode.__func__ is decode_inputs.prepare_decode
    assert type(runner.model.sampler).__name__ == "B200Sampler"

    fcfs = types.SimpleNamespace(scheduler=types.SimpleNamespace(schedule_type="fcfs"),
                                 model_config=None, model_executor=None)
    done = integration.install_engine(fcfs, prefill=False)
    assert done["attention_backend"].startswith("skipped")
    assert done["scheduler"].startswith("skipped") and done["swap"].startswith("skipped")
    assert done["sampler"].startswith("skipped") and done["decode_inputs"].startswith("skipped")


def test_attention_backend_seam(monkeypatch):
    """b2: every Attention layer of a built model gets a B200AttentionImpl carrying the arguments
    its own impl was built with; the backend object keeps the reference's metadata / cache shape."""
    import torch
    R.import_reference()
    monkeypatch.setattr(torch.cuda, "get_device_capability", lambda *a: (10, 0))
    from vllm.attention.backends.abstract import AttentionBackend, AttentionImpl
    from vllm.attention.layer import Attention

    from vllm_ltr_b200 import attention_backend as ab

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = Attention(32, 128, 0.088, num_kv_heads=8)
            self.b = Attention(8, 64, 0.125, alibi_slopes=[0.5] * 8)

    torch.set_default_dtype(torch.float16)       # fp32 models get the xformers backend (absent here)
    try:
        m = Tiny()
    finally:
        torch.set_default_dtype(torch.float32)
    old_a = m.a.impl
    assert ab.install(m) == 2 and ab.install(m) == 0          # idempotent
    backend, impl_cls = ab.backend_classes()
    assert issubclass(backend, AttentionBackend) and issubclass(impl_cls, AttentionImpl)
    assert isinstance(m.a.impl, impl_cls) and m.a.impl.reference_impl is old_a
    assert (m.a.impl.num_heads, m.a.impl.num_kv_heads, m.a.impl.head_size) == (32, 8, 128)
    assert abs(m.a.impl.scale - 0.088) < 1e-9 and m.b.impl.alibi_slopes.shape == (8,)
    assert backend.get_kv_cache_shape(10, 16, 8, 128) == (2, 10, 16 * 8 * 128)
    assert backend.get_impl_cls() is impl_cls
    with pytest.raises(RuntimeError, match="CUDA"):
        m.a.impl.forward(torch.zeros(1, 32 * 128), torch.zeros(1, 8 * 128), torch.zeros(1, 8 * 128),
                         torch.zeros(2, 4, 16 * 8 * 128),
                         __import__("types").SimpleNamespace(
                             kv_cache_dtype="auto", slot_mapping=torch.zeros(1, dtype=torch.long),
                             num_prefill_tokens=0, num_decode_tokens=1, prefill_metadata=None,
                             decode_metadata=None), 1.0)


def test_deferred_rotary(monkeypatch):
    """The rotary embedding of a decoder layer is deferred into the layer's B200AttentionImpl
    (fused with the cache append): install() wraps NeoX full-head RotaryEmbeddings that sit next
    to an Attention layer, leaves everything else alone, and the cache-less profiling run still
    rotates through the reference op."""
    import torch
    R.import_reference()
    monkeypatch.setattr(torch.cuda, "get_device_capability", lambda *a: (10, 0))
    from vllm.attention.layer import Attention
    from vllm.model_executor.layers.rotary_embedding import RotaryEmbedding

    from vllm_ltr_b200 import attention_backend as ab

    class Layer(torch.nn.Module):
        def __init__(self, neox=True, rot=128):
            super().__init__()
            self.rotary_emb = RotaryEmbedding(128, rot, 64, 10000, neox)
            self.attn = Attention(32, 128, 0.088, num_kv_heads=8)

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l0, self.l1, self.l2 = Layer(), Layer(neox=False), Layer(rot=64)

    torch.set_default_dtype(torch.float16)
    try:
        m = Model()
    finally:
        torch.set_default_dtype(torch.float32)
    real = m.l0.rotary_emb
    assert ab.install(m) == 3
    assert isinstance(m.l0.rotary_emb, ab.DeferredRotary) and m.l0.rotary_emb.inner is real
    assert isinstance(m.l1.rotary_emb, RotaryEmbedding)      # GPT-J style: not fused
    assert isinstance(m.l2.rotary_emb, RotaryEmbedding)      # partial rotation: not fused
    assert "rotary_emb.inner.cos_sin_cache" not in "".join(m.state_dict().keys())
    assert ab.install(m) == 0 and m.l0.rotary_emb.inner is real      # idempotent
    ab.install(m, fuse_rope=False)
    # deferral: q / k come back untouched, the impl holds (positions, module) for its next forward
    pos = torch.arange(3)
    q, k = torch.randn(3, 32 * 128), torch.randn(3, 8 * 128)
    dr = ab.DeferredRotary(real, m.l0.attn.impl)
    q2, k2 = dr(pos, q, k)
    assert q2 is q and k2 is k and m.l0.attn.impl._pending_rope[1] is real
    # the profiling run (no KV cache): rotation through the reference op, then the reference impl
    seen = []
    monkeypatch.setattr(real, "forward", lambda p_, q_, k_, o_=None: seen.append("rope") or (q_, k_))
    m.l0.attn.impl.reference_impl = __import__("types").SimpleNamespace(
        forward=lambda *a: seen.append("ref") or "out")
    assert m.l0.attn.impl.forward(q, k, k, None, None, 1.0) == "out"
    assert seen == ["rope", "ref"] and m.l0.attn.impl._pending_rope is None
    # offsets (batched LoRA form) are not deferred
    dr(pos, q, k, torch.zeros(3, dtype=torch.long))
    assert seen == ["rope", "ref", "rope"] and m.l0.attn.impl._pending_rope is None


def test_logits_processor_seam(monkeypatch):
    """f3 (logits half): ``logits.install`` re-binds ``LogitsProcessor._get_logits``; the host
    logic around the matmul (bias, gather, vocabulary trim, None on non-driver ranks) is the
    reference's.  The kernel call is replaced by a torch matmul here (no GPU)."""
    import torch
    R.import_reference()
    import vllm.distributed as vd
    import vllm.model_executor.layers.logits_processor as ref_lp

    from vllm_ltr_b200 import integration, logits as lg

    monkeypatch.setattr(ref_lp, "tensor_model_parallel_gather", lambda x: x)
    monkeypatch.setattr(vd, "tensor_model_parallel_gather", lambda x: x)
    calls = []

    def fake(hidden, emb, bias=None):
        calls.append(tuple(hidden.shape))
        out = hidden.float() @ emb.float().t()
        if bias is not None:
            out = out + bias.float()
        return out.to(hidden.dtype)

    monkeypatch.setattr(lg, "lm_head_logits", fake)
    torch.manual_seed(0)
    hidden = torch.randn(5, 64).to(torch.bfloat16)
    emb = torch.randn(48, 64).to(torch.bfloat16)
    bias = torch.randn(48).to(torch.bfloat16)
    ref = ref_lp.LogitsProcessor(48, org_vocab_size=41)
    ours = ref_lp.LogitsProcessor(48, org_vocab_size=41)
    model = types.SimpleNamespace(logits_processor=ours,
                                  lm_head=types.SimpleNamespace(weight=emb))
    assert lg.install(model)
    for b in (None, bias):
        want = ref._get_logits(hidden, emb, b)
        got = ours._get_logits(hidden, emb, b)
        assert got.shape == want.shape == (5, 41)
        assert torch.allclose(got.float(), want.float(), atol=0.13, rtol=2e-2)
    assert calls == [(5, 64), (5, 64)]
    # a rank the gather does not deliver to gets None, as in the reference
    monkeypatch.setattr(vd, "tensor_model_parallel_gather", lambda x: None)
    assert ours._get_logits(hidden, emb, None) is None
    # models without an f16 / bf16 lm_head are left alone
    assert not lg.install(types.SimpleNamespace(
        logits_processor=ref_lp.LogitsProcessor(48),
        lm_head=types.SimpleNamespace(weight=emb.float())))
    assert not lg.install(types.SimpleNamespace())
    # install_engine reports the seam
    runner = types.SimpleNamespace(model=model)
    engine = types.SimpleNamespace(
        scheduler=None, model_config=None,
        model_executor=types.SimpleNamespace(
            driver_worker=types.SimpleNamespace(model_runner=runner, cache_engine=None)))
    done = integration.install_engine(engine, scheduler=False, predictor=False, swap=False,
                                      decode_inputs=False, sampler=False, prefill=False,
                                      attention_backend=False)
    assert done["logits"].startswith("LogitsProcessor._get_logits")
