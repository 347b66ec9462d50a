"""oracle/opt_oracle.py - fp32 CPU restatement of the LTR predictor forward.  TEST INFRASTRUCTURE
(only tests/, smoke() and bench.py's reference legs import it).

Follows the reference's OPTForSequenceClassification (vllm/model_executor/models/opt.py):
decoder forward :283-310 (embed_tokens, embed_positions with the +2 offset :43-55, project_in/out),
decoder layer :146-178 (pre-LN for do_layer_norm_before, post-LN otherwise; ReLU MLP),
attention :94-98 (causal softmax(q.k^T * head_dim^-0.5).v per prompt - the flash_attn_varlen
prefill path), compute_logits :389-397 (last token of each prompt . score.weight^T, argmax for
classification heads).  Everything in float32/float64, one prompt at a time.

Pinning: the reference has no test of the predictor ("parity unpinned" by the reference's own
tests).  This restatement is pinned against HF transformers'
OPTForSequenceClassification - the class train/trainer.py:214-226 saves the checkpoints from - in
tests/test_oracle_predictor.py, on seeded random weights of the OPT-125M / OPT-350M layouts.
"""
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def _norm_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[len("model."):] if k.startswith("model.") else k): v.detach().to(torch.float32)
            for k, v in sd.items()}


def opt_logits(config: Dict, state_dict: Dict[str, torch.Tensor],
               prompts: Sequence[Sequence[int]], dtype=torch.float64) -> torch.Tensor:
    """[len(prompts), num_labels] logits in float64 math (returned as float32)."""
    sd = {k: v.to(dtype) for k, v in _norm_keys(state_dict).items()}
    H = config["hidden_size"]
    heads = config["num_attention_heads"]
    hd = H // heads
    pre_ln = bool(config.get("do_layer_norm_before", True))
    final_ln = pre_ln and not config.get("_remove_final_layer_norm", False)
    out = []
    for ids in prompts:
        ids = torch.as_tensor(list(ids), dtype=torch.long)
        L = ids.numel()
        x = sd["decoder.embed_tokens.weight"][ids]
        if "decoder.project_in.weight" in sd:
            x = x @ sd["decoder.project_in.weight"].T
        h = x + sd["decoder.embed_positions.weight"][torch.arange(L) + 2]
        mask = torch.full((L, L), float("-inf"), dtype=dtype).triu(1)
        for i in range(config["num_hidden_layers"]):
            p = f"decoder.layers.{i}."
            ln = lambda t, n: F.layer_norm(t, (H,), sd[p + n + ".weight"], sd[p + n + ".bias"], 1e-5)
            lin = lambda t, n: t @ sd[p + n + ".weight"].T + sd[p + n + ".bias"]
            resid = h
            a = ln(h, "self_attn_layer_norm") if pre_ln else h
            q = lin(a, "self_attn.q_proj").view(L, heads, hd).transpose(0, 1)
            k = lin(a, "self_attn.k_proj").view(L, heads, hd).transpose(0, 1)
            v = lin(a, "self_attn.v_proj").view(L, heads, hd).transpose(0, 1)
            s = (q @ k.transpose(1, 2)) * hd ** -0.5 + mask
            o = (torch.softmax(s, dim=-1) @ v).transpose(0, 1).reshape(L, H)
            h = resid + lin(o, "self_attn.out_proj")
            if not pre_ln:
                h = ln(h, "self_attn_layer_norm")
            resid = h
            a = ln(h, "final_layer_norm") if pre_ln else h
            h = resid + lin(torch.relu(lin(a, "fc1")), "fc2")
            if not pre_ln:
                h = ln(h, "final_layer_norm")
        last = h[-1]
        if final_ln:
            last = F.layer_norm(last, (H,), sd["decoder.final_layer_norm.weight"],
                                sd["decoder.final_layer_norm.bias"], 1e-5)
        if "decoder.project_out.weight" in sd:
            last = last @ sd["decoder.project_out.weight"].T
        out.append(last @ sd["score.weight"].T)
    return torch.stack(out).to(torch.float32)


def opt_scores(config, state_dict, prompts, score_dtype=torch.float16) -> List[float]:
    """What the scheduler receives: logits rounded to the model dtype, column 0 - or the argmax
    label as float for classification heads (opt.py:389-409)."""
    lg = opt_logits(config, state_dict, prompts).to(score_dtype)
    if config.get("num_labels", 1) > 1:
        return lg.argmax(dim=-1).float().tolist()
    return lg[:, 0].float().tolist()


def random_opt(kind: str, seed: int = 0, num_labels: int = 1, vocab: int = 50272,
               max_pos: int = 2048, layers=None):
    """(config, state_dict) with seeded random weights in the exact OPT-125M / OPT-350M layouts
    (train/configs: facebook/opt-125m, facebook/opt-350m).  Weights are N(0, 0.02) like HF's
    initializer, LayerNorm weights near 1; values are exactly representable in fp16."""
    if kind == "125m":
        cfg = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, ffn_dim=3072,
                   word_embed_proj_dim=768, do_layer_norm_before=True)
    elif kind == "350m":
        cfg = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, ffn_dim=4096,
                   word_embed_proj_dim=512, do_layer_norm_before=False)
    else:
        raise ValueError(kind)
    if layers is not None:
        cfg["num_hidden_layers"] = layers
    cfg.update(vocab_size=vocab, max_position_embeddings=max_pos, num_labels=num_labels,
               activation_function="relu", _remove_final_layer_norm=False, enable_bias=True,
               layer_norm_elementwise_affine=True, pad_token_id=1)
    g = torch.Generator().manual_seed(seed)
    H, E, Fd = cfg["hidden_size"], cfg["word_embed_proj_dim"], cfg["ffn_dim"]

    def w(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).to(torch.float16).to(torch.float32)

    sd = {"model.decoder.embed_tokens.weight": w(vocab, E),
          "model.decoder.embed_positions.weight": w(max_pos + 2, H),
          "score.weight": w(num_labels, E, std=0.05)}
    if E != H:
        sd["model.decoder.project_in.weight"] = w(H, E)
        sd["model.decoder.project_out.weight"] = w(E, H)
    if cfg["do_layer_norm_before"]:
        sd["model.decoder.final_layer_norm.weight"] = 1 + w(H, std=0.1)
        sd["model.decoder.final_layer_norm.bias"] = w(H)
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.decoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = w(H, H, std=0.04)
            sd[p + f"self_attn.{n}.bias"] = w(H)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = w(Fd, H, std=0.04), w(Fd)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = w(H, Fd, std=0.04), w(H)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = 1 + w(H, std=0.1), w(H)
    return cfg, sd
