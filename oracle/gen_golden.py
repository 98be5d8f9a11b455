"""Generate tests/golden/*.pt by running the REFERENCE ITSELF (test infrastructure).

Runs only in the build container: imports /root/reference/src/model.py (FiDT5) unmodified, with
the 2-line external shim SURVEY.md §8(c) documents (transformers 5.5.0 passes extra positionals to
CheckpointWrapper.forward; EncoderWrapper lacks main_input_name), feeds it seeded synthetic inputs
and the weights produced by oracle.fidt5_oracle.init_state_dict, and stores the reference's
outputs.  tests/test_oracle_golden.py then pins the oracle restatement to these files; the GPU
parity tests compare the CUDA path with the oracle and with these same files.

    python oracle/gen_golden.py            # rewrites tests/golden/
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import transformers  # noqa: E402
import src.model as ref_model  # noqa: E402  (the reference)

from oracle import fidt5_oracle as O  # noqa: E402

# --- the shim ---------------------------------------------------------------------
ref_model.CheckpointWrapper.forward = lambda self, h, *a, **k: self.module(h, *a, **k)


def build_reference(cfg: O.T5Dims, sd):
    hf_cfg = transformers.T5Config(
        vocab_size=cfg.vocab, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff,
        num_layers=cfg.n_layers, num_decoder_layers=cfg.n_layers, num_heads=cfg.n_heads,
        relative_attention_num_buckets=cfg.n_buckets, relative_attention_max_distance=cfg.max_distance,
        dropout_rate=0.0, layer_norm_epsilon=cfg.eps, feed_forward_proj="relu",
        decoder_start_token_id=0, pad_token_id=0, eos_token_id=1,
    )
    model = ref_model.FiDT5(hf_cfg)
    model.encoder.main_input_name = "input_ids"
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    model.eval()
    return model


def install_capture_hooks(model):
    """Restated capture hook following src/model.py:305-329 (the 3.0.2-signature
    cross_attention_forward cannot be patched into 5.5.0): on the first call after a reset store
    q.k^T + additive mask, shape (B,H,qlen,M)."""
    for blk in model.decoder.block:
        attn = blk.layer[1].EncDecAttention
        attn.score_storage = None

        def pre_hook(mod, args, kwargs):
            if mod.score_storage is not None:
                return
            hidden = args[0]
            kv = kwargs["key_value_states"]
            b = hidden.shape[0]
            q = mod.q(hidden).view(b, -1, mod.n_heads, mod.key_value_proj_dim).transpose(1, 2)
            k = mod.k(kv).view(b, -1, mod.n_heads, mod.key_value_proj_dim).transpose(1, 2)
            scores = torch.einsum("bnqd,bnkd->bnqk", q, k)
            mask = kwargs.get("mask", None)
            if mask is not None:
                scores = scores + mask
            mod.score_storage = scores

        attn.register_forward_pre_hook(pre_hook, with_kwargs=True)


def case(name, cfg, B, N, L, T, emb_std, seed, max_length, n_context, fact_format=True, eos_alias=None):
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=seed, emb_std=emb_std), seed=seed + 1)
    if eos_alias is not None:
        # make EOS (id 1) win wherever `eos_alias` would have: exercises the EOS -> pad-fill logic
        sd["shared.weight"][1] = sd["shared.weight"][eos_alias] * 1.05
    ids, mask, labels = O.synth_batch(B, N, L, T, cfg.vocab, seed=seed + 2, fact_format=fact_format)
    model = build_reference(cfg, sd)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, labels=labels, return_dict=False)
        loss, logits = out[0], out[1]
    # gradients of the reference (for backward parity of the CUDA path)
    model.zero_grad()
    out = model(input_ids=ids, attention_mask=mask, labels=labels, return_dict=False)
    out[0].backward()
    grads = {}
    for n_, p in model.named_parameters():
        if p.grad is not None:
            grads[n_] = p.grad.detach().clone()
    grad_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    # a few named gradients only (fixtures must stay small); plus the global norm
    keep = [
        "shared.weight",
        "encoder.encoder.block.0.module.layer.0.SelfAttention.relative_attention_bias.weight",
        "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
        "encoder.encoder.block.0.module.layer.0.SelfAttention.q.weight",
        f"encoder.encoder.block.{cfg.n_layers - 1}.module.layer.1.DenseReluDense.wo.weight",
        f"decoder.block.{cfg.n_layers - 1}.layer.1.EncDecAttention.k.weight",
        "decoder.block.0.layer.1.EncDecAttention.q.weight",
        "decoder.block.0.layer.2.DenseReluDense.wi.weight",
        "encoder.encoder.final_layer_norm.weight",
        "decoder.final_layer_norm.weight",
        "decoder.block.0.layer.1.layer_norm.weight",
    ]
    grads_kept = {k: grads[k] for k in keep}

    install_capture_hooks(model)
    model.reset_score_storage()
    with torch.no_grad():
        gen = model.generate(input_ids=ids, attention_mask=mask, max_length=max_length)
    storage = [blk.layer[1].EncDecAttention.score_storage for blk in model.decoder.block]
    res = {"name": name, "cfg": cfg.__dict__, "B": B, "N": N, "L": L, "T": T, "emb_std": emb_std,
           "seed": seed, "max_length": max_length, "n_context": n_context, "eos_alias": eos_alias,
           "ids": ids, "mask": mask, "labels": labels,
           "loss": loss.detach(), "logits": logits.detach(), "generated": gen,
           "grad_norm": grad_norm, "grads": grads_kept,
           "weights_checksum": float(sum(v.double().sum() for v in O.unique_params(sd).values()))}
    if N % 2 == 0:
        for style in ("mean", "max", "21mean"):
            for last_half in ("yes", "no"):
                opt = types.SimpleNamespace(stream=2, n_context=n_context, attention_score_style=style,
                                            use_last_half_layer_attention=last_half)
                sc = model.get_crossattention_scores(opt, ids, None, mask)
                res[f"scores_{style}_{last_half}"] = sc
        tok, div = O.token_scores_from_storage(storage, mask, last_half=True)
        res["tok_scores_last_half"] = tok
        res["tok_scores_div"] = div
    return res


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)
    cases = [
        # LaKo's own layout: N=2 passages (question+caption, facts), stream=2
        case("tiny_lako_n2", O.T5_TINY, B=3, N=2, L=48, T=6, emb_std=0.3, seed=10, max_length=8, n_context=5,
             eos_alias=121),
        # FiD-style multi passage, ragged lengths, longer memory
        case("tiny_fid_n4", O.T5_TINY, B=2, N=4, L=64, T=10, emb_std=0.3, seed=20, max_length=10, n_context=4),
        # single passage (stream=1 / no facts), odd L
        case("tiny_n1", O.T5_TINY, B=2, N=1, L=40, T=5, emb_std=0.3, seed=30, max_length=6, n_context=1,
             fact_format=False),
    ]
    for c in cases:
        path = os.path.join(out_dir, c["name"] + ".pt")
        torch.save(c, path)
        print(path, os.path.getsize(path), "bytes; loss", float(c["loss"]), "gen", c["generated"].tolist())


if __name__ == "__main__":
    main()
