"""End-to-end parity of lako_b200.FiDT5 (CUDA, bf16 operands / fp32 accumulation) against
(a) the golden outputs of the reference itself (tests/golden/*.pt) and (b) the fp32 oracle."""
import glob
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))

# tolerances from BASELINE.json north_star: logits 2e-2 abs (bf16), cross-attention scores 1e-2 relative.
# The 128-wide toy model of the golden fixtures averages bf16 rounding over 6-8x fewer terms than
# t5-base/large, so its score bound is loosened to 2e-2 here; the 1e-2 bound is asserted on the
# t5-base-shaped case in test_gpu_parity_base.py.
LOGIT_ATOL_BF16 = 2e-2
SCORE_RTOL = 2e-2


def _load(path):
    from oracle import fidt5_oracle as O
    from lako_b200.model import FiDT5, FiDT5Config
    c = torch.load(path, weights_only=False)
    cfg = O.T5Dims(**c["cfg"])
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=c["seed"], emb_std=c["emb_std"]), seed=c["seed"] + 1)
    if c.get("eos_alias") is not None:
        sd["shared.weight"][1] = sd["shared.weight"][c["eos_alias"]] * 1.05
    model = FiDT5(FiDT5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.n_layers,
                              num_heads=cfg.n_heads, dropout_rate=0.0))
    model.load_state_dict(sd)
    model = model.cuda()
    return c, cfg, sd, model


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_forward_loss_logits_vs_reference(path):
    c, cfg, sd, model = _load(path)
    model.eval()
    with torch.no_grad():
        loss, logits = model(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    err = (logits.cpu() - c["logits"]).abs().max().item()
    print("max |dlogit|", err, "loss", loss.item(), "ref", c["loss"].item())
    assert err < LOGIT_ATOL_BF16
    assert abs(loss.item() - c["loss"].item()) < 5e-3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_backward_vs_reference(path):
    c, cfg, sd, model = _load(path)
    model.train()
    model.zero_grad()
    loss, _ = model(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    named = dict(model.named_parameters())
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in named.values())).item()
    print("grad norm", gn, "ref", c["grad_norm"].item())
    assert abs(gn - c["grad_norm"].item()) < 0.02 * c["grad_norm"].item()
    for k, g in c["grads"].items():
        mine = named[k].grad.cpu()
        rel = (mine - g).norm().item() / (g.norm().item() + 1e-12)
        print(f"{k}: relative L2 err {rel:.4f}")
        # yardstick: stock PyTorch CPU autocast(bf16) of the oracle deviates from its own fp32 run by
        # 0.025 (shared), 0.060-0.069 (encoder bias table / q weight), 0.03-0.045 (others) on these cases
        # (tools/noise_floor.py); the CUDA path must stay at that noise floor.
        assert rel < 0.08, (k, rel)
    # second backward accumulates (loss.backward() semantics)
    loss2, _ = model(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    loss2.backward()
    torch.cuda.synchronize()
    gn2 = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in named.values())).item()
    assert abs(gn2 - 2 * gn) < 0.01 * gn


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_generate_and_scores_vs_reference(path):
    c, cfg, sd, model = _load(path)
    model.eval()
    model.overwrite_forward_crossattention()
    model.reset_score_storage()
    out = model.generate(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), max_length=c["max_length"])
    assert out.cpu().tolist() == c["generated"].tolist()
    if c["N"] % 2 == 0:
        tok, div = model.token_scores(last_half=True)
        ref = c["tok_scores_last_half"]
        B, N, L = c["ids"].shape
        err = (tok.cpu().view(B, N, L) - ref).abs().max().item()
        assert err < SCORE_RTOL * ref.abs().max().item(), err
        for style in ("mean", "max", "21mean"):
            for lh in ("yes", "no"):
                opt = types.SimpleNamespace(stream=2, n_context=c["n_context"], attention_score_style=style,
                                            use_last_half_layer_attention=lh)
                sc = model.get_crossattention_scores(opt, c["ids"], None, c["mask"].cuda())
                ref = c[f"scores_{style}_{lh}"]
                assert sc.dtype == torch.float64 and sc.shape == ref.shape
                real = ref != (-5 / (div if lh == "yes" else cfg.n_layers * cfg.n_heads))
                assert torch.allclose(sc, ref, rtol=SCORE_RTOL, atol=SCORE_RTOL * ref[real].abs().max().item())


def test_state_dict_roundtrip_and_t5_names(tmp_path):
    from oracle import fidt5_oracle as O
    from lako_b200.model import FiDT5, FiDT5Config
    cfg = O.T5_TINY
    sd = O.init_state_dict(cfg, seed=3, emb_std=0.3)
    conf = FiDT5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.n_layers,
                       num_heads=cfg.n_heads, dropout_rate=0.0)
    m = FiDT5(conf).cuda()
    m.load_state_dict(sd)
    m.save_pretrained(str(tmp_path))
    m2 = FiDT5.from_pretrained(str(tmp_path)).cuda()
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    plain = {k.replace("encoder.encoder.", "encoder.").replace(".module.", "."): v for k, v in sd.items()}
    m3 = FiDT5(conf)
    m3.load_t5(plain)
    assert torch.equal(m3.state_dict()["encoder.encoder.block.1.module.layer.1.DenseReluDense.wi.weight"],
                       sd["encoder.encoder.block.1.module.layer.1.DenseReluDense.wi.weight"])


@pytest.mark.parametrize("path", GOLDEN[:2], ids=[os.path.basename(p) for p in GOLDEN[:2]])
def test_padded_and_packed_paths_agree(path):
    """The un-padded (packed) execution and the padded layout give the same loss / logits / gradients / tokens."""
    c, cfg, sd, model = _load(path)
    args = dict(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    res = {}
    from lako_b200 import _native as nat
    for unpad in (True, False):
        model.unpad = unpad
        model.train()
        model.zero_grad()
        nat.check(nat.lib().lako_debug_poison_tmem(nat.stream_ptr()))   # stale TMEM must never reach a result
        loss, logits = model(**args)
        loss.backward()
        torch.cuda.synchronize()
        model.eval()
        model.overwrite_forward_crossattention()
        model.reset_score_storage()
        gen = model.generate(c["ids"].cuda(), c["mask"].cuda(), max_length=c["max_length"])
        tok = model._tok_last.clone()
        res[unpad] = (loss.item(), logits.detach().clone(), model._flat_grad.clone(), gen.cpu(), tok)
    a, b = res[True], res[False]
    assert abs(a[0] - b[0]) < 2e-3
    assert (a[1] - b[1]).abs().max().item() < 2e-2
    assert ((a[2] - b[2]).norm() / b[2].norm()).item() < 0.05
    assert a[3].tolist() == b[3].tolist() == c["generated"].tolist()
    assert (a[4] - b[4]).abs().max().item() < 2e-2 * b[4].abs().max().item()


def test_all_padding_passage_and_oracle():
    """Edge case: one passage of a question is ENTIRELY padding (mask all zero).  The packed execution skips it (zero
    rows), the padded execution carries its fully masked rows (HF semantics: uniform attention over masked keys, then
    masked out as cross-attention keys); both must match the fp32 oracle in loss and logits, and each other in
    gradients."""
    from oracle import fidt5_oracle as O
    c, cfg, sd, model = _load(GOLDEN[0])
    ids, mask, labels = c["ids"].clone(), c["mask"].clone(), c["labels"]
    ids[0, 1] = 0
    mask[0, 1] = False
    with torch.no_grad():
        ref_loss, ref_logits = O.fidt5_forward(sd, cfg, ids, mask, labels)[:2]
    res = {}
    for unpad in (True, False):
        model.unpad = unpad
        model.train()          # dropout_rate is 0 in these golden configs
        model.zero_grad()
        loss, logits = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(model._flat_grad).all())
        res[unpad] = (loss.item(), logits.detach().float().cpu(), model._flat_grad.clone())
        assert abs(loss.item() - float(ref_loss)) < 5e-3, (unpad, loss.item(), float(ref_loss))
        assert (res[unpad][1] - ref_logits.reshape(res[unpad][1].shape)).abs().max().item() < 2e-2
    model.unpad = True
    assert ((res[True][2] - res[False][2]).norm() / res[False][2].norm()).item() < 0.05


@pytest.mark.parametrize("L", [300, 520])
def test_passages_longer_than_256_tokens_vs_oracle(L):
    """--text_maxlength is free in the reference (src/options.py:60).  Above 256 tokens the encoder attention runs the
    tiled kernels of csrc/enc_attn_long.cu: loss / logits / greedy tokens / token scores against the fp32 oracle, the
    packed and padded executions against each other in gradients, and (dropout on) a finite training step."""
    from oracle import fidt5_oracle as O
    c, cfg, sd, model = _load(GOLDEN[0])
    ids, mask, labels = O.synth_batch(2, 2, L, 6, cfg.vocab, seed=L)
    with torch.no_grad():
        ref_loss, ref_logits = O.fidt5_forward(sd, cfg, ids, mask, labels)[:2]
    res = {}
    for unpad in (True, False):
        model.unpad = unpad
        model.train()          # dropout_rate is 0 in this config
        model.zero_grad()
        loss, logits = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(model._flat_grad).all())
        res[unpad] = (logits.detach().float().cpu(), model._flat_grad.clone())
        assert abs(loss.item() - float(ref_loss)) < 5e-3, (unpad, loss.item(), float(ref_loss))
        assert (res[unpad][0] - ref_logits.reshape(res[unpad][0].shape)).abs().max().item() < LOGIT_ATOL_BF16
    model.unpad = True
    assert ((res[True][1] - res[False][1]).norm() / res[False][1].norm()).item() < 0.05
    model.set_checkpoint(True)     # encoder layers re-run in backward: through the tiled attention too
    model.zero_grad()
    loss, _ = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    model.set_checkpoint(False)
    # not 1e-5: the split-K dgrads sum fp32 partials with atomics and round the result to bf16, so a different summation
    # order can flip one element of the gradient signal by a bf16 ulp (seen as two identical steps 1.4e-3 apart about
    # once in twelve runs while the cross-attention dq was an atomic sum too, tools/race_probe.py)
    assert ((model._flat_grad - res[True][1]).norm() / res[True][1].norm()).item() < 5e-3
    model.eval()
    model.overwrite_forward_crossattention()
    model.reset_score_storage()
    out = model.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=8)
    ref_out, storage = O.fidt5_generate(sd, cfg, ids, mask, 8, capture_scores=True)
    assert out.cpu().tolist() == ref_out.tolist()
    tok, _ = model.token_scores(last_half=False)
    ref_tok, _ = O.token_scores_from_storage(storage, mask, last_half=False)
    B, N, _ = ids.shape
    assert (tok.cpu().view(B, N, L) - ref_tok.view(B, N, L)).abs().max().item() < SCORE_RTOL * ref_tok.abs().max().item()


@pytest.mark.parametrize("unpad", [True, False])
def test_use_checkpoint_recomputes_encoder_activations(unpad):
    """model.py:84-90 / CheckpointWrapper (src/model.py:236-284): with set_checkpoint(True) the encoder layers are re-run in
    backward from their saved inputs.  Same loss, same gradients (dropout ON: the recomputed masks must be the forward's),
    and a smaller activation workspace."""
    import ctypes
    from lako_b200 import _native as nat
    c, cfg, sd, model = _load(GOLDEN[0])
    model.config.dropout_rate = 0.1
    model.unpad = unpad
    model.train()
    ids, mask, labels = c["ids"].cuda(), c["mask"].cuda(), c["labels"].cuda()
    res = {False: [], True: []}
    for rep in range(3):
        for ckpt in (False, True):
            model.set_checkpoint(ckpt)
            model._drop_seed, model._drop_step = 1234, 0
            model.zero_grad()
            loss, logits = model(input_ids=ids, attention_mask=mask, labels=labels)
            loss.backward()
            torch.cuda.synchronize()
            res[ckpt].append((loss.item(), logits.detach().clone(), model._flat_grad.clone()))
    model.set_checkpoint(False)
    # logits bit-identical; the loss is an atomic fp32 sum over target rows (last-bit noise)
    assert all(abs(r[0] - res[False][0][0]) < 1e-5 and torch.equal(r[1], res[False][0][1]) for r in res[False] + res[True])
    assert all(bool(torch.isfinite(r[2]).all()) for r in res[True]) and res[False][0][2].norm().item() > 0
    # Gradients: the backward sums fp32 partials with atomics (split-K dgrads / wgrads); the order changes the last bits
    # and, where such a sum is then rounded to bf16, can flip one element by a bf16 ulp (~1e-3 of the gradient norm,
    # tools/race_probe.py).  So: some pair of (normal, recompute) runs agrees to summation noise, and no pair differs
    # by more than such a flip.
    dist = [((a[2] - b[2]).norm() / b[2].norm()).item() for a in res[True] for b in res[False]]
    assert min(dist) < 1e-5 and max(dist) < 5e-3, dist
    B, N, L = c["ids"].shape
    T = labels.size(1)
    lib = nat.lib()
    lib.lako_fidt5_workspace_bytes.restype = ctypes.c_size_t
    full = lib.lako_fidt5_workspace_bytes(ctypes.byref(model._dims), B, N, L, T, 1)
    small = lib.lako_fidt5_workspace_bytes(ctypes.byref(model._dims), B, N, L, T, 3)
    assert small < full


def test_host_passage_lengths_skip_the_readback():
    """Passing the Collator's host-side prefix lengths gives the same step as letting forward() read them back."""
    c, cfg, sd, model = _load(GOLDEN[0])
    mask = c["mask"]
    if not bool((mask == (torch.arange(mask.size(-1)) < mask.sum(-1, keepdim=True))).all()):
        pytest.skip("golden case is not a prefix mask")
    args = dict(input_ids=c["ids"].cuda(), attention_mask=mask.cuda(), labels=c["labels"].cuda())
    res = []
    for lengths in (None, mask.sum(-1, dtype=torch.int32)):
        model.train()
        model.zero_grad()
        loss, logits = model(passage_lengths=lengths, **args)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.item(), logits.detach().clone(), model._flat_grad.clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-5
    assert (res[0][1] - res[1][1]).abs().max().item() < 1e-5
    assert ((res[0][2] - res[1][2]).norm() / res[0][2].norm()).item() < 1e-4   # fp32 atomics reorder only


def test_non_prefix_mask_falls_back_to_padded_layout():
    """A mask with holes cannot be packed; the padded kernels must reproduce the oracle."""
    from oracle import fidt5_oracle as O
    c, cfg, sd, model = _load(GOLDEN[1])
    mask = c["mask"].clone()
    mask[:, :, 3] = False          # a hole inside every passage
    mask[0, 1, 10:20] = False
    cu, total = model._packing(mask.cuda())
    assert cu is None and total == 0
    model.eval()
    with torch.no_grad():
        loss, logits = model(input_ids=c["ids"].cuda(), attention_mask=mask.cuda(), labels=c["labels"].cuda())
        ref_loss, ref_logits = O.fidt5_forward(sd, cfg, c["ids"], mask, c["labels"])
    assert (logits.cpu() - ref_logits).abs().max().item() < LOGIT_ATOL_BF16
    assert abs(loss.item() - ref_loss.item()) < 5e-3


def test_grad_stage_events_partition_and_fire():
    """Stage events used to overlap the DP all-reduce: ranges tile the flat gradient, every event completes, and
    enabling them does not change the gradients."""
    c, cfg, sd, model = _load(GOLDEN[1])
    args = dict(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    model.train()
    model.zero_grad()
    model(**args)[0].backward()
    torch.cuda.synchronize()
    g0 = model._flat_grad.clone()
    evs, ranges = model.enable_grad_stage_events()
    live = sorted(r for r in ranges if r[1] > r[0])
    assert live[0][0] == 0 and live[-1][1] == model._total
    assert all(a[1] == b[0] for a, b in zip(live, live[1:]))
    model.zero_grad()
    model(**args)[0].backward()
    torch.cuda.synchronize()
    assert all(e.query() for e in evs)
    assert ((model._flat_grad - g0).norm() / g0.norm()).item() < 1e-3   # atomics reorder sums, nothing else changes


def test_transformers_302_cross_attention_bias_shifts_the_stored_scores():
    """With the 3.0.2 cross-attention bias table present, every stored step-0 score carries + table[0, h]
    (src/model.py:323-329 with compute_bias(qlen=1, klen)): the generated ids are unchanged (softmax invariance) and the
    token scores move by sum_h table[0, h] per selected layer on the valid tokens only."""
    import warnings
    from oracle import fidt5_oracle as O
    from lako_b200.model import FiDT5, FiDT5Config
    cfg = O.T5_TINY
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=4, emb_std=0.3), seed=5)
    ids, mask, _ = O.synth_batch(3, 2, 40, 4, cfg.vocab, seed=6, fact_format=True)
    table = torch.randn(cfg.n_buckets, cfg.n_heads)
    res = []
    for with_bias in (False, True):
        m = FiDT5(FiDT5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.n_layers,
                              num_heads=cfg.n_heads, dropout_rate=0.0))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m.load_state_dict({**sd, **({FiDT5.XATTN_BIAS_KEY: table} if with_bias else {})})
        m = m.cuda().eval()
        m.overwrite_forward_crossattention()
        m.reset_score_storage()
        out = m.generate(ids.cuda(), mask.cuda(), max_length=6)
        res.append((out.cpu(), m.token_scores(last_half=False)[0].cpu(), m.token_scores(last_half=True)[0].cpu()))
    assert torch.equal(res[0][0], res[1][0])
    c = float(table[0].sum())
    valid = mask.reshape(3, -1).float()
    assert torch.allclose(res[1][1] - res[0][1], valid * c * cfg.n_layers, atol=1e-4)
    assert torch.allclose(res[1][2] - res[0][2], valid * c * (cfg.n_layers - (cfg.n_layers + 1) // 2), atol=1e-4)
