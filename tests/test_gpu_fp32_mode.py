"""precision='fp32' (B200): the evaluation / score-extraction path with 3xTF32-split tcgen05 GEMMs and fp32 attention.
north_star's tolerances for this mode: logits within 1e-4 absolute of the reference, cross-attention scores within
1e-2 relative, greedy answers token-identical.  The reference itself computes in plain fp32 (train_reader.py:67)."""
import ctypes
import glob
import os
import types

import pytest
import torch

from lako_b200.model import BUCKET_CENTER as CENTER

pytestmark = pytest.mark.gpu
c_int, c_float = ctypes.c_int, ctypes.c_float
NEG = torch.finfo(torch.float32).min
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def _nat():
    from lako_b200 import _native as nat
    return nat


def _split(nat, x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    nat.check(nat.lib().lako_split_tf32(nat.ptr(x), nat.ptr(hi), nat.ptr(lo), ctypes.c_size_t(x.numel()), nat.stream_ptr()))
    return hi, lo


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (200, 300, 100), (1000, 768, 768), (80, 3072, 1024), (8, 32128, 512),
                                   (4096, 256, 4096)])
@pytest.mark.parametrize("relu", [0, 1])
def test_gemm_tf32x3_is_fp32_accurate(M, N, K, relu):
    nat = _nat()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") * K ** -0.5
    Ah, Al = _split(nat, A)
    Bh, Bl = _split(nat, B)
    assert torch.equal(Ah.view(torch.int32) & 0x1fff, torch.zeros_like(Ah, dtype=torch.int32))
    assert (A - Ah - Al).abs().max().item() <= A.abs().max().item() * 2.0 ** -20
    C = torch.zeros(M, N, device="cuda")
    Ch, Cl = torch.zeros_like(C), torch.zeros_like(C)
    nat.check(nat.lib().lako_gemm_tf32x3(nat.ptr(Ah), nat.ptr(Al), nat.ptr(Bh), nat.ptr(Bl), nat.ptr(C), nat.ptr(Ch), nat.ptr(Cl),
                                         c_int(M), c_int(N), c_int(K), c_int(K), c_int(K), c_int(N), c_int(relu), c_float(0.5),
                                         nat.stream_ptr()))
    torch.cuda.synchronize()
    ref = 0.5 * (A.double() @ B.double().T)
    if relu:
        ref = torch.relu(ref)
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"tf32x3 {M}x{N}x{K}: max err {err:.3e} (max |C| {scale:.2f})")
    # the tensor core's fp32 accumulation truncates instead of rounding to nearest, so the error grows ~K * 2^-24
    # (measured 5.8e-6 * |C| at K = 768, 3.2e-5 * |C| at K = 4096); a single-pass kind::tf32 product is at 1e-3
    assert err < (3e-6 + 1.2e-8 * K) * max(1.0, scale)
    assert (C - Ch - Cl).abs().max().item() <= scale * 2.0 ** -20


def test_fp32_attention_kernels_vs_torch():
    nat = _nat()
    lib = nat.lib()
    from lako_b200.model import bucket_table
    torch.manual_seed(5)
    P, L, H = 3, 150, 2
    inner = H * 64
    lens = torch.tensor([150, 97, 64])
    cu = torch.zeros(P + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    rows = int(cu[-1])
    qkv = torch.randn(rows, 3 * inner, device="cuda") * 0.7
    table = torch.randn(32, H, device="cuda")
    buckets = bucket_table(True, 32, 128).cuda()
    lut = torch.empty(H, 2 * L - 1, device="cuda")
    nat.check(lib.lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, CENTER, nat.stream_ptr()))
    out = torch.zeros(rows, inner, device="cuda")
    hi, lo = torch.zeros_like(out), torch.zeros_like(out)
    cu_d = cu.cuda()
    nat.check(lib.lako_enc_attn_f32(nat.ptr(qkv), None, nat.ptr(lut), nat.ptr(out), nat.ptr(hi), nat.ptr(lo), P, L, H,
                                    nat.ptr(cu_d), nat.stream_ptr()))
    torch.cuda.synchronize()
    for p_ in range(P):
        r0, n = int(cu[p_]), int(lens[p_])
        x = qkv[r0:r0 + n].view(n, 3, H, 64)
        q, k, v = x[:, 0].transpose(0, 1), x[:, 1].transpose(0, 1), x[:, 2].transpose(0, 1)
        rel = torch.arange(n, device="cuda")[None, :] - torch.arange(n, device="cuda")[:, None]
        bias = table[buckets[rel + CENTER].long()].permute(2, 0, 1)
        ref = (torch.softmax(q @ k.transpose(-1, -2) + bias, -1) @ v).transpose(0, 1).reshape(n, inner)
        assert (out[r0:r0 + n] - ref).abs().max().item() < 2e-5
    assert (out - hi - lo).abs().max().item() < 1e-6
    # padded layout with a key mask
    qkv2 = torch.randn(P * L, 3 * inner, device="cuda") * 0.7
    mask = (torch.arange(L)[None] < lens[:, None]).cuda()
    out2 = torch.zeros(P * L, inner, device="cuda")
    nat.check(lib.lako_enc_attn_f32(nat.ptr(qkv2), nat.ptr(mask), nat.ptr(lut), nat.ptr(out2), None, None, P, L, H, None,
                                    nat.stream_ptr()))
    x = qkv2.view(P, L, 3, H, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    rel = torch.arange(L, device="cuda")[None, :] - torch.arange(L, device="cuda")[:, None]
    bias = table[buckets[rel + CENTER].long()].permute(2, 0, 1)[None]
    add = torch.zeros(P, 1, 1, L, device="cuda").masked_fill(~mask[:, None, None, :], NEG)
    ref = (torch.softmax(q @ k.transpose(-1, -2) + bias + add, -1) @ v).transpose(1, 2).reshape(P * L, inner)
    assert (out2 - ref).abs().max().item() < 2e-5
    # RMSNorm + residual, in place
    xr = torch.randn(37, 768, device="cuda") * 2
    dl = torch.randn(37, 768, device="cuda")
    g = 1 + 0.1 * torch.randn(768, device="cuda")
    x_in = xr.clone()
    nh, nl_ = torch.zeros_like(xr), torch.zeros_like(xr)
    nat.check(lib.lako_rmsnorm_f32(nat.ptr(x_in), nat.ptr(dl), nat.ptr(x_in), nat.ptr(g), None, nat.ptr(nh), nat.ptr(nl_), 37, 768,
                                   c_float(1e-6), nat.stream_ptr()))
    y = xr + dl
    refn = g * (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert torch.allclose(x_in, y) and (nh + nl_ - refn).abs().max().item() < 1e-5


def _mk(cfg, sd):
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.n_layers,
                          num_heads=cfg.n_heads, dropout_rate=0.0))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.set_precision("fp32")
    return m


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_fp32_mode_vs_reference_golden(path):
    """Against the outputs of the reference itself: logits 1e-4, greedy identical, scores 1e-2 (all six variants)."""
    from oracle import fidt5_oracle as O
    c = torch.load(path, weights_only=False)
    cfg = O.T5Dims(**c["cfg"])
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=c["seed"], emb_std=c["emb_std"]), seed=c["seed"] + 1)
    if c.get("eos_alias") is not None:
        sd["shared.weight"][1] = sd["shared.weight"][c["eos_alias"]] * 1.05
    model = _mk(cfg, sd)
    with torch.no_grad():
        loss, logits = model(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), labels=c["labels"].cuda())
    err = (logits.cpu() - c["logits"]).abs().max().item()
    print(f"fp32 mode max |dlogit| {err:.2e}; loss {loss.item():.6f} vs {c['loss'].item():.6f}")
    assert err < 1e-4
    assert abs(loss.item() - c["loss"].item()) < 1e-4
    model.overwrite_forward_crossattention()
    model.reset_score_storage()
    out = model.generate(c["ids"].cuda(), c["mask"].cuda(), max_length=c["max_length"])
    assert out.cpu().tolist() == c["generated"].tolist()
    if c["N"] % 2 == 0:
        for style in ("mean", "max", "21mean"):
            for lh in ("yes", "no"):
                opt = types.SimpleNamespace(stream=2, n_context=c["n_context"], attention_score_style=style,
                                            use_last_half_layer_attention=lh)
                sc = model.get_crossattention_scores(opt, c["ids"], None, c["mask"].cuda())
                ref = c[f"scores_{style}_{lh}"]
                assert torch.allclose(sc, ref, rtol=1e-2, atol=1e-5), (style, lh, (sc - ref).abs().max().item())
        tok, _ = model.token_scores(last_half=True)
        ref_tok = c["tok_scores_last_half"]
        rel = (tok.cpu().view_as(ref_tok) - ref_tok).abs().max().item() / ref_tok.abs().max().item()
        print(f"fp32 mode token scores: max|err|/max|score| {rel:.2e}")
        assert rel < 1e-3


@pytest.mark.parametrize("size,B,N,L", [("base", 2, 2, 64), ("base", 1, 10, 200), ("large", 1, 20, 250)])
def test_fp32_mode_real_shapes_vs_oracle(size, B, N, L):
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import fidt5_oracle as O
    cfg = {"base": O.T5_BASE, "large": O.T5_LARGE}[size]
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=70, emb_std=0.25), seed=71)
    model = _mk(cfg, sd)
    sd_gpu = O.sd_to(sd, "cuda")
    ids, mask, labels = [t.cuda() for t in O.synth_batch(B, N, L, 10, cfg.vocab, seed=72, fact_format=True)]
    with torch.no_grad():
        loss, logits = model(input_ids=ids, attention_mask=mask, labels=labels)
        ref_loss, ref_logits = O.fidt5_forward(sd_gpu, cfg, ids, mask, labels)
    err = (logits - ref_logits).abs().max().item()
    print(f"fp32 mode t5-{size} N={N} L={L}: max |dlogit| {err:.2e} (absmax {ref_logits.abs().max().item():.2f}); "
          f"loss {loss.item():.6f} vs {ref_loss.item():.6f}")
    assert err < 1e-4
    model.overwrite_forward_crossattention()
    model.reset_score_storage()
    out = model.generate(ids, mask, max_length=10, min_length=10)
    ref_out, storage = O.fidt5_generate(sd_gpu, cfg, ids, mask, 10, capture_scores=True, min_length=10)
    assert out.cpu().tolist() == ref_out.cpu().tolist()
    for lh in (True, False):
        tok, div = model.token_scores(last_half=lh)
        ref_tok, ref_div = O.token_scores_from_storage(storage, mask, last_half=lh)
        valid = mask.reshape(ref_tok.shape)
        d = (tok.view_as(ref_tok) - ref_tok)[valid]
        rel = d.abs().max().item() / ref_tok[valid].abs().max().item()
        print(f"fp32 mode token scores last_half={lh}: max|err|/max|score| {rel:.2e}")
        assert rel < 1e-2
    opt = types.SimpleNamespace(stream=2, n_context=6, attention_score_style="21mean", use_last_half_layer_attention="yes")
    sc = model.get_crossattention_scores(opt, ids, None, mask)
    ref = O.get_crossattention_scores(storage, ids.cpu(), mask, 6, "21mean", True)
    assert torch.allclose(sc, ref, rtol=1e-2, atol=1e-6)


def test_fp32_mode_refuses_training():
    from oracle import fidt5_oracle as O
    from lako_b200 import _native as nat
    cfg = O.T5_TINY
    model = _mk(cfg, O.init_state_dict(cfg, seed=1, emb_std=0.3))
    ids, mask, labels = [t.cuda() for t in O.synth_batch(2, 2, 32, 4, cfg.vocab, seed=2)]
    with pytest.raises(nat.LakoError):
        model(input_ids=ids, attention_mask=mask, labels=labels)   # grad enabled -> training forward
