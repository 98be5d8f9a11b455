"""Attention kernels vs torch fp32 references on the same bf16-rounded inputs (B200)."""
import ctypes

import pytest
import torch

from lako_b200.model import BUCKET_CENTER as CENTER

pytestmark = pytest.mark.gpu
NEG = torch.finfo(torch.float32).min


def _nat():
    from lako_b200 import _native as nat
    return nat


def _bias_lut(nat, table, buckets, H, L, center=CENTER):
    lut = torch.empty(H, 2 * L - 1, device="cuda")
    nat.check(nat.lib().lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, center,
                                            nat.stream_ptr()))
    return lut


@pytest.mark.parametrize("impl", [2, 1])
@pytest.mark.parametrize("P,L,H", [(3, 64, 2), (4, 200, 3), (2, 250, 2), (2, 256, 1), (5, 40, 2), (2, 130, 4), (40, 250, 16)])
def test_enc_attn_fwd_bwd(P, L, H, impl):
    """impl 2 = persistent warp-specialised forward with P in tensor memory (default), 1 = round 1's one-shot kernel."""
    _enc_attn_fwd_bwd(P, L, H, impl)


@pytest.mark.parametrize("P,L,H,impl", [(2, 300, 2, 2), (2, 512, 3, 2), (1, 700, 1, 2), (2, 1100, 1, 2), (3, 257, 2, 2),
                                        (3, 200, 2, 3), (2, 64, 1, 3), (2, 130, 4, 3), (5, 40, 2, 3)])
def test_enc_attn_fwd_bwd_any_length(P, L, H, impl):
    """The reference takes any --text_maxlength (src/options.py:60): passages longer than 256 tokens go through the tiled
    kernels of csrc/enc_attn_long.cu (impl 3 forces that path at lengths the tensor-memory kernels also serve)."""
    _enc_attn_fwd_bwd(P, L, H, impl)


def _enc_attn_fwd_bwd(P, L, H, impl):
    nat = _nat()
    lib = nat.lib()
    lib.lako_set_attn_impl(impl)
    from lako_b200.model import bucket_table
    torch.manual_seed(L)
    inner = H * 64
    qkv = (torch.randn(P * L, 3 * inner, device="cuda") * 0.7).to(torch.bfloat16)
    lens = torch.randint(int(0.5 * L), L + 1, (P,), device="cuda")
    lens[0] = L
    mask = (torch.arange(L, device="cuda")[None] < lens[:, None])
    table = torch.randn(32, H, device="cuda")
    buckets = bucket_table(True, 32, 128).cuda()
    lut = _bias_lut(nat, table, buckets, H, L)
    out = torch.zeros(P * L, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(P, H, L, device="cuda")
    # every TMEM column of every SM holds NaN bits before the kernels run (tcgen05.alloc does not clear): a kernel that
    # reads a column it did not write -- e.g. the tail of the last 32-key chunk when L % 32 != 0 -- fails this test
    nat.check(lib.lako_debug_poison_tmem(nat.stream_ptr()))
    nat.check(lib.lako_enc_attn_fwd(nat.ptr(qkv), nat.ptr(mask), nat.ptr(lut), nat.ptr(out), nat.ptr(lse), P, L, H,
                                    nat.stream_ptr()))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all()) and bool(torch.isfinite(lse).all())
    # reference
    x = qkv.float().view(P, L, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    tab = table.clone().requires_grad_(True)
    rel = torch.arange(L, device="cuda")[None, :] - torch.arange(L, device="cuda")[:, None]
    bias = tab[buckets[rel + CENTER].long()].permute(2, 0, 1)[None]
    add = torch.zeros(P, 1, 1, L, device="cuda").masked_fill(~mask[:, None, None, :], NEG)
    s = q @ k.transpose(-1, -2) + (bias + add)
    pr = torch.softmax(s, dim=-1)
    ref = (pr @ v).transpose(1, 2).reshape(P * L, inner)
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=3e-2), (out.float() - ref).abs().max().item()
    assert torch.allclose(lse, torch.logsumexp(s, dim=-1), atol=2e-3, rtol=1e-4)
    lib.lako_set_attn_impl(3 if impl == 3 else 2)
    # backward
    dout = (torch.randn(P * L, inner, device="cuda") * 0.5).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dlut = torch.zeros(H, 2 * L - 1, device="cuda")
    nat.check(lib.lako_debug_poison_tmem(nat.stream_ptr()))
    nat.check(lib.lako_enc_attn_bwd(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(mask), nat.ptr(lut),
                                    nat.ptr(dqkv), nat.ptr(dlut), P, L, H, nat.stream_ptr()))
    dtab = torch.zeros(32, H, device="cuda")
    nat.check(lib.lako_fold_dlut(nat.ptr(dlut), nat.ptr(buckets), nat.ptr(dtab), H, L, CENTER, 32, nat.stream_ptr()))
    torch.cuda.synchronize()
    gref = x.grad.reshape(P * L, 3 * inner)
    scale = gref.abs().max().item()
    err = (dqkv.float() - gref).abs().max().item()
    assert err < 0.03 * scale + 1e-3, (err, scale)
    terr = (dtab - tab.grad).abs().max().item()
    lib.lako_set_attn_impl(2)
    assert terr < 0.02 * tab.grad.abs().max().item() + 1e-2, (terr, tab.grad.abs().max().item())


@pytest.mark.parametrize("P,L,H,with_mask,impl", [(3, 250, 2, False, 2), (2, 100, 3, True, 2), (2, 256, 1, False, 2),
                                                  (3, 250, 2, False, 1), (2, 100, 3, True, 1), (2, 256, 1, False, 1),
                                                  (2, 100, 3, True, 3), (2, 250, 2, False, 3), (2, 300, 2, True, 2),
                                                  (1, 520, 1, False, 2)])
def test_enc_attn_dropout_matches_host_mask(P, L, H, with_mask, impl):
    """Dropout on the attention probabilities (HF:332): the keep words the forward saves equal the host restatement
    (lako_b200/philox.py attn_keep_mask) bit for bit, the output equals softmax * mask / (1 - p) @ v in torch, and the
    backward -- reading the saved words, and regenerating them -- gives that graph's gradients."""
    import ctypes as C
    import numpy as np
    nat = _nat()
    lib = nat.lib()
    lib.lako_set_attn_impl(impl)
    from lako_b200.model import bucket_table
    from lako_b200.philox import attn_keep_mask, scale_of
    torch.manual_seed(7 * L + H)
    inner = H * 64
    seed, step, site, pdrop = 0x1234567890ABCDEF, 5, 77, 0.1
    qkv = (torch.randn(P * L, 3 * inner, device="cuda") * 0.6).to(torch.bfloat16)
    lens = torch.randint(int(0.6 * L), L + 1, (P,), device="cuda")
    lens[0] = L
    mask = (torch.arange(L, device="cuda")[None] < lens[:, None]) if with_mask else None
    table = torch.randn(32, H, device="cuda")
    buckets = bucket_table(True, 32, 128).cuda()
    lut = _bias_lut(nat, table, buckets, H, L)
    out = torch.zeros(P * L, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(P, H, L, device="cuda")
    keep = torch.zeros(P, H, L, 8, device="cuda", dtype=torch.int32)
    args = (C.c_ulonglong(seed), C.c_uint(step), C.c_uint(site), C.c_float(pdrop), nat.stream_ptr())
    nat.check(lib.lako_enc_attn_fwd_dropout(nat.ptr(qkv), nat.ptr(mask), nat.ptr(lut), nat.ptr(out), nat.ptr(lse),
                                            nat.ptr(keep), P, L, H, *args))
    torch.cuda.synchronize()
    tiled = impl == 3 or L > 256      # csrc/enc_attn_long.cu: saves no keep words, its backward regenerates them
    lib.lako_set_attn_impl(3 if tiled else 2)
    km = attn_keep_mask(seed, step, site, P, H, L, pdrop)                       # [P, H, L, 32 * max(8, ceil(L/32))] bool
    if not tiled:
        got = keep.cpu().numpy().view(np.uint32)
        j = np.arange(32)
        bit = ((j >> 1) + 16 * (j & 1)).astype(np.uint32)
        got_mask = ((got[..., None] >> bit) & 1).astype(bool).reshape(P, H, L, 256)
        nch = (L + 31) // 32                                                    # chunks the kernels touch
        assert (got_mask[..., :nch * 32] == km[..., :nch * 32]).all()
    assert abs(km[..., :L].mean() - (1 - pdrop)) < 5e-3
    keep_t = torch.from_numpy(km[..., :L]).cuda()
    x = qkv.float().view(P, L, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    tab = table.clone().requires_grad_(True)
    rel = torch.arange(L, device="cuda")[None, :] - torch.arange(L, device="cuda")[:, None]
    bias = tab[buckets[rel + CENTER].long()].permute(2, 0, 1)[None]
    s = q @ k.transpose(-1, -2) + bias
    if with_mask:
        s = s + torch.zeros(P, 1, 1, L, device="cuda").masked_fill(~mask[:, None, None, :], NEG)
    pr = torch.softmax(s, dim=-1) * keep_t * scale_of(pdrop)
    ref = (pr @ v).transpose(1, 2).reshape(P * L, inner)
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=3e-2), (out.float() - ref).abs().max().item()
    dout = (torch.randn(P * L, inner, device="cuda") * 0.5).to(torch.bfloat16)
    ref.backward(dout.float())
    gref = x.grad.reshape(P * L, 3 * inner)
    for saved in ((None,) if tiled else (keep, None)):
        dqkv = torch.zeros_like(qkv)
        dlut = torch.zeros(H, 2 * L - 1, device="cuda")
        nat.check(lib.lako_enc_attn_bwd_dropout(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(mask),
                                                nat.ptr(lut), nat.ptr(dqkv), nat.ptr(dlut), nat.ptr(saved), P, L, H, *args))
        torch.cuda.synchronize()
        scale = gref.abs().max().item()
        err = (dqkv.float() - gref).abs().max().item()
        assert err < 0.03 * scale + 1e-3, (err, scale, saved is None)
    lib.lako_set_attn_impl(2)


@pytest.mark.parametrize("B,T,M,H,nl", [(2, 6, 96, 2, 2), (3, 10, 500, 3, 1), (1, 1, 333, 2, 2), (2, 20, 2000, 12, 1), (2, 40, 300, 2, 1), (1, 64, 129, 1, 1)])
def test_xattn_fwd_bwd(B, T, M, H, nl):
    nat = _nat()
    lib = nat.lib()
    lib.lako_xattn_workspace_bytes.restype = ctypes.c_size_t
    torch.manual_seed(M)
    inner = H * 64
    ldkv = nl * 2 * inner
    layer = nl - 1
    koff, voff = layer * 2 * inner, layer * 2 * inner + inner
    q = (torch.randn(B * T, inner, device="cuda") * 0.5).to(torch.bfloat16)
    kv = (torch.randn(B * M, ldkv, device="cuda") * 0.7).to(torch.bfloat16)
    mask = torch.rand(B, M, device="cuda") > 0.2
    mask[:, 0] = True
    out = torch.zeros(B * T, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    tok = torch.zeros(B, M, device="cuda")
    ws = torch.empty(lib.lako_xattn_workspace_bytes(B, H, T, M), dtype=torch.uint8, device="cuda")
    nat.check(lib.lako_xattn_fwd(nat.ptr(q), nat.ptr(kv), nat.ptr(mask), nat.ptr(out), nat.ptr(lse), nat.ptr(tok),
                                 nat.ptr(ws), B, T, M, H, inner, ldkv, koff, voff, inner, nat.stream_ptr()))
    torch.cuda.synchronize()
    qf = q.float().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    kvf = kv.float().requires_grad_(True)
    kf = kvf[:, koff:koff + inner].view(B, M, H, 64).transpose(1, 2)
    vf = kvf[:, voff:voff + inner].view(B, M, H, 64).transpose(1, 2)
    raw = qf @ kf.transpose(-1, -2)
    s = raw + torch.zeros(B, 1, 1, M, device="cuda").masked_fill(~mask[:, None, None, :], NEG)
    pr = torch.softmax(s, dim=-1)
    ref = (pr @ vf).transpose(1, 2).reshape(B * T, inner)
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=2e-2), (out.float() - ref).abs().max().item()
    assert torch.allclose(lse, torch.logsumexp(s, dim=-1), atol=2e-3, rtol=1e-4)
    tok_ref = raw[:, :, 0, :].masked_fill(~mask[:, None, :], 0.0).sum(1)
    assert torch.allclose(tok, tok_ref.detach(), atol=1e-3, rtol=1e-4)
    dout = (torch.randn(B * T, inner, device="cuda") * 0.5).to(torch.bfloat16)
    ref.backward(dout.float())
    dkv = torch.zeros_like(kv)
    dq = torch.zeros(B * T, inner, device="cuda")
    nat.check(lib.lako_xattn_bwd(nat.ptr(q), nat.ptr(kv), nat.ptr(mask), nat.ptr(out), nat.ptr(dout), nat.ptr(lse),
                                 nat.ptr(dkv), nat.ptr(dq), B, T, M, H, inner, ldkv, koff, voff, inner,
                                 nat.stream_ptr()))
    torch.cuda.synchronize()
    gq = qf.grad.transpose(1, 2).reshape(B * T, inner)
    assert (dq - gq).abs().max().item() < 0.02 * gq.abs().max().item() + 1e-3
    gkv = kvf.grad[:, koff:koff + 2 * inner]
    err = (dkv[:, koff:koff + 2 * inner].float() - gkv).abs().max().item()
    assert err < 0.02 * gkv.abs().max().item() + 1e-3, err


@pytest.mark.parametrize("B,T,H", [(2, 6, 2), (3, 20, 3), (1, 1, 2), (2, 64, 2)])
def test_dself_fwd_bwd(B, T, H):
    nat = _nat()
    lib = nat.lib()
    from lako_b200.model import bucket_table
    torch.manual_seed(T)
    inner = H * 64
    qkv = (torch.randn(B * T, 3 * inner, device="cuda") * 0.6).to(torch.bfloat16)
    table = torch.randn(32, H, device="cuda")
    buckets = bucket_table(False, 32, 128).cuda()
    lut = _bias_lut(nat, table, buckets, H, T)
    out = torch.zeros(B * T, inner, device="cuda", dtype=torch.bfloat16)
    probs = torch.zeros(B, H, T, T, device="cuda")
    nat.check(lib.lako_dself_fwd(nat.ptr(qkv), nat.ptr(lut), nat.ptr(out), nat.ptr(probs), B, T, H, nat.stream_ptr()))
    x = qkv.float().view(B, T, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    tab = table.clone().requires_grad_(True)
    rel = torch.arange(T, device="cuda")[None, :] - torch.arange(T, device="cuda")[:, None]
    bias = tab[buckets[rel + CENTER].long()].permute(2, 0, 1)[None]
    causal = torch.zeros(T, T, device="cuda").masked_fill(rel > 0, NEG)
    s = q @ k.transpose(-1, -2) + (bias + causal)
    pr = torch.softmax(s, dim=-1)
    ref = (pr @ v).transpose(1, 2).reshape(B * T, inner)
    torch.cuda.synchronize()
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=2e-2)
    dout = (torch.randn(B * T, inner, device="cuda") * 0.5).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dlut = torch.zeros(H, 2 * T - 1, device="cuda")
    nat.check(lib.lako_dself_bwd(nat.ptr(qkv), nat.ptr(dout), nat.ptr(probs), nat.ptr(dqkv), nat.ptr(dlut), B, T, H,
                                 nat.stream_ptr()))
    dtab = torch.zeros(32, H, device="cuda")
    nat.check(lib.lako_fold_dlut(nat.ptr(dlut), nat.ptr(buckets), nat.ptr(dtab), H, T, CENTER, 32, nat.stream_ptr()))
    torch.cuda.synchronize()
    g = x.grad.reshape(B * T, 3 * inner)
    assert (dqkv.float() - g).abs().max().item() < 0.02 * g.abs().max().item() + 1e-3
    assert (dtab - tab.grad).abs().max().item() < 0.01 * tab.grad.abs().max().item() + 1e-3
