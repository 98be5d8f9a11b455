"""HBM-bound kernels vs plain torch fp32 on a B200: RMSNorm(+residual) fwd/bwd, embedding, CE, AdamW."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _nat():
    from lako_b200 import _native as nat
    return nat


@pytest.mark.parametrize("rows,d", [(1000, 768), (257, 1024), (33, 128)])
def test_rmsnorm_fwd_bwd(rows, d):
    nat = _nat()
    L = nat.lib()
    torch.manual_seed(0)
    x = torch.randn(rows, d, device="cuda") * 3
    delta = (torch.randn(rows, d, device="cuda")).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(d, device="cuda")
    x_in = x.clone()
    out = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    rstd = torch.empty(rows, device="cuda")
    nat.check(L.lako_rmsnorm_fwd(nat.ptr(x), nat.ptr(delta), nat.ptr(x), nat.ptr(w), nat.ptr(out), nat.ptr(rstd), rows, d,
                                 ctypes.c_float(1e-6), ctypes.c_float(1.0), nat.stream_ptr()))
    xr = (x_in + delta.float()).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    var = xr.pow(2).mean(-1, keepdim=True)
    ref = wr * (xr * torch.rsqrt(var + 1e-6))
    assert torch.allclose(x, xr.detach(), atol=1e-6)
    assert torch.allclose(out.float(), ref.detach(), atol=2e-2, rtol=1e-2)
    assert torch.allclose(rstd, torch.rsqrt(var + 1e-6).squeeze(-1).detach(), rtol=1e-5)
    dy = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    ref.backward(dy.float())
    dres = torch.randn(rows, d, device="cuda")
    dres0 = dres.clone()
    dres_b = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(d, device="cuda")
    nat.check(L.lako_rmsnorm_bwd(nat.ptr(dy), nat.ptr(x), nat.ptr(rstd), nat.ptr(w), nat.ptr(dres), nat.ptr(dres_b),
                                 nat.ptr(dw), rows, d, 1, ctypes.c_float(1.0), nat.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.allclose(dres, dres0 + xr.grad, atol=1e-4, rtol=1e-4)
    assert torch.allclose(dres_b.float(), dres, atol=3e-2, rtol=1e-2)
    assert torch.allclose(dw, wr.grad, atol=1e-2, rtol=1e-3)


def test_embed_and_shift():
    nat = _nat()
    L = nat.lib()
    V, d, rows = 512, 128, 300
    table = torch.randn(V, d, device="cuda")
    ids = torch.randint(0, V, (rows,), device="cuda")
    out = torch.empty(rows, d, device="cuda")
    nat.check(L.lako_embed_gather(nat.ptr(ids), nat.ptr(table), nat.ptr(out), rows, d, V, nat.stream_ptr()))
    assert torch.equal(out, table[ids])
    dres = torch.randn(rows, d, device="cuda")
    dt = torch.zeros(V, d, device="cuda")
    nat.check(L.lako_embed_scatter_add(nat.ptr(ids), nat.ptr(dres), nat.ptr(dt), rows, d, V, nat.stream_ptr()))
    ref = torch.zeros(V, d, device="cuda").index_add_(0, ids, dres)
    assert torch.allclose(dt, ref, atol=1e-5)
    labels = torch.tensor([[5, 6, 1, -100], [7, 1, -100, -100]], device="cuda")
    dec = torch.empty_like(labels)
    nat.check(L.lako_shift_right(nat.ptr(labels), nat.ptr(dec), 2, 4, nat.stream_ptr()))
    assert dec.tolist() == [[0, 5, 6, 1], [0, 7, 1, 0]]


@pytest.mark.parametrize("V", [32128, 512, 1000])
def test_cross_entropy(V):
    nat = _nat()
    L = nat.lib()
    rows = 24
    torch.manual_seed(1)
    logits = (torch.randn(rows, V, device="cuda") * 3)
    labels = torch.randint(0, V, (rows,), device="cuda")
    labels[3] = -100
    labels[10] = -100
    stats = torch.zeros(2, device="cuda")
    row_loss = torch.empty(rows, device="cuda")
    ldd = (V + 7) // 8 * 8
    dlog = torch.zeros(rows, ldd, device="cuda", dtype=torch.bfloat16)
    nat.check(L.lako_count_valid(nat.ptr(labels), rows, nat.ptr(stats), nat.stream_ptr()))
    nat.check(L.lako_ce_fwd_bwd(nat.ptr(logits), nat.ptr(labels), nat.ptr(stats), nat.ptr(row_loss), nat.ptr(dlog), rows,
                                V, V, ldd, nat.stream_ptr()))
    torch.cuda.synchronize()
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-100)
    ref.backward()
    assert stats[1].item() == rows - 2
    assert abs(stats[0].item() / stats[1].item() - ref.item()) < 1e-4
    assert torch.allclose(dlog[:, :V].float(), lr.grad, atol=2e-4, rtol=1e-2)


def test_adamw_clip_and_cast():
    nat = _nat()
    L = nat.lib()
    from oracle import fidt5_oracle as O
    n = 4096 * 3 + 8
    torch.manual_seed(2)
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.1
    p_ref = p.clone().cpu()
    g_ref = g.clone().cpu()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    wb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(1, device="cuda")
    opt = O.HFAdamW([p_ref], lr=1e-3, weight_decay=1e-2)
    for step in range(3):
        ss.zero_()
        nat.check(L.lako_grad_sumsq(nat.ptr(g), ctypes.c_size_t(n), nat.ptr(ss), nat.stream_ptr()))
        nat.check(L.lako_adamw_step(nat.ptr(p), nat.ptr(m), nat.ptr(v), nat.ptr(g), nat.ptr(wb), ctypes.c_size_t(n),
                                    nat.ptr(ss), ctypes.c_float(1.0), ctypes.c_float(1.0), ctypes.c_float(1e-3),
                                    ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-6),
                                    ctypes.c_float(1e-2), nat.stream_ptr()))
        gr = [g_ref.clone()]
        O.clip_grad_norm_(gr, 1.0)
        opt.step(gr)
    torch.cuda.synchronize()
    assert torch.allclose(p.cpu(), p_ref, atol=1e-6, rtol=1e-5)
    assert torch.allclose(wb.float().cpu(), p_ref, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("style", ["mean", "max", "21mean"])
@pytest.mark.parametrize("N,L,n_context", [(2, 64, 4), (2, 130, 10), (20, 250, 10), (3, 40, 50), (2, 7, 3)])
def test_fact_scores_kernel_bit_exact_vs_oracle_host_loop(style, N, L, n_context):
    """lako_fact_scores vs the oracle's restatement of the reference's Python loop (src/model.py:166-203) on the SAME
    token scores: segment boundaries, counts, -5 fillers and the float64 arithmetic order are identical, so the
    results must be equal to the last bit."""
    import ctypes
    from oracle import fidt5_oracle as O
    from lako_b200 import _native as nat
    g = torch.Generator().manual_seed(N * 1000 + L + n_context)
    B = 9
    tok = torch.randn(B, N, L, generator=g)
    ids = torch.randint(6, 100, (B, N, L), generator=g)
    ids[:, 1, 0], ids[:, 1, 1] = 6, 10
    for b in range(B):
        kind = b % 5
        if kind == 0:      # regular facts, padded tail
            pos = torch.randperm(L - 3, generator=g)[: max(1, L // 9)] + 2
            ids[b, 1, pos] = 5
            ids[b, 1, int(pos.max()) + 1:] = 0
        elif kind == 1:    # unterminated last fact, no padding
            pos = torch.randperm(L - 3, generator=g)[: max(1, L // 12)] + 2
            ids[b, 1, pos] = 5
            ids[b, 1, -1] = 77
        elif kind == 2:    # no separator at all
            pass
        elif kind == 3:    # separators back to back (one-token facts) and one at the very end
            ids[b, 1, 2:min(L, 8)] = 5
            ids[b, 1, -1] = 5
        else:              # all padding after the prefix
            ids[b, 1, 2:] = 0
    half = (N + 1) // 2
    ref = O.fact_scores_from_token_scores(tok, ids, n_context, style, 24.0)
    out = torch.empty(B, n_context, dtype=torch.float64, device="cuda")
    tok_d, ids1 = tok.cuda().contiguous(), ids[:, 1].cuda().contiguous()
    nat.check(nat.lib().lako_fact_scores(nat.ptr(tok_d), None, nat.ptr(ids1), ctypes.c_int(L), nat.ptr(out), ctypes.c_int(B),
                                         ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(half), ctypes.c_int(n_context),
                                         ctypes.c_int({"mean": 0, "max": 1, "21mean": 2}[style]), ctypes.c_double(24.0),
                                         ctypes.c_int(0), nat.stream_ptr()))
    assert torch.equal(out.cpu(), ref), (out.cpu() - ref).abs().max()
    sm = torch.empty_like(out)
    nat.check(nat.lib().lako_fact_scores(nat.ptr(tok_d), None, nat.ptr(ids1), ctypes.c_int(L), nat.ptr(sm), ctypes.c_int(B),
                                         ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(half), ctypes.c_int(n_context),
                                         ctypes.c_int({"mean": 0, "max": 1, "21mean": 2}[style]), ctypes.c_double(24.0),
                                         ctypes.c_int(1), nat.stream_ptr()))
    assert torch.allclose(sm.cpu(), torch.softmax(ref, dim=-1), atol=1e-14, rtol=1e-12)
