"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/lako_b200.h declares,
the flat layout / workspace sizing are consistent, host logic (bucket tables, state-dict names, score
post-processing, LR schedule) matches the oracle, and the product refuses to run without a B200."""
import ctypes
import os
import re
import types

import pytest
import torch

from oracle import fidt5_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lako_b200 import _native as nat
    lib = nat.lib()
    hdr = open(os.path.join(ROOT, "include", "lako_b200.h")).read()
    names = set(re.findall(r"\b(lako_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/lako_b200.h but not exported"


def test_layout_matches_parameter_count():
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2))
    n_params = sum(p.numel() for p in m.parameters())
    ref = sum(v.numel() for v in O.unique_params(O.init_state_dict(O.T5_TINY)).values())
    assert n_params == ref
    assert m._total >= n_params
    # every spec is inside the flat buffer and disjoint
    spans = sorted((off, off + int(torch.tensor(shape).prod())) for off, shape in m._specs.values())
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= m._total
    # fused q|k|v and fused cross k|v of all layers are contiguous slabs
    lo = m._layout
    inner, d = 128, 128
    assert m._specs["decoder.block.1.layer.1.EncDecAttention.k.weight"][0] == lo.dec_xkv + 2 * inner * d
    assert m._specs["encoder.encoder.block.0.module.layer.0.SelfAttention.v.weight"][0] == lo.enc_qkv[0] + 2 * inner * d


def test_t5_sizes_parameter_counts():
    from lako_b200 import _native as nat
    from lako_b200.model import _Dims, _Layout
    lib = nat.lib()
    for dims, expect in (((768, 12, 3072, 12, 32128, 32), 222903552), ((1024, 16, 4096, 24, 32128, 32), 737668096)):
        dm = _Dims(*dims, 1e-6)
        lo = _Layout()
        assert lib.lako_fidt5_layout(ctypes.byref(dm), ctypes.byref(lo)) == 0
        assert lo.total == expect  # unique parameter counts of t5-base / t5-large
        lib.lako_fidt5_workspace_bytes.restype = ctypes.c_size_t
        ws_train = lib.lako_fidt5_workspace_bytes(ctypes.byref(dm), 8, 10, 200, 10, 1)
        ws_eval = lib.lako_fidt5_workspace_bytes(ctypes.byref(dm), 8, 10, 200, 10, 0)
        assert ws_eval < ws_train < 170 * 2 ** 30


def test_state_dict_names_match_reference():
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2))
    assert set(m.state_dict().keys()) == set(O.init_state_dict(O.T5_TINY).keys())
    sd = m.state_dict()
    assert sd["lm_head.weight"].data_ptr() == sd["shared.weight"].data_ptr()
    # 3.0.2-era extra key is accepted and ignored (deviation D1)
    extra = dict(O.init_state_dict(O.T5_TINY))
    extra["decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight"] = torch.zeros(32, 2)
    m.load_state_dict(extra)
    with pytest.raises(RuntimeError):
        m.load_state_dict({"bogus": torch.zeros(1)})


def test_bucket_tables_match_oracle():
    from lako_b200.model import BUCKET_CENTER, bucket_table
    for bidir in (True, False):
        mine = bucket_table(bidir, 32, 128)
        ref = torch.from_numpy(O.bucket_table(bidir, BUCKET_CENTER + 1))
        assert torch.equal(mine, ref)


def test_no_cpu_fallback():
    from lako_b200 import _native as nat
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2))
    ids, mask, labels = O.synth_batch(1, 2, 16, 4, 512)
    with pytest.raises(nat.LakoError):
        m(input_ids=ids, attention_mask=mask, labels=labels)
    with pytest.raises(nat.LakoError):
        m.generate(ids, mask, max_length=4)
    if not torch.cuda.is_available():
        assert nat.lib().lako_check_device() != 0


@pytest.mark.gpu
def test_crossattention_score_postprocessing_matches_oracle():
    """get_crossattention_scores (src/model.py:143-204; now the device-side lako_fact_scores kernel) on synthetic
    token scores: equal to the oracle's restatement of the reference's host loop."""
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2))
    B, N, L = 3, 2, 48
    ids, mask, _ = O.synth_batch(B, N, L, 4, 512, seed=5, fact_format=True)
    g = torch.Generator().manual_seed(0)
    first = torch.randn(B, N * L, generator=g) * mask.reshape(B, -1)
    last = torch.randn(B, N * L, generator=g) * mask.reshape(B, -1)
    m._tok_first, m._tok_last = first.cuda(), last.cuda()
    for style in ("mean", "max", "21mean"):
        for lh in ("yes", "no"):
            opt = types.SimpleNamespace(stream=2, n_context=5, attention_score_style=style,
                                        use_last_half_layer_attention=lh)
            got = m.get_crossattention_scores(opt, ids, None, mask.cuda())
            tok = (last if lh == "yes" else first + last).view(B, N, L)
            div = (1 if lh == "yes" else 2) * 2
            ref = O.fact_scores_from_token_scores(tok, ids, 5, style, div)
            assert got.dtype == torch.float64 and not got.is_cuda and torch.equal(got, ref)


def test_schedule_matches_oracle():
    from lako_b200.optim import WarmupLinearSchedule
    s = WarmupLinearSchedule(6, 100)
    for step in (0, 1, 5, 6, 7, 50, 99, 100, 150):
        assert abs(s(step) - O.warmup_linear_lambda(step, 6, 100)) < 1e-12


def test_fused_adamw_checkpoint_roundtrip():
    """FusedAdamW.state_dict / load_state_dict (what src/util.py:105-146 saves and restores) on a CPU-resident model:
    only the host-side bookkeeping is exercised here, the update kernel itself needs the GPU."""
    from lako_b200.model import FiDT5, FiDT5Config
    from lako_b200.optim import FusedAdamW, shard_plan
    cfg = FiDT5Config(vocab_size=64, d_model=128, d_ff=256, num_layers=2, num_heads=2)
    model = FiDT5(cfg, seed=0)
    a = FusedAdamW(model, lr=3e-4, weight_decay=0.01, clip=1.0)
    a.m.normal_()
    a.v.uniform_()
    a.step_count = 17
    sd = a.state_dict()
    b = FusedAdamW(model, lr=1.0, weight_decay=0.5, clip=0.0)
    b.load_state_dict(sd)
    assert b.step_count == 17 and b.lr == 3e-4 and b.wd == 0.01 and b.clip == 1.0
    assert torch.equal(b.m, a.m) and torch.equal(b.v, a.v)
    # the gradient-stage ranges of a real layout split evenly over 2 / 4 / 8 ranks (the sharded mode's precondition)
    import ctypes
    from lako_b200 import _native as nat
    ranges = (ctypes.c_size_t * 14)()
    for size in ("base", "large"):
        dims = FiDT5Config.t5(size)  # layouts only: no weights are allocated
        from lako_b200.model import _Dims
        d = _Dims(dims.d_model, dims.num_heads, dims.d_ff, dims.num_layers, dims.vocab_size,
                  dims.relative_attention_num_buckets, dims.layer_norm_epsilon)
        nat.check(nat.lib().lako_fidt5_grad_stages(ctypes.byref(d), ranges), "lako_fidt5_grad_stages")
        rs = [(int(ranges[2 * i]), int(ranges[2 * i + 1])) for i in range(7)]
        for world in (2, 4, 8):
            plan = shard_plan(rs, world, world - 1)
            assert plan is not None, (size, world)
            assert sum(hi - lo for lo, hi, _, _ in plan) == sum(hi - lo for lo, hi in rs)


def test_packing_plan_prefix_masks_and_host_lengths():
    """FiDT5._packing: prefix masks -> cu_seqlens of the valid prefixes (read back, or taken from the Collator's
    host-side lengths without touching the device); any other mask -> padded layout."""
    from lako_b200.model import FiDT5, FiDT5Config
    model = FiDT5(FiDT5Config(vocab_size=64, d_model=128, d_ff=256, num_layers=1, num_heads=2), seed=0)
    B, N, L = 2, 3, 10
    lens = torch.tensor([[10, 4, 7], [1, 9, 5]])
    mask = torch.arange(L)[None, None, :] < lens[:, :, None]
    cu, total = model._packing(mask)
    assert total == int(lens.sum()) and cu.tolist() == [0, 10, 14, 21, 22, 31, 36]
    cu2, total2 = model._packing(mask, lens)          # host lengths: same plan, no read-back
    assert total2 == total and cu2.tolist() == cu.tolist()
    holes = mask.clone()
    holes[0, 0, 3] = False                            # not a prefix mask: cannot be packed
    assert model._packing(holes) == (None, 0)
    model.unpad = False
    assert model._packing(mask) == (None, 0)
    model.unpad = True
    with pytest.raises(Exception):
        model._packing(mask, torch.tensor([1, 2, 3]))  # wrong number of lengths


def test_transformers_302_cross_attention_bias_key_is_kept_and_announced():
    """SURVEY 8c deviation D1: a checkpoint written under transformers 3.0.2 carries decoder block 0's cross-attention
    relative_attention_bias; it must survive load -> state_dict and the load must say what is (not) reproduced."""
    import warnings
    from lako_b200.model import FiDT5, FiDT5Config
    m = FiDT5(FiDT5Config(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2), seed=0)
    sd = m.state_dict()
    assert FiDT5.XATTN_BIAS_KEY not in sd
    sd[FiDT5.XATTN_BIAS_KEY] = torch.randn(32, 2)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.load_state_dict(sd)
    assert len(w) == 1 and "3.0.2" in str(w[0].message)
    assert torch.equal(m.state_dict()[FiDT5.XATTN_BIAS_KEY], sd[FiDT5.XATTN_BIAS_KEY])
    with pytest.raises(RuntimeError):
        m.load_state_dict({**sd, FiDT5.XATTN_BIAS_KEY: torch.zeros(3, 3)})


def test_attention_keep_words_host_restatement():
    """lako_b200/philox.py attn_keep_words / attn_keep_mask (the bit-sliced attention-dropout masks): keep rate, determinism,
    dependence on (seed, step, site), and the word <-> key mapping key j = bit (j >> 1) + 16 (j & 1)."""
    import numpy as np
    from lako_b200.philox import attn_keep_mask, attn_keep_words, thresh_of
    P, H, L = 3, 2, 70
    m = attn_keep_mask(9, 4, 21, P, H, L, 0.1)
    assert m.shape == (P, H, L, 256) and m.dtype == bool
    assert abs(m.mean() - (1 - thresh_of(0.1) / 65536.0)) < 4e-3
    assert (attn_keep_mask(9, 4, 21, P, H, L, 0.1) == m).all()
    for other in ((10, 4, 21), (9, 5, 21), (9, 4, 22)):
        assert (attn_keep_mask(*other, P, H, L, 0.1) != m).any()
    assert attn_keep_mask(9, 4, 21, P, H, L, 0.0).all()
    # one (passage, head, query, chunk): rebuild the 32 keys from the word by hand
    psg, h, q, c = 2, 1, 33, 5
    word = int(attn_keep_words(9, 4, 21, np.array([(((psg * H + h) * L + q) * 32 + c * 4)], dtype=np.uint64), 0.1)[0])
    for j in range(32):
        assert m[psg, h, q, c * 32 + j] == bool((word >> ((j >> 1) + 16 * (j & 1))) & 1)
    # passages longer than 256 tokens (csrc/enc_attn_long.cu): ceil(L / 32) chunks per query row, 4 counters each
    L2 = 300
    m2 = attn_keep_mask(9, 4, 21, 1, H, L2, 0.1)
    assert m2.shape == (1, H, L2, 320)
    h, q, c = 1, 299, 9
    word = int(attn_keep_words(9, 4, 21, np.array([((h * L2 + q) * 40 + c * 4)], dtype=np.uint64), 0.1)[0])
    for j in range(32):
        assert m2[0, h, q, c * 32 + j] == bool((word >> ((j >> 1) + 16 * (j & 1))) & 1)
    # the bit-sliced comparison IS a 16-bit threshold test of 32 independent uniforms: p = 1 - thresh / 65536 exactly
    big = attn_keep_words(1, 2, 3, np.arange(0, 4 * 20000, 4, dtype=np.uint64), 0.25)
    ones = sum(bin(int(w)).count("1") for w in big) / (32.0 * len(big))
    assert abs(ones - 0.75) < 3e-3


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """include/lako_b200.h compiled as plain C (gcc): sizeof and the offset of every field of lako_dims / lako_layout /
    lako_fidt5 / lako_batch equal the ctypes mirrors in lako_b200/model.py -- the binding a maintainer writes from INTEGRATION.md must
    not drift from the header."""
    import ctypes, os, shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from lako_b200 import model as M
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mirrors = {"lako_dims": M._Dims, "lako_layout": M._Layout, "lako_fidt5": M._Model, "lako_batch": M._Batch}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "lako_b200.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, what, val = ln.split()
        cls = mirrors[cname]
        if what == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, what).offset == int(val), (cname, what, getattr(cls, what).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())
