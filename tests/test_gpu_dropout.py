"""Training dropout (p = config.dropout_rate, 0.1 in the reference; sites HF:734, :332, :375, :95, :149, :768).
RNG streams cannot match PyTorch's, so the checks are: (1) kernels apply exactly the Philox keep-mask that
lako_b200/philox.py restates, with torch math on the same mask as reference; (2) forward and backward use the
same masks everywhere (directional-derivative check of the whole model); (3) keep-rate and eval-mode behaviour."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class DropSpec(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_ulonglong), ("step", ctypes.c_uint), ("site", ctypes.c_uint),
                ("thresh", ctypes.c_uint), ("scale", ctypes.c_float)]


def _mask(seed, step, site, shape, p):
    from lako_b200.philox import keep_mask
    n = int(np.prod(shape))
    return torch.from_numpy(keep_mask(seed, step, site, n, p).reshape(shape)).cuda()


def test_philox_keep_rate_and_determinism():
    from lako_b200.philox import keep_mask, scale_of, thresh_of
    m = keep_mask(42, 3, 17, 1 << 20, 0.1)
    assert abs(m.mean() - 0.9) < 2e-3
    assert (keep_mask(42, 3, 17, 1 << 20, 0.1) == m).all()
    assert (keep_mask(42, 4, 17, 1 << 16, 0.1) != m[:1 << 16]).any()
    assert thresh_of(0.1) == 6554 and abs(scale_of(0.1) - 1 / 0.9) < 1e-4


def _model(p_drop, seed=11):
    from oracle import fidt5_oracle as O
    from lako_b200.model import FiDT5, FiDT5Config
    cfg = O.T5_TINY
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=seed, emb_std=0.3), seed=seed + 1)
    m = FiDT5(FiDT5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_ff=cfg.d_ff, num_layers=cfg.n_layers,
                          num_heads=cfg.n_heads, dropout_rate=p_drop))
    m.load_state_dict(sd)
    return m.cuda(), cfg, sd


def test_dropout_changes_train_loss_not_eval():
    from oracle import fidt5_oracle as O
    model, cfg, sd = _model(0.1)
    ids, mask, labels = O.synth_batch(4, 2, 40, 6, cfg.vocab, seed=3)
    args = dict(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
    model.eval()
    with torch.no_grad():
        le = model(**args)[0].item()
        ref = O.fidt5_forward(sd, cfg, ids, mask, labels)[0].item()
    assert abs(le - ref) < 5e-3                      # eval: no dropout
    model.train()
    model.set_dropout_seed(123)
    l1 = model(**args)[0].item()
    l2 = model(**args)[0].item()                     # next step -> different masks
    model.set_dropout_seed(123)
    l1b = model(**args)[0].item()                    # same seed and step -> identical
    assert abs(l1 - l1b) < 1e-5 and abs(l1 - l2) > 1e-5   # (loss reduction order is atomic -> last-bit jitter)
    assert abs(l1 - le) > 1e-4 and abs(l1 - le) < 0.5


def test_forward_backward_use_same_masks_directional_derivative():
    """d/d eps L(theta + eps * g/|g|) at eps = 0 must equal |g| with the dropout masks held fixed."""
    from oracle import fidt5_oracle as O
    model, cfg, _ = _model(0.1)
    ids, mask, labels = O.synth_batch(4, 2, 40, 6, cfg.vocab, seed=5)
    args = dict(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
    model.train()
    model.set_dropout_seed(7)
    model.zero_grad()
    loss = model(**args)[0]
    loss.backward()
    g = model._flat_grad.clone()
    gn = g.norm().item()
    base = model._flat.clone()
    vals = []
    eps = 0.02
    for sgn in (+1, -1):
        with torch.no_grad():
            model._flat.copy_(base + sgn * eps * g / gn)
        model.set_dropout_seed(7)                    # step counter back to 0 -> same masks as the gradient pass
        vals.append(model(**args)[0].item())
    with torch.no_grad():
        model._flat.copy_(base)
    fd = (vals[0] - vals[1]) / (2 * eps)
    print(f"|g| {gn:.4f}  finite difference {fd:.4f}")
    assert abs(fd - gn) < 0.05 * gn
    # and the same check without dropout must hold too (sanity of the method)
    model0, _, _ = _model(0.0)
    model0.train()
    model0.zero_grad()
    model0(**args)[0].backward()
    g0 = model0._flat_grad.clone()
    gn0 = g0.norm().item()
    base0 = model0._flat.clone()
    v = []
    for sgn in (+1, -1):
        with torch.no_grad():
            model0._flat.copy_(base0 + sgn * eps * g0 / gn0)
        v.append(model0(**args)[0].item())
    assert abs((v[0] - v[1]) / (2 * eps) - gn0) < 0.05 * gn0


def test_device_mask_equals_host_restatement():
    """csrc/common.cuh drop_keep8 == lako_b200/philox.py keep_mask, bit for bit."""
    from lako_b200 import _native as nat
    from lako_b200.philox import keep_mask
    lib = nat.lib()
    for seed, step, site, n, p in ((1, 1, 0, 1000, 0.1), (0xDEADBEEFCAFE, 77, 16 * 11 + 4, 123457, 0.1),
                                   (5, 2 ** 31, 3, 4096, 0.25), (9, 9, 9, 64, 0.0)):
        keep = torch.zeros(n, dtype=torch.uint8, device="cuda")
        nat.check(lib.lako_dropout_mask(nat.ptr(keep), ctypes.c_size_t(n), ctypes.c_ulonglong(seed), ctypes.c_uint(step),
                                        ctypes.c_uint(site), ctypes.c_float(p), nat.stream_ptr()))
        ref = torch.from_numpy(keep_mask(seed, step, site, n, p))
        assert torch.equal(keep.cpu().bool(), ref), (seed, step, site)
