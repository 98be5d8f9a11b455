"""Pins oracle/fidt5_oracle.py to outputs of the reference itself (tests/golden/*.pt, written by
oracle/gen_golden.py from /root/reference/src/model.py FiDT5), plus the bucket KAT of SURVEY §4."""
import glob
import os

import pytest
import torch

from oracle import fidt5_oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def load_case(path):
    c = torch.load(path, weights_only=False)
    cfg = O.T5Dims(**c["cfg"])
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=c["seed"], emb_std=c["emb_std"]), seed=c["seed"] + 1)
    if c.get("eos_alias") is not None:
        sd["shared.weight"][1] = sd["shared.weight"][c["eos_alias"]] * 1.05
    return c, cfg, sd


def test_bucket_known_answers():
    # SURVEY.md §4 table, generated from HF:modeling_t5.py:188-234
    rel = torch.tensor([-300, -128, -64, -48, -32, -24, -16, -15, -8, -1, 0, 1, 2, 8, 15, 16, 24, 32, 64, 100, 128])
    exp = [15, 15, 14, 13, 12, 11, 10, 9, 8, 1, 0, 17, 18, 24, 25, 26, 27, 28, 30, 31, 31]
    assert O.relative_position_bucket(rel, True).tolist() == exp
    starts = {0: 0, 8: 8, 12: 9, 16: 10, 23: 11, 32: 12, 46: 13, 64: 14, 91: 15}
    for dist, b in starts.items():
        assert int(O.relative_position_bucket(torch.tensor([-dist]), True)) == b
        if dist > 0:
            assert int(O.relative_position_bucket(torch.tensor([-(dist - 1)]), True)) == (b - 1 if dist > 8 else dist - 1)
    causal_starts = {16: 16, 19: 17, 21: 18, 24: 19, 27: 20, 31: 21, 35: 22, 40: 23, 46: 24, 52: 25, 59: 26,
                     67: 27, 77: 28, 87: 29, 99: 30, 113: 31}
    for dist, b in causal_starts.items():
        assert int(O.relative_position_bucket(torch.tensor([-dist]), False)) == b
        assert int(O.relative_position_bucket(torch.tensor([-(dist - 1)]), False)) == b - 1
    assert O.relative_position_bucket(torch.tensor([1, 5, 100]), False).tolist() == [0, 0, 0]
    tab = O.bucket_table(True, 256)
    assert tab.shape == (511,) and tab[255] == 0 and tab[254] == 1 and tab[256] == 17


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_forward_matches_reference(path):
    c, cfg, sd = load_case(path)
    assert abs(float(sum(v.double().sum() for v in O.unique_params(sd).values())) - c["weights_checksum"]) < 1e-6
    with torch.no_grad():
        loss, logits = O.fidt5_forward(sd, cfg, c["ids"], c["mask"], c["labels"])
    assert torch.allclose(logits, c["logits"], atol=2e-5, rtol=1e-5)
    assert abs(float(loss) - float(c["loss"])) < 1e-5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_backward_matches_reference(path):
    c, cfg, sd = load_case(path)
    params = O.unique_params(sd)
    for p in params.values():
        p.requires_grad_(True)
    loss, _ = O.fidt5_forward(sd, cfg, c["ids"], c["mask"], c["labels"])
    loss.backward()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params.values())).float()
    assert abs(float(gn) - float(c["grad_norm"])) < 1e-4 * float(c["grad_norm"])
    for k, g in c["grads"].items():
        assert torch.allclose(sd[k].grad, g, atol=1e-6, rtol=1e-4), k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_generate_and_scores_match_reference(path):
    c, cfg, sd = load_case(path)
    gen, storage = O.fidt5_generate(sd, cfg, c["ids"], c["mask"], c["max_length"], capture_scores=True)
    assert gen.tolist() == c["generated"].tolist()
    if c["N"] % 2 == 0:
        tok, div = O.token_scores_from_storage(storage, c["mask"], last_half=True)
        assert div == c["tok_scores_div"]
        assert torch.allclose(tok, c["tok_scores_last_half"], atol=1e-4, rtol=1e-5)
        for style in ("mean", "max", "21mean"):
            for lh in ("yes", "no"):
                sc = O.get_crossattention_scores(storage, c["ids"], c["mask"], c["n_context"], style, lh == "yes")
                assert sc.dtype == torch.float64
                assert torch.allclose(sc, c[f"scores_{style}_{lh}"], atol=1e-5, rtol=1e-5), (style, lh)


def test_adamw_and_schedule_semantics():
    # src/util.py:149-168 lambda
    assert O.warmup_linear_lambda(0, 6, 100) == 0.0
    assert abs(O.warmup_linear_lambda(3, 6, 100) - 0.5) < 1e-12
    assert abs(O.warmup_linear_lambda(6, 6, 100) - 1.0) < 1e-12
    assert abs(O.warmup_linear_lambda(53, 6, 100) - 0.5) < 1e-12
    assert O.warmup_linear_lambda(200, 6, 100) == 0.0
    # HF AdamW without bias correction, first step: m=(1-b1)g, v=(1-b2)g^2
    p = torch.tensor([1.0, -2.0]); g = torch.tensor([0.5, 0.25])
    opt = O.HFAdamW([p], lr=0.1, weight_decay=0.01)
    opt.step([g])
    m = 0.1 * g; v = 0.001 * g * g
    exp = torch.tensor([1.0, -2.0]) - 0.1 * m / (v.sqrt() + 1e-6)
    exp = exp - 0.1 * 0.01 * exp
    assert torch.allclose(p, exp, atol=1e-7)
    gs = [torch.tensor([3.0, 4.0])]
    n = O.clip_grad_norm_(gs, 1.0)
    assert abs(float(n) - 5.0) < 1e-6 and torch.allclose(gs[0], torch.tensor([0.6, 0.8]), atol=1e-6)
