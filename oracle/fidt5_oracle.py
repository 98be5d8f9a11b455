"""CPU/fp32 restatement of the LaKo FiDT5 reader hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product (lako_b200/) never does: it fails loudly if the CUDA
library is missing.

What is restated, and from where (reference = /root/reference, HF = the un-vendored
`transformers` T5 code the reference inherits from; README.md:21 pins 3.0.2, this image has
5.5.0 at site-packages/transformers/models/t5/modeling_t5.py):

  FiDT5.forward            src/model.py:39-51     (B,N,L) -> (B,N*L) flatten, then T5 forward
  EncoderWrapper.forward   src/model.py:227-234   (B,N*L) -> (B*N,L) -> encoder -> (B,N*L,d)
  FiDT5.generate           src/model.py:54-60     greedy, start token 0, eos 1, pad 0
  cross_attention_forward  src/model.py:286-349   score_storage = q.k^T + mask (+bias), pre-softmax,
                                                  stored on the first call after a reset (:327-329)
  get_crossattention_scores src/model.py:143-204  + find_index :100, get_attention_score :107
  T5LayerNorm              HF:modeling_t5.py:55-68
  relative position bucket HF:modeling_t5.py:188-234, compute_bias :236-251
  T5Attention              HF:modeling_t5.py:253-344 (no 1/sqrt(dk); fp32 softmax)
  T5 FFN (ReLU)            HF:modeling_t5.py:84-103,135-150
  T5Stack                  HF:modeling_t5.py:617-792 (bias computed in block 0, shared by all layers)
  LM head + loss           HF:modeling_t5.py:1105-1117 (h * d^-0.5 @ shared^T, CE ignore_index=-100)
  _shift_right             HF:modeling_t5.py:595-614
  AdamW(correct_bias=False) src/util.py:185-227 (HF AdamW, removed from 5.5.0; restated from its
                                                  published algorithm)
  WarmupLinearScheduler    src/util.py:149-168
  clip_grad_norm_          train_reader.py:76

Parity pinning: the reference ships NO tests or golden vectors for this path, so
this restatement is pinned against *outputs of the reference itself run in the build container*:
oracle/gen_golden.py imports /root/reference/src/model.py (FiDT5) unmodified + a 2-line shim and
writes tests/golden/*.pt; tests/test_oracle_golden.py checks this file against those fixtures.
"""
from __future__ import annotations

import heapq
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

NEG = torch.finfo(torch.float32).min  # HF 5.5.0 additive mask value (deviation D2 in SURVEY §8c)


@dataclass
class T5Dims:
    d_model: int = 768
    n_heads: int = 12
    d_kv: int = 64
    d_ff: int = 3072
    n_layers: int = 12          # encoder == decoder depth (t5-base/large are symmetric)
    vocab: int = 32128
    n_buckets: int = 32
    max_distance: int = 128
    eps: float = 1e-6

    @property
    def inner(self):
        return self.n_heads * self.d_kv


T5_BASE = T5Dims()
T5_LARGE = T5Dims(d_model=1024, n_heads=16, d_kv=64, d_ff=4096, n_layers=24)
T5_TINY = T5Dims(d_model=128, n_heads=2, d_kv=64, d_ff=256, n_layers=2, vocab=512)


# ----------------------------------------------------------------------------------------------
# state dict with the reference's names
# ----------------------------------------------------------------------------------------------
def enc_prefix(i):
    return f"encoder.encoder.block.{i}.module.layer"


def dec_prefix(i):
    return f"decoder.block.{i}.layer"


def init_state_dict(cfg: T5Dims, seed: int = 0, emb_std: float = 1.0, dtype=torch.float32):
    """HF T5 `_init_weights` distributions (HF:modeling_t5.py:540-590) with a configurable embedding std."""
    g = torch.Generator().manual_seed(seed)
    d, inner, ff = cfg.d_model, cfg.inner, cfg.d_ff

    def n(shape, std):
        return (torch.randn(shape, generator=g, dtype=torch.float32) * std).to(dtype)

    sd = {}
    sd["shared.weight"] = n((cfg.vocab, d), emb_std)

    def attn(prefix, with_bias):
        sd[f"{prefix}.q.weight"] = n((inner, d), (d * cfg.d_kv) ** -0.5)
        sd[f"{prefix}.k.weight"] = n((inner, d), d ** -0.5)
        sd[f"{prefix}.v.weight"] = n((inner, d), d ** -0.5)
        sd[f"{prefix}.o.weight"] = n((d, inner), inner ** -0.5)
        if with_bias:
            sd[f"{prefix}.relative_attention_bias.weight"] = n((cfg.n_buckets, cfg.n_heads), d ** -0.5)

    def ffn(prefix):
        sd[f"{prefix}.DenseReluDense.wi.weight"] = n((ff, d), d ** -0.5)
        sd[f"{prefix}.DenseReluDense.wo.weight"] = n((d, ff), ff ** -0.5)

    for i in range(cfg.n_layers):
        p = enc_prefix(i)
        attn(f"{p}.0.SelfAttention", i == 0)
        sd[f"{p}.0.layer_norm.weight"] = torch.ones(d, dtype=dtype)
        ffn(f"{p}.1")
        sd[f"{p}.1.layer_norm.weight"] = torch.ones(d, dtype=dtype)
    sd["encoder.encoder.final_layer_norm.weight"] = torch.ones(d, dtype=dtype)
    for i in range(cfg.n_layers):
        p = dec_prefix(i)
        attn(f"{p}.0.SelfAttention", i == 0)
        sd[f"{p}.0.layer_norm.weight"] = torch.ones(d, dtype=dtype)
        attn(f"{p}.1.EncDecAttention", False)
        sd[f"{p}.1.layer_norm.weight"] = torch.ones(d, dtype=dtype)
        ffn(f"{p}.2")
        sd[f"{p}.2.layer_norm.weight"] = torch.ones(d, dtype=dtype)
    sd["decoder.final_layer_norm.weight"] = torch.ones(d, dtype=dtype)
    # tied tensors (HF:modeling_t5.py:956-960): one storage, four names
    sd["encoder.encoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["lm_head.weight"] = sd["shared.weight"]
    return sd


def sd_to(sd, device):
    """Copy of the state dict on `device` with the tying of the four embedding names preserved."""
    moved, out = {}, {}
    for k, v in sd.items():
        key = v.data_ptr()
        if key not in moved:
            moved[key] = v.to(device)
        out[k] = moved[key]
    return out


def perturb_norm_weights(sd, seed=1, scale=0.1):
    """Make layer-norm gains non-trivial so tests exercise them (init is all ones)."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if "layer_norm.weight" in k:
            v.add_(torch.randn(v.shape, generator=g) * scale)
    return sd


# ----------------------------------------------------------------------------------------------
# T5 arithmetic
# ----------------------------------------------------------------------------------------------
def relative_position_bucket(rel, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128):
    """HF:modeling_t5.py:188-234; rel = key_pos - query_pos (LongTensor) -> bucket LongTensor."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).long()
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


def bucket_table(bidirectional: bool, max_len: int, num_buckets: int = 32, max_distance: int = 128):
    """bucket for rel in [-(max_len-1), max_len-1]; index = rel + max_len - 1 (int32 numpy).

    This integer table is what the CUDA side consumes, so the float-log bucket boundaries are
    computed exactly once, on the host, with the reference's own float32 expression."""
    rel = torch.arange(-(max_len - 1), max_len, dtype=torch.long)
    return relative_position_bucket(rel, bidirectional, num_buckets, max_distance).to(torch.int32).numpy()


def compute_bias(table, qlen, klen, bidirectional, cfg: T5Dims, q_offset=0):
    """HF:modeling_t5.py:236-251 -> (1,H,qlen,klen)."""
    ctx = torch.arange(qlen, dtype=torch.long, device=table.device)[:, None] + q_offset
    mem = torch.arange(klen, dtype=torch.long, device=table.device)[None, :]
    b = relative_position_bucket(mem - ctx, bidirectional, cfg.n_buckets, cfg.max_distance)
    return table.float()[b.to(table.device)].permute(2, 0, 1).unsqueeze(0)


def rmsnorm(x, w, eps):
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def _heads(x, cfg):
    b, s, _ = x.shape
    return x.view(b, s, cfg.n_heads, cfg.d_kv).transpose(1, 2)


_DROP = {"p": 0.0}  # training dropout of the oracle (torch RNG); 0 for every parity run


def _drop(x):
    return F.dropout(x, p=_DROP["p"], training=True) if _DROP["p"] > 0 else x


def attention(sd, prefix, cfg, hidden, kv_states, bias_plus_mask, capture=None, cached_kv=None):
    """T5Attention (HF:modeling_t5.py:253-344) / cross_attention_forward (src/model.py:286-349).

    bias_plus_mask: additive (b|1, H|1, q, k).  capture: list to which the pre-softmax scores are
    appended (src/model.py:327-329 stores `scores` after `+= mask` and `+= position_bias`)."""
    q = _heads(hidden @ sd[f"{prefix}.q.weight"].T, cfg)
    if cached_kv is not None:
        k, v = cached_kv
    else:
        k = _heads(kv_states @ sd[f"{prefix}.k.weight"].T, cfg)
        v = _heads(kv_states @ sd[f"{prefix}.v.weight"].T, cfg)
    scores = torch.einsum("bnqd,bnkd->bnqk", q, k)
    scores = scores + bias_plus_mask
    if capture is not None:
        capture.append(scores)
    attn = _drop(F.softmax(scores.float(), dim=-1).type_as(scores))   # HF:331-332
    out = torch.matmul(attn, v)
    b, _, s, _ = out.shape
    out = out.transpose(1, 2).contiguous().view(b, s, cfg.inner)
    return out @ sd[f"{prefix}.o.weight"].T, (k, v)


def ffn(sd, prefix, x):
    h = _drop(torch.relu(x @ sd[f"{prefix}.DenseReluDense.wi.weight"].T))   # HF:92-95
    return h @ sd[f"{prefix}.DenseReluDense.wo.weight"].T


def encoder_forward(sd, cfg: T5Dims, ids, mask):
    """T5 encoder stack on (B*N, L) passages.  ids int64, mask bool -> (B*N, L, d)."""
    x = _drop(sd["shared.weight"][ids])   # HF:734
    L = ids.shape[1]
    table = sd[f"{enc_prefix(0)}.0.SelfAttention.relative_attention_bias.weight"]
    bias = compute_bias(table, L, L, True, cfg).to(x.dtype)
    add_mask = torch.zeros(mask.shape, dtype=x.dtype, device=x.device).masked_fill(~mask.bool(), NEG)
    bias = bias + add_mask[:, None, None, :]
    for i in range(cfg.n_layers):
        p = enc_prefix(i)
        n = rmsnorm(x, sd[f"{p}.0.layer_norm.weight"], cfg.eps)
        a, _ = attention(sd, f"{p}.0.SelfAttention", cfg, n, n, bias)
        x = x + _drop(a)   # HF:375
        n = rmsnorm(x, sd[f"{p}.1.layer_norm.weight"], cfg.eps)
        x = x + _drop(ffn(sd, f"{p}.1", n))   # HF:149
    return _drop(rmsnorm(x, sd["encoder.encoder.final_layer_norm.weight"], cfg.eps))   # HF:767-768


def shift_right(labels):
    """HF:modeling_t5.py:595-614: start token 0, -100 -> pad 0."""
    dec = labels.new_zeros(labels.shape)
    dec[:, 1:] = labels[:, :-1]
    dec[:, 0] = 0
    return dec.masked_fill(dec == -100, 0)


def decoder_forward(sd, cfg: T5Dims, dec_ids, memory, mem_mask, cache=None, capture=None):
    """T5 decoder stack.  dec_ids (B,T) (T=1 with cache for cached decode), memory (B,M,d),
    mem_mask (B,M) bool.  cache: dict with 'self' (list of (k,v)), 'cross' (list of (k,v)), 'len'."""
    x = _drop(sd["shared.weight"][dec_ids])
    B, T = dec_ids.shape
    past = 0 if cache is None else cache["len"]
    table = sd[f"{dec_prefix(0)}.0.SelfAttention.relative_attention_bias.weight"]
    total = past + T
    self_bias = compute_bias(table, T, total, False, cfg, q_offset=past).to(x.dtype)
    qpos = torch.arange(T, device=x.device)[:, None] + past
    kpos = torch.arange(total, device=x.device)[None, :]
    causal = torch.zeros(T, total, dtype=x.dtype, device=x.device).masked_fill(kpos > qpos, NEG)
    self_bias = self_bias + causal[None, None]
    cross_bias = torch.zeros(mem_mask.shape, dtype=x.dtype, device=x.device).masked_fill(~mem_mask.bool(), NEG)[:, None, None, :]
    for i in range(cfg.n_layers):
        p = dec_prefix(i)
        n = rmsnorm(x, sd[f"{p}.0.layer_norm.weight"], cfg.eps)
        if cache is None:
            a, _ = attention(sd, f"{p}.0.SelfAttention", cfg, n, n, self_bias)
        else:
            pre = f"{p}.0.SelfAttention"
            k_new = _heads(n @ sd[f"{pre}.k.weight"].T, cfg)
            v_new = _heads(n @ sd[f"{pre}.v.weight"].T, cfg)
            if cache["self"][i] is not None:
                k_all = torch.cat([cache["self"][i][0], k_new], dim=2)
                v_all = torch.cat([cache["self"][i][1], v_new], dim=2)
            else:
                k_all, v_all = k_new, v_new
            cache["self"][i] = (k_all, v_all)
            a, _ = attention(sd, pre, cfg, n, None, self_bias, cached_kv=(k_all, v_all))
        x = x + _drop(a)
        n = rmsnorm(x, sd[f"{p}.1.layer_norm.weight"], cfg.eps)
        cap = None
        if capture is not None and capture[i] is None:
            cap = []
        ckv = None if cache is None else cache["cross"][i]
        a, kv = attention(sd, f"{p}.1.EncDecAttention", cfg, n, memory, cross_bias, capture=cap, cached_kv=ckv)
        if cache is not None:
            cache["cross"][i] = kv
        if cap is not None:
            capture[i] = cap[0]
        x = x + _drop(a)
        n = rmsnorm(x, sd[f"{p}.2.layer_norm.weight"], cfg.eps)
        x = x + _drop(ffn(sd, f"{p}.2", n))
    if cache is not None:
        cache["len"] = total
    return _drop(rmsnorm(x, sd["decoder.final_layer_norm.weight"], cfg.eps))


def lm_logits(sd, cfg, h):
    return (h * (cfg.d_model ** -0.5)) @ sd["shared.weight"].T


def fidt5_encode(sd, cfg, input_ids, attention_mask):
    """FiDT5.forward + EncoderWrapper.forward reshapes (src/model.py:39-51, 227-234)."""
    B, N, L = input_ids.shape
    enc = encoder_forward(sd, cfg, input_ids.reshape(B * N, L), attention_mask.reshape(B * N, L))
    return enc.view(B, N * L, cfg.d_model), attention_mask.reshape(B, N * L)


def fidt5_forward(sd, cfg: T5Dims, input_ids, attention_mask, labels):
    """-> (loss, logits(B,T,V)) as `FiDT5.forward(...)[0:2]` (train_reader.py:67-71)."""
    memory, mem_mask = fidt5_encode(sd, cfg, input_ids, attention_mask)
    h = decoder_forward(sd, cfg, shift_right(labels), memory, mem_mask)
    logits = lm_logits(sd, cfg, h)
    loss = F.cross_entropy(logits.view(-1, logits.size(-1)), labels.reshape(-1), ignore_index=-100)
    return loss, logits


@torch.no_grad()
def fidt5_generate(sd, cfg: T5Dims, input_ids, attention_mask, max_length, capture_scores=False,
                   return_step_logits=False, min_length=0):
    """Greedy decode as FiDT5.generate (src/model.py:54-60 -> HF greedy).

    Returns (ids (B,<=max_length) int64, score_storage list[Ld] of (B,H,1,M) or None[, logits list])."""
    memory, mem_mask = fidt5_encode(sd, cfg, input_ids, attention_mask)
    B = input_ids.shape[0]
    cache = {"self": [None] * cfg.n_layers, "cross": [None] * cfg.n_layers, "len": 0}
    capture = [None] * cfg.n_layers if capture_scores else None
    dev = input_ids.device
    out = torch.zeros(B, 1, dtype=torch.long, device=dev)
    unfinished = torch.ones(B, dtype=torch.long, device=dev)
    step_logits = []
    while out.shape[1] < max_length:
        h = decoder_forward(sd, cfg, out[:, -1:], memory, mem_mask, cache=cache, capture=capture)
        logits = lm_logits(sd, cfg, h)[:, -1]
        if return_step_logits:
            step_logits.append(logits)
        if out.shape[1] < min_length:   # HF MinLengthLogitsProcessor: no EOS while the decoder length is below min_length
            logits = logits.clone()
            logits[:, 1] = -float("inf")
        nxt = logits.argmax(-1)
        nxt = nxt * unfinished + 0 * (1 - unfinished)
        out = torch.cat([out, nxt[:, None]], dim=1)
        unfinished = unfinished * (nxt != 1).long()
        if unfinished.max() == 0:
            break
    if return_step_logits:
        return out, capture, step_logits
    return out, capture


# ----------------------------------------------------------------------------------------------
# score post-processing (src/model.py:100-204), restated
# ----------------------------------------------------------------------------------------------
def find_index(fact_list, start):
    try:
        return fact_list.index(5, start)
    except ValueError:
        return -1


def get_attention_score(scores, start, end, style):
    if style == "mean":
        return sum(scores[start:end]) / (end - start)
    if style == "max":
        return max(scores[start:end])
    if style == "21mean":
        num = max(int((end - start + 1) / 2), 1)
        return sum(heapq.nlargest(num, scores[start:end])) / num
    raise ValueError(style)


def token_scores_from_storage(score_storage, context_mask, last_half: bool):
    """src/model.py:152-167 up to the heads+layers sum, for ALL passages: -> (B,N,L) fp32 and the
    divisor n_layers*n_heads (:202)."""
    scores = torch.cat(score_storage, dim=2)
    if last_half:
        _, scores = torch.chunk(scores, 2, dim=2)
    bsz, n_heads, n_layers, _ = scores.shape
    n_passages = context_mask.size(1)
    scores = scores.view(bsz, n_heads, n_layers, n_passages, -1)
    scores = scores.masked_fill(~context_mask[:, None, None].bool(), 0.0)
    return scores.sum(dim=[1, 2]), n_layers * n_heads


def fact_scores_from_token_scores(tok_scores, context_ids, n_context, style, divisor):
    """src/model.py:166-203: second half of the passages, passage 1's tokens split on id 5."""
    _, fact_scores = torch.chunk(tok_scores, 2, dim=1)
    out = []
    for b in range(tok_scores.shape[0]):
        fact_list = context_ids[b][1].cpu().numpy().tolist()
        sc = fact_scores[b][0].cpu().numpy().tolist()
        res, start = [], 2
        for _ in range(n_context):
            end = find_index(fact_list, start) + 1
            if end == 0:
                break
            res.append(get_attention_score(sc, start, end, style))
            start = end
        if len(res) < n_context and fact_list[-1] != 0:
            end = len(fact_list)
            if end > start:
                res.append(get_attention_score(sc, start, end, style))
        while len(res) < n_context:
            res.append(-5)
        out.append(np.array(res))
    return torch.from_numpy(np.array(out)) / divisor


def get_crossattention_scores(score_storage, context_ids, context_mask, n_context, style="21mean",
                              last_half=True):
    tok, div = token_scores_from_storage(score_storage, context_mask, last_half)
    return fact_scores_from_token_scores(tok, context_ids, n_context, style, div)


# ----------------------------------------------------------------------------------------------
# optimizer / schedule (src/util.py:149-245, train_reader.py:76-79)
# ----------------------------------------------------------------------------------------------
def warmup_linear_lambda(step, warmup_steps, scheduler_steps, min_ratio=0.0, fixed_lr=False):
    if step < warmup_steps:
        return (1 - min_ratio) * step / float(max(1, warmup_steps)) + min_ratio
    if fixed_lr:
        return 1.0
    return max(0.0, 1.0 + (min_ratio - 1) * (step - warmup_steps) / float(max(1.0, scheduler_steps - warmup_steps)))


def clip_grad_norm_(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_: g *= min(1, max_norm/(||g||+1e-6)); returns the norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


class HFAdamW:
    """transformers.AdamW(correct_bias=False) as used at src/util.py:225:
    m=b1*m+(1-b1)*g; v=b2*v+(1-b2)*g^2; p -= lr*m/(sqrt(v)+eps); then p -= lr*wd*p."""

    def __init__(self, params, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-6):
        self.params = list(params)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    @torch.no_grad()
    def step(self, grads, lr_scale=1.0):
        lr = self.lr * lr_scale
        b1, b2 = self.betas
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.addcdiv_(m, v.sqrt().add_(self.eps), value=-lr)
            if self.wd > 0:
                p.add_(p, alpha=-lr * self.wd)


def unique_params(sd):
    """name -> tensor for the untied parameter set, in state-dict order (tied names dropped)."""
    seen, out = set(), {}
    for k, v in sd.items():
        if v.data_ptr() in seen:
            continue
        seen.add(v.data_ptr())
        out[k] = v
    return out


def train_step(sd, cfg, opt: HFAdamW, input_ids, attention_mask, labels, lr_scale, clip=1.0, dropout=0.0):
    """One iteration of train_reader.py:67-79 on the oracle. `sd` tensors must require grad.
    dropout > 0 reproduces model.train() (p = 0.1 in the reference) with torch's RNG; parity runs use 0."""
    params = opt.params
    for p in params:
        p.grad = None
    _DROP["p"] = dropout
    try:
        loss, _ = fidt5_forward(sd, cfg, input_ids, attention_mask, labels)
    finally:
        _DROP["p"] = 0.0
    loss.backward()
    grads = [p.grad for p in params]
    gnorm = clip_grad_norm_(grads, clip)
    opt.step(grads, lr_scale)
    return loss.detach(), gnorm


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_batch(B, N, L, T, vocab, seed=1234, fact_format=False):
    g = torch.Generator().manual_seed(seed)
    hi = min(32000, vocab)
    ids = torch.randint(11, hi, (B, N, L), generator=g)
    lens = torch.randint(int(0.6 * L), L + 1, (B, N), generator=g)
    pos = torch.arange(L)[None, None, :]
    mask = pos < lens[:, :, None]
    ids = ids * mask
    if fact_format and N >= 2:
        # passage 1 looks like "fact: s1 . s2 . ..." : 2 prefix tokens then id 5 every 8-20 tokens
        for b in range(B):
            ids[b, 1, 0], ids[b, 1, 1] = 6, 10
            p = 2
            while True:
                p += int(torch.randint(8, 21, (1,), generator=g))
                if p >= int(lens[b, 1]):
                    break
                ids[b, 1, p] = 5
    tl = torch.randint(2, T + 1, (B,), generator=g)
    labels = torch.randint(11, hi, (B, T), generator=g)
    tpos = torch.arange(T)[None, :]
    labels = torch.where(tpos == (tl[:, None] - 1), torch.ones_like(labels), labels)
    labels = labels.masked_fill(tpos >= tl[:, None], -100)
    return ids, mask, labels
