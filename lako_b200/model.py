"""FiDT5 -- drop-in for the reference's `src.model.FiDT5` (/root/reference/src/model.py:20-214), with
every device operation running in the hand-written sm_100a kernels of liblako_b200.so.

Kept from the reference's Python surface:
  forward(input_ids (B,N,L), attention_mask (B,N,L), labels (B,T)) -> (loss, logits)   model.py:39-51
  generate(input_ids, attention_mask, max_length) -> int64 (B, <=max_length)           model.py:54-60
  overwrite_forward_crossattention / reset_score_storage / get_crossattention_scores   model.py:92-98,143-213
  load_t5 / wrap_encoder / unwrap_encoder / set_checkpoint                             model.py:62-90
  state_dict()/load_state_dict() with the reference's key names (encoder.encoder.block.N.module...)
  save_pretrained / from_pretrained, parameters(), .cuda()/.train()/.eval(), encoder.n_passages

There is no PyTorch/CPU fallback: without the CUDA library, or off sm_100, calls raise.
"""
from __future__ import annotations

import ctypes
import heapq
import json
import math
import os
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from . import _native as nat

MAX_LAYERS = 48


class _Dims(ctypes.Structure):
    _fields_ = [("d_model", ctypes.c_int), ("n_heads", ctypes.c_int), ("d_ff", ctypes.c_int),
                ("n_layers", ctypes.c_int), ("vocab", ctypes.c_int), ("n_buckets", ctypes.c_int),
                ("eps", ctypes.c_float)]


_SZ = ctypes.c_size_t * MAX_LAYERS


class _Layout(ctypes.Structure):
    _fields_ = [("shared", ctypes.c_size_t),
                ("enc_ln0", _SZ), ("enc_qkv", _SZ), ("enc_o", _SZ), ("enc_ln1", _SZ), ("enc_wi", _SZ), ("enc_wo", _SZ),
                ("enc_relbias", ctypes.c_size_t), ("enc_final_ln", ctypes.c_size_t),
                ("dec_ln0", _SZ), ("dec_qkv", _SZ), ("dec_o", _SZ), ("dec_ln1", _SZ), ("dec_xq", _SZ), ("dec_xo", _SZ),
                ("dec_ln2", _SZ), ("dec_wi", _SZ), ("dec_wo", _SZ),
                ("dec_xkv", ctypes.c_size_t), ("dec_relbias", ctypes.c_size_t), ("dec_final_ln", ctypes.c_size_t),
                ("total", ctypes.c_size_t)]


class _Model(ctypes.Structure):
    _fields_ = [("dims", _Dims), ("w_f32", ctypes.c_void_p), ("w_bf16", ctypes.c_void_p), ("g_f32", ctypes.c_void_p),
                ("enc_buckets", ctypes.c_void_p), ("dec_buckets", ctypes.c_void_p), ("bucket_center", ctypes.c_int),
                ("w_tf32_hi", ctypes.c_void_p), ("w_tf32_lo", ctypes.c_void_p)]


class _Batch(ctypes.Structure):
    _fields_ = [("ids", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("labels", ctypes.c_void_p),
                ("B", ctypes.c_int), ("N", ctypes.c_int), ("L", ctypes.c_int), ("T", ctypes.c_int),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
                ("logits", ctypes.c_void_p), ("stats", ctypes.c_void_p), ("count_is_set", ctypes.c_int),
                ("dropout_p", ctypes.c_float), ("dropout_step", ctypes.c_uint), ("dropout_seed", ctypes.c_ulonglong),
                ("cu_seqlens", ctypes.c_void_p), ("packed_tokens", ctypes.c_int), ("stage_events", ctypes.c_void_p),
                ("labels_are_decoder_ids", ctypes.c_int), ("precision", ctypes.c_int), ("recompute", ctypes.c_int),
                ("weight_events", ctypes.c_void_p)]


class FiDT5Config:
    """The subset of transformers.T5Config the reader path reads (t5-base defaults)."""

    def __init__(self, vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=None,
                 num_heads=12, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                 dropout_rate=0.1, layer_norm_epsilon=1e-6, initializer_factor=1.0, feed_forward_proj="relu",
                 pad_token_id=0, eos_token_id=1, decoder_start_token_id=0, **unused):
        self.vocab_size, self.d_model, self.d_kv, self.d_ff = vocab_size, d_model, d_kv, d_ff
        self.num_layers = num_layers
        self.num_decoder_layers = num_layers if num_decoder_layers is None else num_decoder_layers
        self.num_heads = num_heads
        self.relative_attention_num_buckets = relative_attention_num_buckets
        self.relative_attention_max_distance = relative_attention_max_distance
        self.dropout_rate = dropout_rate
        self.layer_norm_epsilon = layer_norm_epsilon
        self.initializer_factor = initializer_factor
        self.feed_forward_proj = feed_forward_proj
        self.pad_token_id, self.eos_token_id, self.decoder_start_token_id = pad_token_id, eos_token_id, decoder_start_token_id

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        if isinstance(cfg, dict):
            return cls(**cfg)
        return cls(**{k: getattr(cfg, k) for k in (
            "vocab_size", "d_model", "d_kv", "d_ff", "num_layers", "num_decoder_layers", "num_heads",
            "relative_attention_num_buckets", "relative_attention_max_distance", "dropout_rate",
            "layer_norm_epsilon", "initializer_factor", "feed_forward_proj") if hasattr(cfg, k)})

    def to_dict(self):
        return dict(self.__dict__)

    @classmethod
    def t5(cls, size):
        table = {"small": dict(d_model=512, d_ff=2048, num_layers=6, num_heads=8),
                 "base": dict(d_model=768, d_ff=3072, num_layers=12, num_heads=12),
                 "large": dict(d_model=1024, d_ff=4096, num_layers=24, num_heads=16)}
        return cls(**table[size])


def relative_position_bucket(rel, bidirectional, num_buckets=32, max_distance=128):
    """Integer bucket of rel = key_pos - query_pos, evaluated with the same float32 expression as
    HF:modeling_t5.py:188-234 so that bucket boundaries agree bit-for-bit."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


BUCKET_CENTER = 2047  # offsets covered by the device bucket tables: text_maxlength and max_length up to 2048


def bucket_table(bidirectional, num_buckets, max_distance):
    rel = torch.arange(-BUCKET_CENTER, BUCKET_CENTER + 1, dtype=torch.long)
    return relative_position_bucket(rel, bidirectional, num_buckets, max_distance).to(torch.int32)


class _Node(nn.Module):
    """Bare container used to reproduce the reference's module tree (and so its state_dict names)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("lako_b200 sub-modules are containers; call FiDT5.forward/generate")


def _enc_name(i, wrapped=True):
    return f"encoder.encoder.block.{i}.module.layer" if wrapped else f"encoder.block.{i}.layer"


class ReaderOutput(tuple):
    """What FiDT5.forward returns: indexable like the reference's tuple (`model(...)[0]` is the loss,
    train_reader.py:67-71) and attribute-readable like HF's Seq2SeqLMOutput (.loss, .logits)."""

    def __new__(cls, loss, logits):
        self = super().__new__(cls, (logits,) if loss is None else (loss, logits))
        self.loss, self.logits = loss, logits
        return self


class _LossFn(torch.autograd.Function):
    """Ties `loss.backward()` (train_reader.py:73) to lako_fidt5_backward: gradients are accumulated by
    the CUDA side directly into the flat gradient buffer that every Parameter.grad is a view of."""

    @staticmethod
    def forward(ctx, anchor, model, batch_keep, loss, logits):
        ctx.model = model
        ctx.batch_keep = batch_keep
        ctx.mark_non_differentiable(logits)
        return loss.clone(), logits

    @staticmethod
    def backward(ctx, grad_loss, _grad_logits):
        ctx.model._run_backward(ctx.batch_keep, grad_loss)
        return torch.zeros(1, device=grad_loss.device), None, None, None, None


class FiDT5(nn.Module):
    def __init__(self, config, emb_std=None, seed=None):
        super().__init__()
        self.config = FiDT5Config.from_any(config)
        c = self.config
        if c.d_kv != 64:
            raise nat.LakoError("lako_b200 kernels are specialised for d_kv = 64 (every public T5 size)")
        if c.num_decoder_layers != c.num_layers:
            raise nat.LakoError("encoder/decoder depth must match (t5-small/base/large/3b/11b)")
        if "relu" not in c.feed_forward_proj or "gated" in c.feed_forward_proj:
            raise nat.LakoError("only the ReLU FFN of T5 v1.0 (what LaKo uses) is implemented")
        self._dims = _Dims(c.d_model, c.num_heads, c.d_ff, c.num_layers, c.vocab_size,
                           c.relative_attention_num_buckets, c.layer_norm_epsilon)
        self._layout = _Layout()
        nat.check(nat.lib().lako_fidt5_layout(ctypes.byref(self._dims), ctypes.byref(self._layout)), "lako_fidt5_layout")
        self._total = int(self._layout.total)
        self._flat = torch.zeros(self._total, dtype=torch.float32)
        self._flat_grad = None
        self._bf16 = None
        self._bf16_version = -1
        self._tf32 = None
        self._tf32_version = -1
        self._enc_buckets = bucket_table(True, c.relative_attention_num_buckets, c.relative_attention_max_distance)
        self._dec_buckets = bucket_table(False, c.relative_attention_num_buckets, c.relative_attention_max_distance)
        self._specs = self._param_specs()
        self._build_tree()
        self._anchor = torch.zeros(1, requires_grad=True)
        self._ws = {}
        self._ws_gen = 0
        self.precision = "bf16"   # "bf16" (bf16 operands, fp32 accumulate) | "fp32" (evaluation only, see set_precision)
        self._capture = False
        self._tok_first = self._tok_last = None
        self.use_checkpoint = False
        # training dropout (config.dropout_rate, 0.1 for every T5 size): Philox masks keyed by (seed, step)
        self._drop_seed = None
        self._drop_step = 0
        # un-padded execution: when every passage mask is a prefix mask (what src/data.py's Collator emits), the
        # encoder side runs on the valid tokens only.  Set to False to force the padded layout.
        self.unpad = True
        # data-parallel overlap: CUDA events recorded by backward as gradient ranges become final (optim.py)
        self._stage_events = None
        self._stage_ranges = None
        self._master_sync = None   # set by a sharded FusedAdamW: all-gathers the fp32 masters before they are read
        self._pending_wev = None   # (events, handle array): an all-gather of updated weights still running on another stream
        # transformers 3.0.2 (the version the reference pins, README.md:21) gives decoder block 0's cross-attention its
        # own relative_attention_bias table (32, H); later versions have none.  Kept when a checkpoint carries it.
        self._xattn_relbias = None
        self._init_weights(emb_std, seed)

    # ------------------------------------------------------------------ parameters
    def _param_specs(self):
        """reference-name -> (offset, shape) into the flat buffer."""
        c, lo = self.config, self._layout
        d, inner, ff, H, nb = c.d_model, c.num_heads * 64, c.d_ff, c.num_heads, c.relative_attention_num_buckets
        s = {"shared.weight": (lo.shared, (c.vocab_size, d))}
        for i in range(c.num_layers):
            p = _enc_name(i)
            s[f"{p}.0.SelfAttention.q.weight"] = (lo.enc_qkv[i], (inner, d))
            s[f"{p}.0.SelfAttention.k.weight"] = (lo.enc_qkv[i] + inner * d, (inner, d))
            s[f"{p}.0.SelfAttention.v.weight"] = (lo.enc_qkv[i] + 2 * inner * d, (inner, d))
            s[f"{p}.0.SelfAttention.o.weight"] = (lo.enc_o[i], (d, inner))
            if i == 0:
                s[f"{p}.0.SelfAttention.relative_attention_bias.weight"] = (lo.enc_relbias, (nb, H))
            s[f"{p}.0.layer_norm.weight"] = (lo.enc_ln0[i], (d,))
            s[f"{p}.1.DenseReluDense.wi.weight"] = (lo.enc_wi[i], (ff, d))
            s[f"{p}.1.DenseReluDense.wo.weight"] = (lo.enc_wo[i], (d, ff))
            s[f"{p}.1.layer_norm.weight"] = (lo.enc_ln1[i], (d,))
        s["encoder.encoder.final_layer_norm.weight"] = (lo.enc_final_ln, (d,))
        for i in range(c.num_layers):
            p = f"decoder.block.{i}.layer"
            s[f"{p}.0.SelfAttention.q.weight"] = (lo.dec_qkv[i], (inner, d))
            s[f"{p}.0.SelfAttention.k.weight"] = (lo.dec_qkv[i] + inner * d, (inner, d))
            s[f"{p}.0.SelfAttention.v.weight"] = (lo.dec_qkv[i] + 2 * inner * d, (inner, d))
            s[f"{p}.0.SelfAttention.o.weight"] = (lo.dec_o[i], (d, inner))
            if i == 0:
                s[f"{p}.0.SelfAttention.relative_attention_bias.weight"] = (lo.dec_relbias, (nb, H))
            s[f"{p}.0.layer_norm.weight"] = (lo.dec_ln0[i], (d,))
            s[f"{p}.1.EncDecAttention.q.weight"] = (lo.dec_xq[i], (inner, d))
            s[f"{p}.1.EncDecAttention.k.weight"] = (lo.dec_xkv + (2 * i) * inner * d, (inner, d))
            s[f"{p}.1.EncDecAttention.v.weight"] = (lo.dec_xkv + (2 * i + 1) * inner * d, (inner, d))
            s[f"{p}.1.EncDecAttention.o.weight"] = (lo.dec_xo[i], (d, inner))
            s[f"{p}.1.layer_norm.weight"] = (lo.dec_ln1[i], (d,))
            s[f"{p}.2.DenseReluDense.wi.weight"] = (lo.dec_wi[i], (ff, d))
            s[f"{p}.2.DenseReluDense.wo.weight"] = (lo.dec_wo[i], (d, ff))
            s[f"{p}.2.layer_norm.weight"] = (lo.dec_ln2[i], (d,))
        s["decoder.final_layer_norm.weight"] = (lo.dec_final_ln, (d,))
        return s

    TIED = ("encoder.encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight")

    def _view(self, flat, name):
        off, shape = self._specs[name]
        n = int(np.prod(shape))
        return flat[off:off + n].view(shape)

    def _register(self, name, param):
        parts = name.split(".")
        node = self
        for comp in parts[:-1]:
            if comp not in node._modules:
                node.add_module(comp, _Node())
            node = node._modules[comp]
        node.register_parameter(parts[-1], param)

    def _build_tree(self):
        self._params = {}
        for name in self._specs:
            p = nn.Parameter(self._view(self._flat, name))
            self._params[name] = p
            self._register(name, p)
        for name in self.TIED:
            self._register(name, self._params["shared.weight"])
        self.encoder.n_passages = 1

    def _rebind(self):
        for name, p in self._params.items():
            p.data = self._view(self._flat, name)
            if self._flat_grad is not None:
                p.grad = self._view(self._flat_grad, name)

    def _apply(self, fn, recurse=True):
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise nat.LakoError("FiDT5 keeps fp32 master weights; bf16 compute copies are managed internally")
        self._flat = new_flat.contiguous()
        if self._flat_grad is not None:
            self._flat_grad = fn(self._flat_grad)
        self._enc_buckets = fn(self._enc_buckets)
        self._dec_buckets = fn(self._dec_buckets)
        if self._xattn_relbias is not None:
            self._xattn_relbias = fn(self._xattn_relbias)
        with torch.no_grad():
            self._anchor = fn(self._anchor.detach()).requires_grad_(True)
        self._bf16 = None
        self._bf16_version = -1
        self._tf32 = None
        self._tf32_version = -1
        self._ws = {}
        self._tok_first = self._tok_last = None
        self._rebind()
        return self

    @torch.no_grad()
    def _init_weights(self, emb_std=None, seed=None):
        """HF T5 `_init_weights` distributions (HF:modeling_t5.py:540-590)."""
        c = self.config
        f = c.initializer_factor
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        d, inner, ff = c.d_model, c.num_heads * 64, c.d_ff

        def normal_(p, std):
            p.copy_(torch.randn(p.shape, generator=g) * std)

        for name, p in self._params.items():
            if name == "shared.weight":
                normal_(p, f * 1.0 if emb_std is None else emb_std)
            elif "layer_norm" in name:
                p.fill_(f)
            elif name.endswith(".q.weight"):
                normal_(p, f * (d * c.d_kv) ** -0.5)
            elif name.endswith(".k.weight") or name.endswith(".v.weight") or "relative_attention_bias" in name:
                normal_(p, f * d ** -0.5)
            elif name.endswith(".o.weight"):
                normal_(p, f * inner ** -0.5)
            elif name.endswith("wi.weight"):
                normal_(p, f * d ** -0.5)
            elif name.endswith("wo.weight"):
                normal_(p, f * ff ** -0.5)

    # ------------------------------------------------------------------ state dict plumbing
    def state_dict(self, *args, **kwargs):
        self._wait_weights()
        if self._master_sync is not None:   # sharded data-parallel optimizer: a collective, call it on every rank
            self._master_sync()
        sd = super().state_dict(*args, **kwargs)
        if self._xattn_relbias is not None:
            sd[kwargs.get("prefix", "") + self.XATTN_BIAS_KEY] = self._xattn_relbias.detach().clone()
        return sd

    XATTN_BIAS_KEY = "decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight"

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = set(self._specs) | set(self.TIED)
        ignore = {self.XATTN_BIAS_KEY}  # 3.0.2-era key (SURVEY 8c deviation D1), handled below
        missing = [k for k in self._specs if k not in state_dict and k != "shared.weight"]
        if "shared.weight" not in state_dict and not any(t in state_dict for t in self.TIED):
            missing.append("shared.weight")
        unexpected = [k for k in state_dict if k not in own and k not in ignore]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        with torch.no_grad():
            for k, v in state_dict.items():
                if k in self._specs:
                    self._params[k].copy_(v.to(torch.float32))
                elif k in self.TIED and "shared.weight" not in state_dict:
                    self._params["shared.weight"].copy_(v.to(torch.float32))
        self._xattn_relbias = None
        if self.XATTN_BIAS_KEY in state_dict:
            t = state_dict[self.XATTN_BIAS_KEY].detach().to(torch.float32)
            if tuple(t.shape) != (self.config.relative_attention_num_buckets, self.config.num_heads):
                raise RuntimeError(f"load_state_dict: {self.XATTN_BIAS_KEY} has shape {tuple(t.shape)}")
            self._xattn_relbias = t.clone().to(self._flat.device)
            if bool((t != 0).any()):
                import warnings
                warnings.warn(
                    "this checkpoint carries the cross-attention relative_attention_bias of transformers 3.0.2 (the version "
                    "the reference pins).  generate() and get_crossattention_scores() reproduce it exactly: at every "
                    "cached decode step it is the per-head constant table[0, h] (src/model.py:323-325 calls "
                    "compute_bias(qlen=1, klen)), which leaves the softmax unchanged and shifts the stored step-0 "
                    "scores.  forward(labels=...) does NOT apply its teacher-forced form (a t-dependent bias on the "
                    "first T-1 memory tokens), so training-time logits differ slightly from a 3.0.2 run, and the table "
                    "is carried through state_dict() unchanged (it receives no gradient).", stacklevel=2)
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    # the reference wraps/unwraps the encoder to load plain T5 checkpoints (model.py:62-82); here it is a
    # key translation since there is no nn.Module encoder to swap.
    def wrap_encoder(self, use_checkpoint=False):
        self.use_checkpoint = use_checkpoint

    def unwrap_encoder(self):
        pass

    @staticmethod
    def _t5_to_fid_key(k):
        if k.startswith("encoder.block."):
            parts = k.split(".")
            return "encoder.encoder.block." + parts[2] + ".module." + ".".join(parts[3:])
        if k.startswith("encoder.") and not k.startswith("encoder.encoder."):
            return "encoder." + k
        return k

    def load_t5(self, state_dict):
        self.load_state_dict({self._t5_to_fid_key(k): v for k, v in state_dict.items()})

    def set_checkpoint(self, use_checkpoint):
        """model.py:84-90 (--use_checkpoint): recompute the encoder layers' activations in backward instead of keeping
        them (lako_batch.recompute).  Same loss and gradients; ~7x less encoder activation memory for ~25 % more step
        time.  Not needed at 180 GB for the reference's shapes (cfg4 trains in 45 GB); it is what lets N = 100 passages
        at B = 8 train on one GPU."""
        self.use_checkpoint = bool(use_checkpoint)

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.config.to_dict(), f)
        torch.save({k: v.detach().cpu() for k, v in self.state_dict().items()}, os.path.join(path, "pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, path):
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        model = cls(FiDT5Config(**cfg))
        sd = cls._read_checkpoint(path)
        if any(k.startswith("encoder.block.") for k in sd):   # a plain T5 checkpoint written by transformers
            model.load_t5(sd)
        else:
            model.load_state_dict(sd)
        return model

    @staticmethod
    def _read_checkpoint(path):
        """Every on-disk form `PreTrainedModel.save_pretrained` produces: pytorch_model.bin, model.safetensors, or
        either one sharded behind a *.index.json (src/util.py:128 loads the reader through from_pretrained)."""
        def load_one(fn):
            if fn.endswith(".safetensors"):
                from safetensors.torch import load_file
                return load_file(fn, device="cpu")
            return torch.load(fn, map_location="cpu", weights_only=True)
        for name in ("model.safetensors", "pytorch_model.bin"):
            fn = os.path.join(path, name)
            if os.path.exists(fn):
                return load_one(fn)
            idx = fn + ".index.json"
            if os.path.exists(idx):
                with open(idx) as f:
                    shards = sorted(set(json.load(f)["weight_map"].values()))
                sd = {}
                for sh in shards:
                    sd.update(load_one(os.path.join(path, sh)))
                return sd
        raise FileNotFoundError(f"no pytorch_model.bin / model.safetensors (or sharded index) under {path}")

    # ------------------------------------------------------------------ device state
    def _require_cuda(self):
        if not self._flat.is_cuda:
            raise nat.LakoError("FiDT5 runs only on a B200 (sm_100a): call .cuda() first; there is no CPU path")

    def _ensure_grad(self):
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
            self._rebind()

    def zero_grad(self, set_to_none=False):
        if self._flat_grad is not None:
            self._flat_grad.zero_()
            for name, p in self._params.items():
                if p.grad is None:
                    p.grad = self._view(self._flat_grad, name)

    def _wait_weights(self):
        """A sharded optimizer may still be all-gathering the updated weights on its communication stream
        (FusedAdamW._step_sharded): make the current stream wait for all of it.  The training forward does not come here --
        it hands the events to the C side, which waits range by range right before first use."""
        if self._pending_wev is not None:
            cur = torch.cuda.current_stream()
            for e in self._pending_wev[0]:
                if e is not None:
                    cur.wait_event(e)
            self._pending_wev = None

    def _refresh_bf16(self):
        sig = self._weights_signature()
        if self._bf16 is None or self._bf16_version != sig:
            self._wait_weights()
        if self._master_sync is not None and self._bf16 is not None and self._bf16_version != sig:
            # sharded optimizer + someone touched the masters: foreign slices of the masters are stale, and casting them
            # would overwrite the correct all-gathered bf16 weights -- bring the masters up to date first (collective)
            self._master_sync()
        if self._bf16 is None:
            self._bf16 = torch.empty(self._total, dtype=torch.bfloat16, device=self._flat.device)
            self._bf16_version = -1
        sig = self._weights_signature()
        if self._bf16_version != sig:
            nat.check(nat.lib().lako_cast_f32_to_bf16(nat.ptr(self._flat), nat.ptr(self._bf16),
                                                      ctypes.c_size_t(self._total), nat.stream_ptr()), "cast")
            self._bf16_version = sig

    def _weights_signature(self):
        # in-place updates (optimizer.step, copy_) bump the version counters of the Parameters / the flat buffer
        return self._flat._version + sum(p._version for p in self._params.values())

    def mark_bf16_fresh(self):
        """Called by the fused optimizer, which re-casts the weights itself."""
        if self._bf16 is None:
            self._bf16 = torch.empty(self._total, dtype=torch.bfloat16, device=self._flat.device)
        self._bf16_version = self._weights_signature()

    def _refresh_tf32(self):
        """(hi, lo) TF32 split of the fp32 masters for precision='fp32' (re-made when the weights change)."""
        sig = self._weights_signature()
        if self._tf32 is None or self._tf32_version != sig:
            if self._master_sync is not None:
                self._master_sync()
            if self._tf32 is None:
                self._tf32 = (torch.empty_like(self._flat), torch.empty_like(self._flat))
            nat.check(nat.lib().lako_split_tf32(nat.ptr(self._flat), nat.ptr(self._tf32[0]), nat.ptr(self._tf32[1]),
                                                ctypes.c_size_t(self._total), nat.stream_ptr()), "lako_split_tf32")
            self._tf32_version = self._weights_signature()

    def _model_struct(self, with_grad, precision=0, wait=True):
        if wait:
            self._wait_weights()
        self._refresh_bf16()
        hi = lo = 0
        if precision == 1:
            self._refresh_tf32()
            hi, lo = self._tf32[0].data_ptr(), self._tf32[1].data_ptr()
        return _Model(self._dims, self._flat.data_ptr(), self._bf16.data_ptr(),
                      self._flat_grad.data_ptr() if with_grad else 0,
                      self._enc_buckets.data_ptr(), self._dec_buckets.data_ptr(), BUCKET_CENTER, hi, lo)

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key[0])
        if ws is None or ws[0] != key or ws[1].numel() < nbytes:
            self._ws.pop(key[0], None)
            ws = (key, torch.empty(nbytes, dtype=torch.uint8, device=self._flat.device))
            if os.environ.get("LAKO_POISON_WORKSPACE") == "1":   # test support: every fp32 / bf16 word a NaN, so that a
                ws[1].fill_(0xFF)                                # kernel reading what nobody wrote shows up in the results
            self._ws[key[0]] = ws
        return ws[1]

    def _packing(self, mask, lengths=None):
        """(cu_seqlens int32 device tensor [B*N+1], total valid tokens) if every passage mask is a prefix mask,
        else (None, 0).  Costs one small D2H copy (B*N+1 ints) per call -- unless the caller passes `lengths`, the
        HOST tensor / list [B][N] of valid-prefix lengths (a by-product of src/data.py's Collator, which builds the
        masks on the host): then nothing is read back and the host keeps running ahead of the device; the claim
        "mask[b,n,:] == arange(L) < lengths[b,n]" is checked on the device (asynchronous device-side assert)."""
        if not self.unpad:
            return None, 0
        B, N, L = mask.shape
        if lengths is not None:
            lens_h = torch.as_tensor(lengths, dtype=torch.int32).reshape(-1).cpu()
            if lens_h.numel() != B * N:
                raise nat.LakoError(f"passage_lengths must have {B * N} entries, got {lens_h.numel()}")
            cu = torch.zeros(B * N + 1, dtype=torch.int32)
            cu[1:] = torch.cumsum(lens_h, 0)
            total = int(cu[-1])
            if total == 0 or int(lens_h.min()) < 0 or int(lens_h.max()) > L:
                raise nat.LakoError("passage_lengths out of range")
            cu_d = cu.to(mask.device, non_blocking=True)
            lens_d = (cu_d[1:] - cu_d[:-1]).reshape(B, N, 1)
            torch._assert_async((mask == (torch.arange(L, device=mask.device)[None, None, :] < lens_d)).all())
            return cu_d, total
        lens = mask.sum(-1, dtype=torch.int32)
        prefix = (mask == (torch.arange(L, device=mask.device)[None, None, :] < lens[..., None])).all()
        info = torch.cat([lens.reshape(-1), prefix.reshape(1).to(torch.int32)]).cpu()
        if int(info[-1]) == 0:
            return None, 0
        cu = torch.zeros(B * N + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(info[:-1], 0)
        total = int(cu[-1])
        if total == 0:
            return None, 0
        return cu.to(mask.device), total

    def set_precision(self, precision):
        """"bf16" (default): bf16 tensor-core operands with fp32 accumulation everywhere.  "fp32": the evaluation /
        score-extraction path (forward without grad, generate, get_crossattention_scores) runs 3xTF32 split tcgen05
        GEMMs on the fp32 master weights and fp32 attention kernels -- the reference itself is fp32-only
        (train_reader.py:67, no autocast anywhere); north_star's 1e-4 logit tolerance is stated for this mode."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision

    def _precision_code(self, override=None):
        p = self.precision if override is None else override
        if p not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        return 1 if p == "fp32" else 0

    def set_dropout_seed(self, seed):
        """Seed of the dropout masks (defaults to torch.initial_seed() at the first training forward; use a
        different seed per data-parallel rank, as train_reader.py:39 does with torch.manual_seed(rank + seed))."""
        self._drop_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._drop_step = 0

    # ------------------------------------------------------------------ forward / backward
    _IGNORED_KWARGS = ("return_dict", "use_cache", "output_attentions", "output_hidden_states")

    def forward(self, input_ids=None, attention_mask=None, labels=None, decoder_input_ids=None,
                global_target_count=None, passage_lengths=None, precision=None, **kwargs):
        """model.py:39-51.  With `labels`: (loss, logits).  With `decoder_input_ids` and no labels (HF's other calling
        convention): (logits,) for exactly those decoder inputs, no loss and no backward."""
        self._require_cuda()
        bad = [k for k, v in kwargs.items() if k not in self._IGNORED_KWARGS and v is not None]
        if bad:
            raise TypeError(f"FiDT5.forward: unsupported arguments {bad} (supported: input_ids, attention_mask, labels, "
                            "decoder_input_ids; return_dict/use_cache/output_* are accepted and have no effect)")
        if kwargs.get("output_attentions") or kwargs.get("output_hidden_states"):
            raise TypeError("FiDT5.forward: attention maps / hidden states are never materialised by the fused kernels")
        dec_given = labels is None
        if dec_given:
            if decoder_input_ids is None:
                raise nat.LakoError("FiDT5.forward needs labels (model(..., labels=...)[0], train_reader.py:67-71) or "
                                    "decoder_input_ids; use generate() for inference")
            labels = decoder_input_ids
        if input_ids.dim() == 3:
            self.encoder.n_passages = input_ids.size(1)
        B = input_ids.size(0)
        N = self.encoder.n_passages
        ids = input_ids.reshape(B, N, -1).contiguous()
        L = ids.size(2)
        if attention_mask is None:
            attention_mask = torch.ones_like(ids, dtype=torch.bool)
        mask = attention_mask.reshape(B, N, L).to(torch.bool).contiguous()
        labels = labels.contiguous()
        T = labels.size(1)
        train = torch.is_grad_enabled() and not dec_given
        mode = 1 if train else 0
        prec = self._precision_code(precision)
        if train and prec != 0:
            raise nat.LakoError("precision='fp32' is an evaluation mode (forward / generate / score extraction); "
                                "training runs bf16 operands with fp32 accumulation")
        if train:
            self._ensure_grad()
        L_ = nat.lib()
        L_.lako_fidt5_workspace_bytes.restype = ctypes.c_size_t
        recompute = int(train and self.use_checkpoint)
        nbytes = L_.lako_fidt5_workspace_bytes(ctypes.byref(self._dims), B, N, L, T,
                                               2 if prec == 1 else (3 if recompute else mode))
        ws = self._workspace(("fwd", B, N, L, T, mode, prec, recompute), nbytes)
        self._ws_gen += 1   # the workspace now holds THIS forward's activations (see _run_backward)
        logits = torch.empty(B, T, self.config.vocab_size, dtype=torch.float32, device=ids.device)
        stats = torch.zeros(2, dtype=torch.float32, device=ids.device)
        count_is_set = 0
        if global_target_count is not None:
            stats[1] = global_target_count
            count_is_set = 1
        drop_p = float(self.config.dropout_rate) if (train and self.training) else 0.0
        if drop_p > 0.0:
            if self._drop_seed is None:
                self._drop_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
            self._drop_step += 1
        drop = (drop_p, self._drop_step & 0xFFFFFFFF, self._drop_seed or 0)
        cu, packed = self._packing(mask, passage_lengths)
        pack = (cu.data_ptr() if cu is not None else 0, packed)
        # a training forward waits for an overlapped weight all-gather range by range inside the C call
        wev = self._pending_wev if (train and prec == 0) else None
        m = self._model_struct(with_grad=train, precision=prec, wait=wev is None)
        bt = _Batch(ids.data_ptr(), mask.data_ptr(), labels.data_ptr(), B, N, L, T, ws.data_ptr(), ws.numel(),
                    logits.data_ptr(), stats.data_ptr(), count_is_set, *drop, *pack, 0, int(dec_given), prec, recompute,
                    ctypes.cast(wev[1], ctypes.c_void_p).value if wev is not None else 0)
        nat.check(L_.lako_fidt5_forward(ctypes.byref(m), ctypes.byref(bt), mode, nat.stream_ptr()), "lako_fidt5_forward")
        if wev is not None:
            self._pending_wev = None   # the stream has now waited on every range
        if dec_given:
            return ReaderOutput(None, logits)
        loss = stats[0] / torch.clamp(stats[1], min=1.0)
        if train:
            keep = {"t": (ids, mask, labels, ws, logits, stats, (B, N, L, T), drop, cu, pack), "gen": self._ws_gen,
                    "done": False, "recompute": recompute}
            loss, logits = _LossFn.apply(self._anchor, self, keep, loss, logits)
        return ReaderOutput(loss, logits)

    def enable_grad_stage_events(self):
        """Create the 6 stage events + flat ranges used to overlap the gradient all-reduce with backward."""
        if self._stage_events is None:
            evs = [torch.cuda.Event() for _ in range(6)]
            for e in evs:
                e.record()          # forces creation of the underlying cudaEvent_t
            ranges = (ctypes.c_size_t * 14)()
            nat.check(nat.lib().lako_fidt5_grad_stages(ctypes.byref(self._dims), ranges), "lako_fidt5_grad_stages")
            self._stage_events = evs
            self._stage_handles = (ctypes.c_void_p * 6)(*[e.cuda_event for e in evs])
            self._stage_ranges = [(int(ranges[2 * i]), int(ranges[2 * i + 1])) for i in range(7)]
        return self._stage_events, self._stage_ranges

    def _run_backward(self, keep, grad_loss):
        # The activation workspace is one buffer per shape, reused by every forward: a later grad-enabled forward with the
        # same shapes overwrites what this graph saved, and backward scales the stored dlogits in place.
        if keep["gen"] != self._ws_gen:
            raise nat.LakoError("loss.backward(): another forward with the same shapes ran after this loss was computed and "
                                "overwrote its saved activations; call backward() before the next forward (gradient "
                                "accumulation: forward -> backward per micro-batch, the gradients add up in .grad)")
        if keep["done"]:
            raise nat.LakoError("loss.backward() called twice on the same graph (retain_graph is not supported: backward "
                                "consumes the saved activations)")
        keep["done"] = True
        ids, mask, labels, ws, logits, stats, (B, N, L, T), drop, _cu, pack = keep["t"]
        L_ = nat.lib()
        m = self._model_struct(with_grad=True)
        stage = ctypes.cast(self._stage_handles, ctypes.c_void_p).value if self._stage_events is not None else 0
        bt = _Batch(ids.data_ptr(), mask.data_ptr(), labels.data_ptr(), B, N, L, T, ws.data_ptr(), ws.numel(),
                    logits.data_ptr(), stats.data_ptr(), 1, *drop, *pack, stage, 0, 0, keep["recompute"])
        gl = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        nat.check(L_.lako_fidt5_backward(ctypes.byref(m), ctypes.byref(bt), nat.ptr(gl), nat.stream_ptr()),
                  "lako_fidt5_backward")

    # ------------------------------------------------------------------ generate + scores
    @torch.no_grad()
    def generate(self, input_ids, attention_mask, max_length, min_length=0, precision=None):
        self._require_cuda()
        B, N, L = input_ids.shape
        self.encoder.n_passages = N
        ids = input_ids.contiguous()
        mask = attention_mask.to(torch.bool).contiguous()
        L_ = nat.lib()
        L_.lako_fidt5_generate_workspace_bytes.restype = ctypes.c_size_t
        prec = self._precision_code(precision)
        nbytes = L_.lako_fidt5_generate_workspace_bytes(ctypes.byref(self._dims), B, N, L, max_length, prec)
        ws = self._workspace(("gen", B, N, L, max_length, prec), nbytes)
        out = torch.empty(B, max_length, dtype=torch.long, device=ids.device)
        tf = tl = None
        if self._capture and self._tok_first is None:
            # reference semantics: scores are stored on the first decoder call after a reset (model.py:327-329)
            self._tok_first = torch.zeros(B, N * L, dtype=torch.float32, device=ids.device)
            self._tok_last = torch.zeros(B, N * L, dtype=torch.float32, device=ids.device)
            tf, tl = self._tok_first, self._tok_last
        cu, packed = self._packing(mask)
        m = self._model_struct(with_grad=False, precision=prec)
        bt = _Batch(ids.data_ptr(), mask.data_ptr(), 0, B, N, L, 1, ws.data_ptr(), ws.numel(), 0, 0, 0, 0.0, 0, 0,
                    cu.data_ptr() if cu is not None else 0, packed, 0, 0, prec)
        out_len = ctypes.c_int(0)
        nat.check(L_.lako_fidt5_generate(ctypes.byref(m), ctypes.byref(bt), max_length, int(min_length), nat.ptr(out),
                                         ctypes.byref(out_len), nat.ptr(tf), nat.ptr(tl), nat.stream_ptr()),
                  "lako_fidt5_generate")
        if tf is not None and self._xattn_relbias is not None:
            # 3.0.2 semantics (see load_state_dict): every stored score carries + table[0, h]; after the masked sum over
            # heads and layers (model.py:161-167) that is a constant per valid token and selected layer
            c = self._xattn_relbias[0].sum()
            nl = self.config.num_layers
            valid = mask.reshape(B, N * L).to(torch.float32)
            tf.add_(valid * (c * ((nl + 1) // 2)))
            tl.add_(valid * (c * (nl - (nl + 1) // 2)))
        return out[:, :out_len.value]

    def overwrite_forward_crossattention(self):
        """model.py:206-213: from now on the first decoder step after each reset records its scores."""
        self._capture = True

    def reset_score_storage(self):
        """model.py:92-98."""
        self._tok_first = self._tok_last = None

    def token_scores(self, last_half):
        """(B, N*L) fp32: masked step-0 pre-softmax cross-attention scores summed over heads and the selected
        decoder layers -- what model.py:152-167 reduces score_storage to -- and the divisor of model.py:202."""
        if self._tok_first is None:
            raise nat.LakoError("no scores stored: call overwrite_forward_crossattention() and generate() first")
        nl, H = self.config.num_layers, self.config.num_heads
        if last_half:
            return self._tok_last, (nl - (nl + 1) // 2) * H
        return self._tok_first + self._tok_last, nl * H

    SCORE_STYLES = {"mean": 0, "max": 1, "21mean": 2}

    def get_crossattention_scores(self, opt, context_ids, tokenizer, context_mask, softmax=False):
        """model.py:143-204 (same argument list, same float64 CPU (B, opt.n_context) result) on the fused token-score
        tensor.  The per-example fact segmentation of passage 1 (split on token id 5 from index 2, mean / max / 21mean,
        -5 fillers, / (layers * heads)) runs in ONE kernel (lako_fact_scores) instead of the reference's Python loop over
        a (B, N, L) device->host copy: what crosses the bus is the (B, L) ids of passage 1 in and the (B, n_context)
        result out.  softmax=True additionally applies test_reader.py:76's torch.softmax(scores, dim=-1)."""
        assert opt.stream == 2
        if opt.attention_score_style not in self.SCORE_STYLES:
            raise ValueError(f"attention_score_style must be one of {sorted(self.SCORE_STYLES)}")
        bsz, n_passages, L = context_mask.shape
        assert n_passages >= 2, "the score pass needs the fact passage (stream 2: passage 1)"
        if self._tok_first is None:
            raise nat.LakoError("no scores stored: call overwrite_forward_crossattention() and generate() first")
        nl, H = self.config.num_layers, self.config.num_heads
        if opt.use_last_half_layer_attention == "yes":
            a, b2, div = self._tok_last, None, (nl - (nl + 1) // 2) * H
        else:
            a, b2, div = self._tok_first, self._tok_last, nl * H
        dev = a.device
        ids1 = context_ids[:, 1].to(dev, non_blocking=True).contiguous()
        out = torch.empty(bsz, int(opt.n_context), dtype=torch.float64, device=dev)
        # torch.chunk(scores, 2, dim=1)[1][:, 0]: the first passage of the second half (model.py:166,175)
        nat.check(nat.lib().lako_fact_scores(nat.ptr(a), nat.ptr(b2), nat.ptr(ids1), ctypes.c_int(L), nat.ptr(out),
                                             ctypes.c_int(bsz), ctypes.c_int(n_passages), ctypes.c_int(L),
                                             ctypes.c_int((n_passages + 1) // 2), ctypes.c_int(int(opt.n_context)),
                                             ctypes.c_int(self.SCORE_STYLES[opt.attention_score_style]),
                                             ctypes.c_double(float(div)), ctypes.c_int(int(bool(softmax))),
                                             nat.stream_ptr()), "lako_fact_scores")
        return out.cpu()
