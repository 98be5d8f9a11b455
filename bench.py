#!/usr/bin/env python
"""bench.py -- FiDT5 reader throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4|cfg2|cfg3|cfg5]

Default workload = the north-star configuration (BASELINE.json configs[3] / north_star "Target"): FiDT5 t5-large
training, n_context = 20 passages x text_maxlength = 250 tokens, 8 questions per GPU, T = 10 target tokens, bf16
operands with fp32 accumulation, dropout 0.1.  A "step" = one pass of the hot path over one synthetic batch: FiDT5
forward (loss) + backward + gradient reduce-scatter/all-gather (N>1) + clip + AdamW, i.e. train_reader.py:67-79 of the
reference.  For N>1 every rank takes its own 8 questions (question-sharded data parallelism, weak scaling).
`--workload cfg2` = configs[1] (t5-base, B=8, N=10, L=200).  `--workload cfg3` / `cfg5` time the evaluation /
attention-signal pass instead (test_reader.py:60-76: greedy generate + step-0 cross-attention score capture +
get_crossattention_scores) and report questions/s.

Prints ONE JSON line (rank 0).  `value` is samples/s with inputs already resident in HBM; `e2e` is the same metric
through the public API with pinned HOST buffers (H2D of ids/mask/labels, the host-side passage-length computation and a
D2H read of the loss inside the timed region).  `roofline` describes the dominant kernel (the tcgen05 CTA-pair GEMM)
from CUDA-event timings taken live in this process; `cpu_baseline` times the oracle (CPU restatement of the
reference) on a bounded sample on this box's host cores; `gpu_eager_baseline` times the same oracle in stock PyTorch
eager on this GPU (fp32 as the reference runs it, and under bf16 autocast) -- the on-box library path the kernels here
replace; `generate` carries the evaluation-pass numbers at the same model size.

--impl reference: the reference's own PyTorch CPU path for the same step, restated in oracle/fidt5_oracle.py (the
reference is Python over an un-vendored transformers and cannot travel to the GPU box), timed on the
host cores on a bounded sample (1 question / step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (t5 size, per-GPU questions B, passages N, text_maxlength L, target length T, kind)
    "cfg2": ("base", 8, 10, 200, 10, "train"),
    "cfg4": ("large", 8, 20, 250, 10, "train"),
    "cfg3": ("large", 8, 20, 250, 10, "eval"),
    "cfg5": ("large", 2, 50, 250, 10, "eval"),
}
METRIC = "fidt5_train_samples_per_sec"
METRIC_EVAL = "fidt5_generate_questions_per_sec"
EMB_STD = 0.25
DIMS = {"base": (768, 768, 3072, 12, 32128), "large": (1024, 1024, 4096, 24, 32128)}


def fwd_flops(cfgname, B, N, L, T):
    """SURVEY.md 8(d) algorithmic FLOPs of one forward on the padded shapes; train = 3x."""
    d, inner, ff, nl, V = DIMS[cfgname]
    enc = nl * (8 * N * L * d * inner + 4 * N * L * L * inner + 4 * N * L * d * ff)
    dec = nl * (12 * T * d * inner + 4 * T * T * inner + 4 * N * L * d * inner + 4 * T * N * L * inner + 4 * T * d * ff)
    return B * (enc + dec + 2 * T * d * V)


def fwd_flops_executed(cfgname, lens, T):
    """Same formula on the rows the packed execution actually processes: `lens` is the [B][N] tensor of valid
    passage lengths (pad tokens are neither projected nor attended)."""
    d, inner, ff, nl, V = DIMS[cfgname]
    B = lens.shape[0]
    tok = float(lens.sum())
    sq = float((lens.double() ** 2).sum())
    enc = nl * (tok * (8 * d * inner + 4 * d * ff) + 4 * sq * inner)
    dec = nl * (B * (12 * T * d * inner + 4 * T * T * inner + 4 * T * d * ff) + tok * (4 * d * inner + 4 * T * inner))
    return enc + dec + B * 2 * T * d * V


def decode_step_bytes(cfgname, mem_tokens):
    """SURVEY 8(d) algorithmic HBM bytes of ONE cached decode step: the cross-attention K/V of every decoder layer for
    the `mem_tokens` memory tokens of the batch + the decoder weights and the tied LM head once."""
    d, inner, ff, nl, V = DIMS[cfgname]
    kv = nl * 2 * mem_tokens * inner * 2
    w = (nl * (6 * d * inner + 2 * d * ff) + V * d) * 2
    return kv + w


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.window = None  # (t0, t1) host times of the timed region; samples outside it are dropped

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        """NVML polled every ~5 ms; same fields as the nvidia-smi query."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        while not self.stop_flag:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
            r = get_reasons(h)
            row = [str(sm), str(mx), f"{pw:.0f}"] + ["Not Active"] * 4
            for bit, col in bits:
                if r & bit:
                    row[col] = "Active"
            self.samples.append((time.perf_counter(), row))
            time.sleep(0.004)

    def _run_smi(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append((time.perf_counter(), [x.strip() for x in out.split(",")]))
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, pw, mx, reasons = [], [], 0, set()
        rows = [r for t, r in self.samples if self.window is None or self.window[0] <= t <= self.window[1]]
        for s in rows:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                pw.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "sm_min_mhz": sm[0] if sm else None, "power_w": pw[len(pw) // 2] if pw else None}


def oracle_step_time(size, B, N, L, T, steps, warmup, device="cpu", autocast=False, seed=0):
    """One train step (fwd + bwd + clip + HF-AdamW) of the oracle -- the restated reference path in stock PyTorch
    eager -- on `device`, batch of B questions.  -> (seconds per step, host threads)."""
    import torch
    from oracle import fidt5_oracle as O
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    else:
        torch.backends.cuda.matmul.allow_tf32 = False   # the reference runs plain fp32 (no autocast / tf32 anywhere)
    cfg = {"base": O.T5_BASE, "large": O.T5_LARGE}[size]
    sd = O.init_state_dict(cfg, seed=seed, emb_std=EMB_STD)
    if device != "cpu":
        sd = O.sd_to(sd, device)
    params = O.unique_params(sd)
    for p in params.values():
        p.requires_grad_(True)
    opt = O.HFAdamW(list(params.values()), lr=1e-4, weight_decay=1e-4)
    ids, mask, labels = [t.to(device) for t in O.synth_batch(B, N, L, T, cfg.vocab, seed=1234)]
    times = []
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                O.train_step(sd, cfg, opt, ids, mask, labels, lr_scale=1.0, clip=1.0, dropout=0.1)
        else:
            O.train_step(sd, cfg, opt, ids, mask, labels, lr_scale=1.0, clip=1.0, dropout=0.1)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size, B, N, L, T, kind = WORKLOADS[args.workload]
    B = args.batch or B
    N = args.passages or N
    L = args.text_maxlength or L
    steps = max(1, args.steps)
    warmup = max(1, min(args.warmup, 2))
    # bounded: cap the number of timed CPU steps so the arm ends within minutes (one t5-large step of ONE question
    # takes ~15 s on 16 cores)
    steps = min(steps, 5 if size == "base" else 3)
    if size == "large":
        warmup = 1
    sec, cores = oracle_step_time(size, 1, N, L, T, steps, warmup)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: FiDT5 t5-{size} train step, n_passages={N}, text_maxlength={L}, "
                               f"T={T}, dropout=0.1 (bounded sample: 1 question per step instead of {B})",
                   "parallelism": "cpu"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} train steps of 1 question ({N}x{L} tokens) after {warmup} warm-up"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def synth_host_batches(B, N, L, T, rank, n_batches=4, fact_format=False):
    """Seeded synthetic batches in the Collator's layout (src/data.py:148), pinned host tensors."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    host = []
    for _ in range(n_batches):
        ids = torch.randint(11, 32000, (B, N, L), generator=g)
        lens = torch.randint(int(0.6 * L), L + 1, (B, N), generator=g)
        mask = torch.arange(L)[None, None, :] < lens[:, :, None]
        ids = ids * mask
        if fact_format and N >= 2:   # passage 1 = "fact: s1 . s2 . ...": ids 6,10 then id 5 every 8-20 tokens
            for b in range(B):
                ids[b, 1, 0], ids[b, 1, 1] = 6, 10
                p = 2
                while True:
                    p += int(torch.randint(8, 21, (1,), generator=g))
                    if p >= int(lens[b, 1]):
                        break
                    ids[b, 1, p] = 5
        tl = torch.randint(2, T + 1, (B,), generator=g)
        labels = torch.randint(11, 32000, (B, T), generator=g)
        tpos = torch.arange(T)[None, :]
        labels = torch.where(tpos == (tl[:, None] - 1), torch.ones_like(labels), labels)
        labels = labels.masked_fill(tpos >= tl[:, None], -100)
        host.append((ids.pin_memory(), mask.pin_memory(), labels.pin_memory()))
    return host


def time_generate(model, ids, mask, max_length, reps=3, with_scores=True, opt=None):
    """ms per evaluation-pass call (test_reader.py:63-76): reset, greedy generate forced to `max_length` columns
    (bench-only min_length = max_length: random-init weights emit EOS at once), score extraction, D2H of answers."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def one():
        model.reset_score_storage()
        out = model.generate(input_ids=ids, attention_mask=mask, max_length=max_length, min_length=max_length)
        sc = None
        if with_scores:
            sc = model.get_crossattention_scores(opt, ids, None, mask)
        return out.cpu(), sc

    out, _ = one()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out, sc = one()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, int(out.shape[1])


def eval_pass_numbers(model, size, B, N, L, dev, rank, peaks, lengths=(10, 50)):
    """Evaluation / attention-signal pass at (B, N, L): questions/s per max_length, ms per decoded token and the
    achieved HBM GB/s of one cached decode step against SURVEY 8(d)'s bytes (cross-KV of the batch + decoder
    weights), from the difference between the two max_length timings (the encoder cancels)."""
    import types
    import torch
    host = synth_host_batches(B, N, L, 4, rank, n_batches=1, fact_format=True)
    ids, mask = host[0][0].to(dev), host[0][1].to(dev)
    opt = types.SimpleNamespace(stream=2, n_context=10, attention_score_style="21mean",
                                use_last_half_layer_attention="yes")
    model.eval()
    model.overwrite_forward_crossattention()
    res = {"B": B, "n_passages": N, "text_maxlength": L, "score_capture": "21mean, last-half layers"}
    ms = {}
    for ml in lengths:
        ms[ml], cols = time_generate(model, ids, mask, ml, opt=opt)
        res[f"max_length_{ml}"] = {"questions_per_s": round(B / (ms[ml] * 1e-3), 2), "ms_per_call": round(ms[ml], 3),
                                   "columns_generated": cols}
    if len(lengths) >= 2:
        a, b = lengths[0], lengths[-1]
        step_ms = (ms[b] - ms[a]) / (b - a)
        mem_tokens = int(mask.sum().item())
        nbytes = decode_step_bytes(size, mem_tokens)
        res["decode_step"] = {"ms_per_token": round(step_ms, 4), "algorithmic_bytes": nbytes,
                              "achieved_GBs": round(nbytes / (step_ms * 1e-3) / 1e9, 1),
                              "frac_hbm_peak": round(nbytes / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                              "encoder_plus_first_step_ms": round(ms[a] - (a - 2) * step_ms, 3)}
    model.train()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lako_b200", choices=["lako_b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU questions (default: the workload's own; cfg3 is also "
                    "quoted at B = 1 and 32)")
    ap.add_argument("--passages", type=int, default=0, help="passages per question (default: the workload's own; cfg5 is "
                    "also quoted at N = 100)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="evaluation workloads only: fp32 = 3xTF32 GEMMs + fp32 attention (the mode that meets 1e-4 / 1e-2)")
    ap.add_argument("--text-maxlength", type=int, default=0, help="tokens per passage (default: the workload's own; above "
                    "256 the encoder attention takes the tiled mma.sync path)")
    ap.add_argument("--use-checkpoint", action="store_true",
                    help="training workloads: model.set_checkpoint(True) (the reference's --use_checkpoint): encoder "
                         "activations recomputed in backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip generate / gpu_eager_baseline probes")
    args = ap.parse_args()
    if args.impl == "reference":
        if WORKLOADS[args.workload][5] != "train":
            args.workload = "cfg4"
        run_reference(args)
        return

    import ctypes
    import torch
    import torch.distributed as dist
    from lako_b200 import _native as nat
    from lako_b200.model import FiDT5, FiDT5Config
    from lako_b200.optim import FusedAdamW, train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        # NCCL prints its version banner on stdout at the first collective; the driver wants exactly one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                    timeout=datetime.timedelta(seconds=180))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local_rank)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)

    size, B, N, L, T, kind = WORKLOADS[args.workload]
    B = args.batch or B
    N = args.passages or N
    L = args.text_maxlength or L
    conf = FiDT5Config.t5(size)
    model = FiDT5(conf, emb_std=EMB_STD, seed=0).cuda()
    peaks, which = measured_peaks()
    lib = nat.lib()
    lib.lako_launch_count.restype = ctypes.c_ulonglong

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if kind == "eval":
        run_eval_workload(args, model, size, B, N, L, dev, rank, world, peaks, which, lib, barrier, warmup, steps)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    model.train()
    model.set_checkpoint(args.use_checkpoint)
    opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4, clip=1.0, shard=True)
    n_batches = 4  # rotate distinct synthetic batches
    host = synth_host_batches(B, N, L, T, rank, n_batches)
    # valid-prefix lengths: a host by-product of the Collator (it builds the masks on the host); passing them lets the
    # step run without reading anything back from the device.  The resident-input loop uses precomputed lengths; the
    # e2e loop below computes them from the host mask INSIDE its timed region.
    host_lens = [b.sum(-1, dtype=torch.int32) for _, b, _ in host]
    resident = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c in host]
    valid_frac = float(sum(b.float().mean().item() for _, b, _ in host) / len(host))

    # the sampler thread starts before the warm-up (NVML initialisation takes a while); only its samples taken
    # inside the timed region are reported
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- warm-up
    for i in range(warmup):
        train_step(model, opt, *resident[i % n_batches], passage_lengths=host_lens[i % n_batches])
    barrier()

    # ---- timed region: K steps, inputs resident in HBM
    launches0 = lib.lako_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = bool(os.environ.get("LAKO_PROFILER_RANGE"))  # ncu --profile-from-start off: capture the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    t_win0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = train_step(model, opt, *resident[i % n_batches], passage_lengths=host_lens[i % n_batches])
    e1.record()
    barrier()
    t_win1 = time.perf_counter()
    if prof_range:
        torch.cuda.profiler.stop()
    launches = lib.lako_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if sampler:
        sampler.stop_flag = True
        sampler.window = (t_win0, t_win1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * steps / (ms * 1e-3)
    final_loss = loss.item()

    # ---- e2e: same metric through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        barrier()
        e0.record()
        for i in range(steps):
            a, b, c = host[i % n_batches]
            lens_h = b.sum(-1, dtype=torch.int32)  # host-side passage lengths, inside the timed region
            ids_d = a.to(dev, non_blocking=True)
            mask_d = b.to(dev, non_blocking=True)
            lab_d = c.to(dev, non_blocking=True)
            l_ = train_step(model, opt, ids_d, mask_d, lab_d, passage_lengths=lens_h)
            _ = l_.item()  # D2H of the step's loss (train_reader.py:82)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a, b, c = host[0]
        e2e = {"value": world * B * steps / (t.item() * 1e-3), "unit": "samples/s",
               "h2d_bytes_per_step": a.numel() * 8 + b.numel() + c.numel() * 8 + (B * N + 1) * 4,
               "d2h_bytes_per_step": 4,
               "host_work_in_region": "passage lengths from the host mask, cu_seqlens, H2D copies, loss.item()"}

    # ---- per-kernel timing (CUDA events on the launching stream, same process, right after the timed region)
    roofline = None
    kernel_share = None
    clocks = sampler.summary() if sampler else None
    nprof = 2
    if rank == 0:
        lib.lako_prof_enable(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pe0.record()
    for i in range(nprof):  # every rank steps (the step contains collectives); only rank 0 records events
        train_step(model, opt, *resident[i % n_batches], passage_lengths=host_lens[i % n_batches])
    pe1.record()
    torch.cuda.synchronize()
    if rank == 0:
        ncat = 9
        msa = (ctypes.c_double * ncat)()
        wka = (ctypes.c_double * ncat)()
        cna = (ctypes.c_longlong * ncat)()
        lib.lako_prof_collect(msa, wka, cna)
        lib.lako_prof_enable(0)
        step_ms_prof = pe0.elapsed_time(pe1) / nprof
        names = ["gemm_fwd_large", "enc_attn_fwd", "enc_attn_bwd", "rmsnorm", "cross_attn", "adamw_gradnorm",
                 "gemm_dgrad_large", "gemm_wgrad_large", "gemm_decoder_small"]
        kernel_share = {}
        for i, nme in enumerate(names):
            if cna[i] == 0:
                continue
            per_step_ms = msa[i] / nprof
            ent = {"ms_per_step": round(per_step_ms, 4), "launches_per_step": cna[i] // nprof,
                   "share_of_step": round(per_step_ms / step_ms_prof, 4)}
            if nme in ("rmsnorm", "adamw_gradnorm"):
                ent["achieved_GBs"] = round(wka[i] / (msa[i] * 1e-3) / 1e9, 1)
                ent["frac_hbm_peak"] = round(ent["achieved_GBs"] / peaks["hbm_gbs"], 4)
            else:
                ent["achieved_TFLOPs"] = round(wka[i] / (msa[i] * 1e-3) / 1e12, 1)
            if nme == "cross_attn":
                # HBM-bound: K and V of one decoder layer per forward launch; K, V read + dK, dV written per backward
                d_, inner_, ff_, nl_, V_ = DIMS[size]
                tok = float(sum(float(x.sum()) for x in host_lens) / len(host_lens))
                nbytes = nl_ * tok * inner_ * 2 * 2 * (1 + 2)
                ent["achieved_GBs"] = round(nbytes / (per_step_ms * 1e-3) / 1e9, 1)
                ent["frac_hbm_peak"] = round(ent["achieved_GBs"] / peaks["hbm_gbs"], 4)
            kernel_share[nme] = ent
        # dominant kernel: the tcgen05 GEMM on the token-major operands (M > 128 rows: encoder Linear fwd/dgrad/wgrad and
        # the fused cross-K/V projection, run by the CTA-pair kernel).  The decoder's <= 80-row GEMMs are latency bound
        # and are reported separately in "kernels" (gemm_decoder_small), not against the tensor roofline.
        gi = (0, 6, 7)
        g_ms, g_wk, g_n = sum(msa[i] for i in gi), sum(wka[i] for i in gi), sum(cna[i] for i in gi)
        gemm_tf = g_wk / (g_ms * 1e-3) / 1e12
        # denominator: the burst cuBLAS figure when the SM clock held >= 95 % of its maximum through the timed region
        # with no power cap, the sustained figure when the run was power-capped (cfg4 draws the full 1000 W)
        # (decided by the sampled clock, not by the reason flag alone: at t5-base the flag shows up in single samples
        # while the SM clock never leaves its maximum)
        capped = bool(clocks) and bool(clocks["sm_mhz"] and clocks["sm_max_mhz"] and
                                       clocks["sm_mhz"] < 0.95 * clocks["sm_max_mhz"])
        sustained = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        peak = sustained if capped else peaks["bf16_tflops"]
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tj.get(args.workload, {}).get("gemm_dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2), all launches with M > 128",
                    "achieved": round(gemm_tf, 1),
                    "peak": peak,
                    "peak_source": f"{which} cuBLAS bf16, " + ("sustained (median SM clock below 95 % of max during the timed region)"
                                                                if capped else "burst (median SM clock >= 95 % of max)"),
                    "unit": "TFLOP/s", "frac": round(gemm_tf / peak, 4),
                    "frac_of_burst": round(gemm_tf / peaks["bf16_tflops"], 4),
                    "frac_of_sustained": round(gemm_tf / sustained, 4), "traffic": traffic,
                    "launches_per_step": g_n // nprof, "avg_launch_ms": round(g_ms / max(1, g_n), 5),
                    "share_of_step": round(g_ms / nprof / step_ms_prof, 4)}

    if world > 1:
        dist.barrier()

    # ---- supplementary (rank 0, outside every timed region of the training metric)
    gen = None
    eager = None
    if rank == 0 and not args.no_e2e and not args.no_extras:
        try:   # evaluation / attention-signal pass at the same model size (cfg3 shape for t5-large)
            gen = eval_pass_numbers(model, size, B, N, L, dev, rank, peaks)
        except Exception as e:  # never let a supplementary probe break the contract line
            gen = {"error": str(e)[:200]}
            model.train()
    if rank == 0 and world == 1 and not args.no_extras:
        # the reference's own library path on THIS GPU: the oracle in stock PyTorch eager (cuBLAS GEMMs, unfused
        # softmax / norms), same step; bounded to EB questions per step, reported per sample
        try:
            EB = 2
            torch.cuda.empty_cache()
            s32, _ = oracle_step_time(size, EB, N, L, T, steps=2, warmup=1, device=dev)
            torch.cuda.empty_cache()
            s16, _ = oracle_step_time(size, EB, N, L, T, steps=2, warmup=1, device=dev, autocast=True)
            torch.cuda.empty_cache()
            eager = {"fp32_samples_per_s": round(EB / s32, 2), "bf16_autocast_samples_per_s": round(EB / s16, 2),
                     "unit": "samples/s", "sample": f"2 oracle train steps of {EB} questions after 1 warm-up, torch "
                     f"{torch.__version__} eager on the same GPU (TF32 off in the fp32 run)", "kind": "port"}
        except Exception as e:
            eager = {"error": str(e)[:200]}

    # ---- CPU baseline (rank 0, N=1 only): oracle on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores = oracle_step_time(size, 1, N, L, T, steps=1 if size == "large" else 2, warmup=1)
        cpu_baseline = {"value": 1.0 / sec, "unit": "samples/s", "cores": cores, "kind": "port",
                        "sample": f"oracle train steps of 1 question ({N}x{L} tokens, t5-{size}, fp32) after 1 warm-up"}

    if rank == 0:
        algo_tf = 3 * fwd_flops(size, B, N, L, T) / 1e12  # per rank per step, padded shapes
        exec_tf = 3 * sum(fwd_flops_executed(size, x, T) for x in host_lens) / len(host_lens) / 1e12
        sec_step = ms * 1e-3 / steps
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: FiDT5 t5-{size} training step (fwd+bwd+clip+AdamW), per-GPU batch "
                                   f"{B} questions, n_passages={N}, text_maxlength={L}, T={T}, dropout=0.1 (model.train(), Philox masks)",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "inputs": "ids / masks / labels + the Collator's host-side passage lengths",
                       "l2": "step working set (>4 GB of activations) far exceeds the 126 MB L2; batches rotate; no explicit flush",
                       "emb_std": EMB_STD, "use_checkpoint": bool(args.use_checkpoint),
                       "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)},
            # algorithmic = SURVEY 8(d)'s dense count on the padded shapes (what the reference's own path computes);
            # executed = the same formula on the valid tokens only (the packed execution skips the pad rows).  The
            # tensor-pipe utilisation claim is the EXECUTED figure.
            "model_tflops_per_gpu": round(algo_tf / sec_step, 1),
            "executed_tflops_per_gpu": round(exec_tf / sec_step, 1),
            "frac_of_bf16_peak_executed": round(exec_tf / sec_step / peaks["bf16_tflops"], 4),
            "frac_of_bf16_peak_algorithmic": round(algo_tf / sec_step / peaks["bf16_tflops"], 4),
            "valid_token_fraction": round(valid_frac, 4),
            "unpadded_execution": bool(model.unpad),
            "peaks": {"bf16_tflops": peaks["bf16_tflops"], "bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"),
                      "hbm_gbs": peaks["hbm_gbs"], "source": which},
            "final_loss": final_loss,
            "gpu_launches": int(launches),
            "e2e": e2e,
            "roofline": roofline,
            "kernels": kernel_share,
            "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": eager,
            "generate": gen,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()  # rank 0 did the supplementary probes; leave together
        dist.destroy_process_group()


def run_eval_workload(args, model, size, B, N, L, dev, rank, world, peaks, which, lib, barrier, warmup, steps):
    """cfg3 / cfg5: the evaluation + attention-signal pass (test_reader.py:60-76) as the measured step: reset, greedy
    generate(max_length=50) forced to full length, get_crossattention_scores (21mean, last-half layers), answers and
    scores read back.  Replicas only under N>1 (inference shards over questions, nothing is exchanged but the final
    gather of answers / scores)."""
    import types
    import torch
    import torch.distributed as dist
    from lako_b200.optim import gather_answers_and_scores
    ML = 50
    host = synth_host_batches(B, N, L, 4, rank, n_batches=2, fact_format=True)
    resident = [(a.to(dev), b.to(dev)) for a, b, _ in host]
    opt = types.SimpleNamespace(stream=2, n_context=10, attention_score_style="21mean",
                                use_last_half_layer_attention="yes")
    model.eval()
    if args.precision == "fp32":
        model.set_precision("fp32")
    model.overwrite_forward_crossattention()

    def step(ids, mask):
        model.reset_score_storage()
        out = model.generate(input_ids=ids, attention_mask=mask, max_length=ML, min_length=ML)
        sc = model.get_crossattention_scores(opt, ids, None, mask)
        return out, sc

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(warmup):
        step(*resident[i % 2])
    barrier()
    launches0 = lib.lako_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        out, sc = step(*resident[i % 2])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.lako_launch_count() - launches0
    if sampler:
        sampler.stop_flag = True
        sampler.window = (t0, t1)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # e2e: pinned host inputs, answers + scores back on the host (and gathered on rank 0 when N>1)
    barrier()
    e0.record()
    for i in range(steps):
        a, b, _ = host[i % 2]
        out, sc = step(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True))
        ans, scs = gather_answers_and_scores(out, sc.to(dev), ML)
        _ = ans.cpu(), scs.cpu()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    a, b, _ = host[0]
    e2e = {"value": world * B * steps / (t.item() * 1e-3), "unit": "questions/s",
           "h2d_bytes_per_step": a.numel() * 8 + b.numel(), "d2h_bytes_per_step": B * ML * 8 + B * opt.n_context * 8}
    detail = eval_pass_numbers(model, size, B, N, L, dev, rank, peaks) if rank == 0 else None
    if rank == 0:
        ds = detail.get("decode_step", {})
        roofline = {"bound": "hbm", "kernel": "one cached decode step (xattn_decode + decoder GEMV kernels)",
                    "achieved": ds.get("achieved_GBs"), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ds.get("frac_hbm_peak"), "traffic": None, "peak_source": f"{which} copy bandwidth"}
        line = {"metric": METRIC_EVAL, "value": world * B * steps / (ms * 1e-3), "unit": "questions/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": ("tf32x3" if args.precision == "fp32" else "bf16"), "data": "synthetic",
                "config": {"workload": f"{args.workload}: FiDT5 t5-{size} evaluation + attention-signal pass, per-GPU batch {B} "
                                       f"questions, n_passages={N}, text_maxlength={L}, greedy generate(max_length={ML}) forced "
                                       f"to full length + get_crossattention_scores(21mean, last half)",
                           "parallelism": f"replicas x{world}", "l2": "cross-KV cache of the batch exceeds L2; no explicit flush"},
                "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "generate": detail,
                "cpu_baseline": None, "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
