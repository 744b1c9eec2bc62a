#!/usr/bin/env python
"""bench.py -- headline benchmark of the kjarni-b200 hot path.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): MiniLM-L6 seq-128 sentence embeddings per second (whole job, all GPUs);
the index top-k scan (queries/s) rides along in the same JSON line under "index_topk".
One step = one batch of synthetic token ids through the encoder forward + mean-pool + L2.
  value     : inputs already resident in HBM, device-timed (CUDA events), max over ranks
  e2e       : the same batch through the host-buffer C ABI call (H2D of ids/mask and D2H of the
              embeddings inside the timed region)
  roofline  : dominant kernel class, algorithmic FLOPs / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the fp32 oracle port of the reference CPU path, timed on this box's host cores
`--impl reference` times that CPU port alone (the reference is Rust and cannot be built here:
no cargo in the image -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = "minilm-l6"
SEQ = 128
METRIC = "MiniLM-L6 seq128 embeddings/sec"
UNIT = "embeddings/s"


def flops_per_seq(H, L, I, S):
    """projection/FFN GEMMs + attention matmuls, 2 flop/MAC."""
    return L * S * (8 * H * H + 4 * H * I + 4 * S * H)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md).  The poller is started BEFORE the warm-up
    steps -- nvidia-smi's start-up (NVML initialisation, first query) stalls kernel launches for milliseconds, which belongs into the
    warm-up, not into a 50-150 ms timed region -- and `stop(t0, t1)` keeps the samples whose arrival time lies inside the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_wait = time.time() + 3.0  # the first sample has arrived = the poller's start-up is over
            while not self.rows and time.time() < t_wait and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1 + 0.15)]  # a sample is reported up to one period after it was taken
        if not rows:  # timed region shorter than one sampling period: the samples closest to it (taken under the same load, in the warm-up)
            rows = [r for _, r in self.rows[-2:]]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def _host_threads():
    """Threads the CPU arm uses: every physical core, as kjarni_init sizes the reference's rayon pool
    (kjarni-ffi/src/lib.rs:36-40).  torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the port to one thread."""
    try:
        import psutil

        n = psutil.cpu_count(logical=False) or 0
    except Exception:
        n = 0
    n = n or (os.cpu_count() or 1)
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, n)


def _cpu_model(model_dir):
    """The CPU port of the reference path: the C restatement (oracle/kjarni_oracle.c, AVX2 + OpenMP, structured like the
    reference's own kernels) when its library was built, else the numpy restatement.  Returns (embed_fn, cores, what)."""
    from oracle import kjarni_oracle as ko

    m = ko.load_model_dir(model_dir)
    try:
        from oracle import kjarni_oracle_c as koc

        cm = koc.CModel(m)
        koc.set_num_threads(_host_threads())
        return (lambda ids, mask: cm.embed(ids, mask.astype(np.float32))), koc.num_threads(), \
            "C restatement of the reference CPU kernels (oracle/kjarni_oracle.c: 64-token blocks x 4x3 AVX2/FMA micro-kernel, OpenMP)"
    except (ImportError, OSError):
        return (lambda ids, mask: ko.embed(m, ids, mask)), os.cpu_count(), "numpy fp32 restatement (oracle/kjarni_oracle.py)"


def cpu_port_embed_rate(model_dir, nseq, seconds=12.0):
    """The reference CPU path restated (oracle/), timed here as the CPU baseline on BASELINE configs[0] batches."""
    from kjarni_b200 import synth

    fn, cores, what = _cpu_model(model_dir)
    ids, mask, _ = synth.synth_tokens(nseq, SEQ, synth.ARCHS[ARCH][5], regime="T", seed=42)
    fn(ids[:4], mask[:4])  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        fn(ids, mask)
        n += nseq
        dt = time.perf_counter() - t0
        if dt > seconds:
            break
    return n / dt, dt, n, cores, what


def cpu_scan_rate(dim, k, rows, nq):
    """Segment::search_vectors restated in C (scalar pass per query, OpenMP over queries) on a bounded slice."""
    from oracle import kjarni_oracle as ko
    from oracle import kjarni_oracle_c as koc

    koc.set_num_threads(_host_threads())
    r = ko.synth_rows(7, 0, rows, dim)
    q = ko.synth_rows(11, 0, nq, dim)
    koc.scan_topk(r[:1000], q, k)
    t0 = time.perf_counter()
    koc.scan_topk(r, q, k)
    dt = time.perf_counter() - t0
    return rows * nq / dt, dt, koc.num_threads()


def model_dir_for(rank):
    from kjarni_b200 import synth

    d = os.path.join(tempfile.gettempdir(), f"kjarni_b200_bench_{os.getuid()}_{rank}", ARCH)
    if not os.path.exists(os.path.join(d, "model.safetensors")):
        synth.write_model_dir(d, ARCH)
    return d


def run_reference(args, rank, world):
    """The reference arm: the reference's CPU path restated (kind "port": the Rust reference cannot be built here) with all
    host threads OpenMP gives it, on BASELINE configs[0] batches (32 x 128 tokens) of the same synthetic workload."""
    if rank != 0:
        return
    nseq = 32  # BASELINE configs[0]: batch 32 x seq 128, the reference's own CPU-runnable case
    d = model_dir_for(0)
    from kjarni_b200 import synth

    fn, cores, what = _cpu_model(d)
    ids, mask, _ = synth.synth_tokens(nseq, SEQ, synth.ARCHS[ARCH][5], regime="T", seed=42)
    for _ in range(max(args.warmup, 1)):
        fn(ids, mask)
    steps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(ids, mask)
    dt = time.perf_counter() - t0
    val = nseq * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic token ids, random-init weights (no network)",
        "config": {"workload": "all-MiniLM-L6-v2 architecture sentence embedding (mean-pool + L2), seq 128",
                   "sample": f"{nseq} sequences x {SEQ} tokens per step (BASELINE configs[0])"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {nseq} seqs x {SEQ} tokens; {what}; "
                                   "the Rust reference cannot be built in this image (no cargo)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def emit(line: dict) -> None:
    """Exactly one JSON line on the process's real stdout (see main(): fd 1 is pointed at stderr for everything else)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # Libraries print banners on stdout (NCCL: "NCCL version ..." at communicator creation); the contract is ONE JSON line
    # there, so fd 1 is redirected to stderr and the result line goes to a duplicate of the original stdout.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (default 28 micro-batches)")
    ap.add_argument("--index-rows", type=int, default=6_250_000, help="index shard rows per GPU (BASELINE config 4: 50M/8)")
    ap.add_argument("--no-index", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1-latency / C2 / C3 / C5 lines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from kjarni_b200 import _native as N
    from kjarni_b200 import api, synth

    torch.cuda.set_device(local_rank)
    # a real (non-NULL) stream: the library treats NULL as "my own stream", and torch events only see the current stream
    torch.cuda.set_stream(torch.cuda.Stream())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = N.lib()
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d = model_dir_for(rank)
    enc = api.EncoderModel(d, device=local_rank)
    info = enc.info
    H, L, I = info.hidden_size, info.num_layers, info.intermediate_size
    mb = enc.micro_batch(SEQ)
    B = args.batch or 28 * mb  # 28 x 148 = 4144 sequences per GPU per step
    ids_np, mask_np, _ = synth.synth_tokens(B, SEQ, info.vocab_size, regime="T", seed=42 + rank)
    maskf_np = mask_np.astype(np.float32)
    ids_d = torch.from_numpy(ids_np.view(np.int32)).cuda()
    mask_d = torch.from_numpy(maskf_np).cuda()
    out_d = torch.empty((B, H), dtype=torch.float32, device="cuda")
    opts = N.KjcForwardOptions(N.OUT_POOLED, N.POOL_MEAN, 1, N.MASK_AUTO)
    stream = torch.cuda.current_stream().cuda_stream

    def step_dev():
        N.check(lib.kjc_encoder_forward_device_async(enc._h, ids_d.data_ptr(), mask_d.data_ptr(), None, B, SEQ, C.byref(opts),
                                                     out_d.data_ptr(), stream))

    # ---------------------------------------------------------------- value: device-resident inputs
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_dev()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    barrier()
    ms = max_over_ranks(ms)
    launches_per_step = enc.last_launch_count
    value = world * B * args.steps / (ms / 1e3)

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI
    out_h = np.empty((B, H), np.float32)
    for _ in range(2):
        enc.encode_batch_from_ids(ids_np, maskf_np)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        N.check(lib.kjc_encoder_forward(enc._h, ids_np.ctypes.data, maskf_np.ctypes.data, None, B, SEQ, C.byref(opts), out_h.ctypes.data))
    dt = time.perf_counter() - t0
    barrier()
    dt = max_over_ranks(dt)
    e2e_val = world * B * e2e_steps / dt

    # ---------------------------------------------------------------- roofline: per-kernel-class CUDA events
    N.check(lib.kjc_encoder_set_profiling(enc._h, 1))
    prof_steps = 2
    for _ in range(prof_steps):
        step_dev()
    torch.cuda.synchronize()
    pms = (C.c_double * 8)()
    pn = (C.c_int64 * 8)()
    N.check(lib.kjc_encoder_get_profile(enc._h, pms, pn))
    N.check(lib.kjc_encoder_set_profiling(enc._h, 0))
    M = B * SEQ
    tf = {"gemm_qkv": 2.0 * M * 3 * H * H * L, "gemm_out": 2.0 * M * H * H * L, "gemm_ffn_up": 2.0 * M * H * I * L,
          "gemm_ffn_down": 2.0 * M * H * I * L, "attention": 4.0 * M * SEQ * H * L}
    gb = {"layernorm": 2.0 * L * M * H * 10, "embed_ln": M * 4.0 + M * H * 4.0 + M * H * 6.0, "output": M * H * 4.0 + B * H * 4.0}
    if pn[N.KERNEL_CLASSES.index("gemm_ffn_down")] == 0:  # fused feed-forward kernel: up + GELU + down + residual + LayerNorm in one launch
        tf["gemm_ffn_up"] += tf["gemm_ffn_down"]
    chained = bool(lib.kjc_encoder_chained(enc._h))
    contains = {}
    if chained:
        # chained launches (gemm_ln_gemm.cuh): class gemm_ffn_up = out-proj + LN1 + FFN-up, class gemm_ffn_down = FFN-down + LN2 + the
        # NEXT layer's QKV (plain FFN-down + LN2 in the last layer), class gemm_qkv = layer 0 only; their algorithmic flops follow
        qkv_l = tf["gemm_qkv"] / L
        tf["gemm_ffn_up"] += tf.pop("gemm_out")
        tf["gemm_ffn_down"] += qkv_l * (L - 1)
        tf["gemm_qkv"] = qkv_l
        embed_chained = pn[N.KERNEL_CLASSES.index("embed_ln")] == 0
        contains = {"gemm_qkv": "embedding gather + embed LN -> QKV projection of layer 0, one launch" if embed_chained else "QKV projection of layer 0", "gemm_ffn_up": "out-proj + residual + LN1 -> FFN-up + GELU, one launch per layer",
                    "gemm_ffn_down": "FFN-down + residual + LN2 -> next layer's QKV, one launch per layer (last layer: FFN-down + LN2 only)"}
    kernels = {}
    tot_ms = sum(pms[i] for i in range(8)) / prof_steps
    for i, name in enumerate(N.KERNEL_CLASSES):
        ms_i = pms[i] / prof_steps
        if pn[i] == 0:
            continue
        k = {"ms_per_step": round(ms_i, 4), "share": round(ms_i / tot_ms, 4), "launches_per_step": int(pn[i] // prof_steps)}
        if name in tf:
            k.update(bound="tensor", achieved=round(tf[name] / (ms_i * 1e-3) / 1e12, 2), unit="TFLOP/s", peak=peaks["tf_sustained"])
        else:
            k.update(bound="hbm", achieved=round(gb[name] / (ms_i * 1e-3) / 1e9, 1), unit="GB/s", peak=peaks["hbm_gbs"])
        k["frac"] = round(k["achieved"] / k["peak"], 4)
        if name in contains:
            k["contains"] = contains[name]
        kernels[name] = k
    dom = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
    traffic = traffic_in_step = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(dom)                               # cold-cache: one ncu --set full capture
            traffic_in_step = (tj.get("in_step") or {}).get(dom)  # ncu --cache-control none: what the launch moves inside a step
        except Exception:
            traffic = traffic_in_step = None
    roofline = {"kernel": dom, "bound": kernels[dom]["bound"], "achieved": kernels[dom]["achieved"], "peak": kernels[dom]["peak"],
                "unit": kernels[dom]["unit"], "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_in_step": traffic_in_step,
                "peak_source": peaks["src"] + (" bf16_tflops_sustained" if kernels[dom]["bound"] == "tensor" else " hbm_gbs"),
                "how": f"CUDA events around every launch, {prof_steps} extra steps after the timed region",
                "whole_step": {"achieved": round(value / world * flops_per_seq(H, L, I, SEQ) / 1e12, 2), "unit": "TFLOP/s",
                               "frac": round(value / world * flops_per_seq(H, L, I, SEQ) / 1e12 / peaks["tf_sustained"], 4)},
                "kernels": kernels}

    # ---------------------------------------------------------------- index top-k scan (second half of the metric)
    index = None
    if not args.no_index:
        try:
            index = bench_index(args, rank, world, local_rank, lib, N, api, torch, dist, barrier, max_over_ranks, peaks)
        except Exception as ex:  # keep the headline line even if the shard does not fit
            index = {"error": str(ex)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, secs, nseq, cores, what = cpu_port_embed_rate(d, 32)
        cpu = {"value": round(rate, 2), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nseq} sequences (batches of 32 x {SEQ} tokens = BASELINE configs[0]) in {secs:.1f} s; {what}"}
        if isinstance(index, dict) and "error" not in index:
            try:
                rq, secs_s, cores_s = cpu_scan_rate(384, 10, 1_000_000, 64 * cores)
                index["cpu_baseline"] = {"value": round(rq / index["rows_per_gpu"], 3), "unit": "queries/s", "cores": cores_s, "kind": "port",
                                         "sample": f"{64 * cores} queries x 1000000 rows x 384 dims in {secs_s:.1f} s (C restatement of "
                                                   "Segment::search_vectors, OpenMP over queries), scaled linearly to rows_per_gpu"}
            except Exception as ex:
                index["cpu_baseline"] = {"error": str(ex)}

    extra = {}
    if not args.no_extra:
        for name in EXTRA_CONFIGS:
            try:
                extra[name] = bench_extra(name, args, rank, world, local_rank, lib, N, api, torch, barrier, max_over_ranks, peaks,
                                          with_cpu=(rank == 0 and world == 1 and not args.no_cpu))
            except Exception as ex:
                extra[name] = {"error": str(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic token ids, random-init weights (no network)",
            "config": {"workload": f"all-MiniLM-L6-v2 architecture sentence embedding (mean-pool + L2), seq {SEQ}, "
                                   f"{B} sequences per GPU per step in micro-batches of {mb}",
                       "global_batch": B * world, "seq_len": SEQ, "parallelism": f"dp{world} (batch split, no collective)",
                       "l2": "activations of one step exceed L2 many times over; the 45 MB of bf16 weights stay L2-resident by design"},
            "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": int(ids_np.nbytes + maskf_np.nbytes),
                    "d2h_bytes_per_step": int(out_h.nbytes), "steps": e2e_steps, "timing": "wall clock around synchronous C-ABI calls, max over ranks"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "index_topk": index, "configs": extra,
        }
        emit(line)
    enc.close()
    if world > 1:
        dist.destroy_process_group()


# BASELINE.json configs beyond the headline ("plus classifier seq/s (C2), rerank pairs/s (C3), BERT-base seq/s (C5)"),
# and the reference's own batch-32 case as a latency line.  name -> (arch, batch per GPU per step, seq, output, pair inputs, unit, what)
EXTRA_CONFIGS = {
    "c1_batch32_latency": ("minilm-l6", 32, 128, "pooled", False, "embeddings/s",
                           "configs[0]: all-MiniLM-L6-v2 architecture, ONE batch of 32 x 128 per call (latency case)"),
    "c2_classify": ("distilbert-sst2", 256, 128, "logits", False, "sequences/s",
                    "configs[1]: DistilBERT (distilbert-sentiment architecture) sequence classification, batch 256 x seq 128 per GPU"),
    "c3_rerank": ("minilm-l6-cross-encoder", 1000, 256, "logits", True, "pairs/s",
                  "configs[2]: MiniLM-L6 cross-encoder, 64 queries x 1000 candidate passages, seq 256; one step = one query's 1000 pairs"),
    "c5_bertbase": ("bert-base", 512, 512, "pooled", False, "sequences/s",
                    "configs[4]: BERT-base-size (768-dim, 12-layer) embedding, batch 512 x seq 512 per GPU"),
}


def _model_dir(arch, rank):
    from kjarni_b200 import synth

    d = os.path.join(tempfile.gettempdir(), f"kjarni_b200_bench_{os.getuid()}_{rank}", arch)
    if not os.path.exists(os.path.join(d, "model.safetensors")):
        synth.write_model_dir(d, arch)
    return d


def cpu_port_rate(model_dir, S, pair, seconds, nseq):
    """The reference CPU forward (encoder hidden states; the head is O(H^2) per sequence) restated in C, on a bounded sample."""
    from kjarni_b200 import synth
    from oracle import kjarni_oracle as ko
    from oracle import kjarni_oracle_c as koc

    m = ko.load_model_dir(model_dir)
    cm = koc.CModel(m)
    koc.set_num_threads(_host_threads())
    ids, mask, types = synth.synth_tokens(nseq, S, m.word.shape[0], regime="T", seed=42, pair=pair)
    maskf = mask.astype(np.float32)
    t0 = time.perf_counter()
    n = 0
    while True:
        cm.hidden_states(ids, maskf, types, noalloc=False)
        n += nseq
        dt = time.perf_counter() - t0
        if dt > seconds:
            break
    return n / dt, dt, n, koc.num_threads()


def bench_extra(name, args, rank, world, local_rank, lib, N, api, torch, barrier, max_over_ranks, peaks, with_cpu):
    from kjarni_b200 import synth

    arch, B, S, out, pair, unit, what = EXTRA_CONFIGS[name]
    enc = api.EncoderModel(_model_dir(arch, rank), device=local_rank)
    info = enc.info
    H, L, I = info.hidden_size, info.num_layers, info.intermediate_size
    ids_np, mask_np, types_np = synth.synth_tokens(B, S, info.vocab_size, regime="T", seed=4242 + rank, pair=pair)
    maskf_np = mask_np.astype(np.float32)
    ids_d = torch.from_numpy(ids_np.view(np.int32)).cuda()
    mask_d = torch.from_numpy(maskf_np).cuda()
    types_d = torch.from_numpy(types_np.view(np.int32)).cuda() if pair else None
    cols = H if out == "pooled" else info.num_labels
    out_d = torch.empty((B, cols), dtype=torch.float32, device="cuda")
    out_h = np.empty((B, cols), np.float32)
    opts = N.KjcForwardOptions(N.OUT_POOLED if out == "pooled" else N.OUT_LOGITS, N.POOL_MEAN, 1, N.MASK_AUTO)
    stream = torch.cuda.current_stream().cuda_stream
    tptr = types_d.data_ptr() if pair else None
    tptr_h = types_np.ctypes.data if pair else None

    def step_dev():
        N.check(lib.kjc_encoder_forward_device_async(enc._h, ids_d.data_ptr(), mask_d.data_ptr(), tptr, B, S, C.byref(opts), out_d.data_ptr(), stream))

    def step_host():
        N.check(lib.kjc_encoder_forward(enc._h, ids_np.ctypes.data, maskf_np.ctypes.data, tptr_h, B, S, C.byref(opts), out_h.ctypes.data))

    fl = flops_per_seq(H, L, I, S)
    # steps: C3 = the whole 64-query job; the others enough steps for ~0.3 s of device time
    steps = 64 if name == "c3_rerank" else max(args.steps, int(min(2000, max(10, 0.15 * peaks["tf_sustained"] * 1e12 / (fl * B)))))
    for _ in range(3):
        step_dev()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_dev()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    launches = enc.last_launch_count
    barrier()
    for _ in range(2):
        step_host()
    e_steps = max(3, steps // 4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        step_host()
    e_ms = max_over_ranks((time.perf_counter() - t0) / e_steps * 1e3)
    # per-kernel-class shares (CUDA events around every launch)
    N.check(lib.kjc_encoder_set_profiling(enc._h, 1))
    step_dev()
    torch.cuda.synchronize()
    pms = (C.c_double * 8)()
    pn = (C.c_int64 * 8)()
    N.check(lib.kjc_encoder_get_profile(enc._h, pms, pn))
    N.check(lib.kjc_encoder_set_profiling(enc._h, 0))
    tot = sum(pms[i] for i in range(8)) or 1.0
    shares = {n: {"ms": round(pms[i], 4), "share": round(pms[i] / tot, 4), "launches": int(pn[i])} for i, n in enumerate(N.KERNEL_CLASSES) if pn[i]}
    rate = world * B / (ms * 1e-3)
    tfs = rate / world * fl / 1e12
    res = {"workload": what, "value": round(rate, 1), "unit": unit, "ms_per_step": round(ms, 4), "steps": steps, "batch_per_gpu": B, "seq_len": S,
           "e2e": {"value": round(world * B / (e_ms * 1e-3), 1), "unit": unit, "ms_per_step": round(e_ms, 4), "steps": e_steps,
                   "h2d_bytes_per_step": int(ids_np.nbytes + maskf_np.nbytes + (types_np.nbytes if pair else 0)), "d2h_bytes_per_step": int(out_h.nbytes)},
           "roofline": {"bound": "tensor", "achieved": round(tfs, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": round(tfs / peaks["tf_sustained"], 4),
                        "flops_per_seq": fl, "note": "F_seq = L*S*(8H^2 + 4HI + 4SH) (SURVEY 8d) x sequences/s per GPU vs bf16_tflops_sustained"},
           "gpu_launches_per_step": int(launches), "kernels": shares}
    if name == "c1_batch32_latency":
        res["latency_ms"] = {"device": round(ms, 4), "e2e_host_buffers": round(e_ms, 4)}
    if with_cpu:
        try:
            nseq = {"c1_batch32_latency": 32, "c2_classify": 32, "c3_rerank": 16, "c5_bertbase": 4}[name]
            r, secs, n, cores = cpu_port_rate(_model_dir(arch, rank), S, pair, 4.0, nseq)
            res["cpu_baseline"] = {"value": round(r, 2), "unit": unit, "cores": cores, "kind": "port",
                                   "sample": f"{n} sequences x {S} tokens (batches of {nseq}) in {secs:.1f} s; C restatement of the reference CPU forward (oracle/kjarni_oracle.c)"}
        except Exception as ex:
            res["cpu_baseline"] = {"error": str(ex)}
    enc.close()
    return res


def bench_index(args, rank, world, local_rank, lib, N, api, torch, dist, barrier, max_over_ranks, peaks):
    """Row-sharded cosine top-10 (BASELINE config 4): every rank searches its shard for the same query batch, the per-shard
    [Q,k] candidates are gathered over NCCL (all_gather) and merged by the merge kernel.  Two regimes are timed:
      batch  4096 queries per step: tensor-core filter GEMM over the bf16 shadow + exact fp32 rescoring (tensor-bound)
      small  8 queries per step: the same filter path, one pass over the bf16 shadow per step (HBM-bound)
      exact  8 queries per step on the exact fp32 scan kernel alone (the fallback path; fp32 index bytes read once per step)"""
    dim, k = 384, 10
    n = args.index_rows
    sh = api.IndexShard(dim, n, id_base=rank * n, device=local_rank)
    sh.append_synthetic(7, rank * n, n)
    from oracle import kjarni_oracle as ko

    stream = torch.cuda.current_stream().cuda_stream

    def regime(nq, steps, exact_only=False):
        sh.set_filter(min_queries=(1 << 30) if exact_only else 1)
        q_np = ko.synth_rows(11, 0, nq, dim)
        q_d = torch.from_numpy(q_np).cuda()
        # this rank's candidates as ONE packed record ([nq,k] u64 ids | [nq,k] f32 scores): one all_gather per search
        rb = int(lib.kjc_packed_record_bytes(nq, k))
        rec_d = torch.empty((rb,), dtype=torch.uint8, device="cuda")
        cnt_d = torch.empty((nq,), dtype=torch.int32, device="cuda")
        if world > 1:
            g_rec = torch.empty((world * rb,), dtype=torch.uint8, device="cuda")  # rank-major concat
            f_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            f_sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")

        def step():
            # the always-exact entry: one 4-byte read-back per search, unproven queries re-run on the exact scan
            N.check(lib.kjc_index_search_device(sh._h, q_d.data_ptr(), nq, k, N.SCAN_SEGMENT, rec_d.data_ptr(), rec_d.data_ptr() + nq * k * 8,
                                                cnt_d.data_ptr(), stream))
            if world > 1:
                dist.all_gather_into_tensor(g_rec, rec_d)
                N.check(lib.kjc_topk_merge_packed_device_async(local_rank, g_rec.data_ptr(), world, nq, k, f_ids.data_ptr(), f_sc.data_ptr(),
                                                               None, stream))

        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        launches = sh.last_launch_count
        barrier()
        # host-buffer call (H2D of the queries, D2H of ids/scores inside) on this rank's shard
        sh.search_batch(q_np, k)
        t0 = time.perf_counter()
        e2e_steps = max(2, steps // 2)
        for _ in range(e2e_steps):
            sh.search_batch(q_np, k)
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
        return {"value": round(nq / (ms * 1e-3), 1), "unit": "queries/s", "ms_per_step": round(ms, 4), "queries_per_step": nq,
                "e2e": {"value": round(nq / (e2e_ms * 1e-3), 1), "unit": "queries/s", "h2d_bytes_per_step": int(q_np.nbytes),
                        "d2h_bytes_per_step": nq * k * 12 + nq * 4, "note": "host-buffer kjc_index_search on this rank's shard"},
                "gpu_launches_per_step": launches}, ms

    steps = max(args.steps, 10)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    batch, ms_b = regime(4096, steps)
    tf = 2.0 * 4096 * n * dim / (ms_b * 1e-3) / 1e12  # per GPU: the filter GEMM is the dominant kernel
    batch["roofline"] = {"bound": "tensor", "achieved": round(tf, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": round(tf / peaks["tf_sustained"], 4), "traffic": None,
                         "note": "2*Q*N*D flops of the bf16 filter GEMM per GPU / whole search time (seed + filter + select + exact "
                                 "rescoring); arithmetic intensity 2Q/2 flop per shadow byte = 4096 flop/B, far above machine balance"}
    small, ms_s = regime(8, steps)
    gbs = (n * dim * 2.0) / (ms_s * 1e-3) / 1e9  # one pass over the bf16 shadow per GPU
    small["roofline"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": traffic.get("index_filter_8q"),
                         "fp32_index_equivalent": {"achieved": round(2 * gbs, 1), "unit": "GB/s", "frac": round(2 * gbs / peaks["hbm_gbs"], 4),
                                                   "note": "SURVEY 8(d) counts N*D*4 bytes (the fp32 index) per search; this path returns the same "
                                                           "exact fp32 top-k while reading half of them, hence a figure above the copy roof"},
                         "note": "algorithmic bytes = N*D*2: the bf16 shadow of the index is read ONCE per search (a different, smaller byte "
                                 "count than SURVEY 8(d)'s N*D*4 -- see fp32_index_equivalent) / whole search time (prep + seed + filter + "
                                 "select + exact fp32 rescoring of 32 candidates per query); traffic = ncu dram bytes of the seed + filter launches"}
    unverified = sh.unverified_count
    exact, ms_e = regime(8, steps, exact_only=True)
    gbs = (n * dim * 4.0 + n * 4.0) / (ms_e * 1e-3) / 1e9  # rows + cached norms, per GPU
    exact["roofline"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": traffic.get("index_exact_8q"),
                         "note": "exact fp32 scan kernel alone (fallback / proof path): query norms + scan + merge per GPU; algorithmic bytes = "
                                 "N*D*4 + N*4 (fp32 rows + cached norms, SURVEY 8(d)'s definition) read once per 8-query pass"}
    sh.set_filter(min_queries=1)
    res = dict(batch)
    res.update({"k": k, "rows_per_gpu": n, "rows_total": n * world, "dim": dim, "dtype": "bf16 filter + f32 exact rescoring",
                "merge": "none (1 shard)" if world == 1 else "ONE NCCL all_gather of the packed per-shard [Q,k] (id, score) records + merge kernel",
                "exactness": "kjc_index_search_device: every query proven exact or re-run on the exact fp32 scan inside the timed region",
                "unverified_queries": unverified, "small_batch_8q": small, "exact_scan_8q": exact})
    sh.close()
    return res


if __name__ == "__main__":
    main()
