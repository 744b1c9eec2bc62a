"""Where the hidden-state error of the bf16 encoder comes from (CPU emulation on the fp32 oracle, MiniLM-L6, random-init weights):
   A  bf16 GEMM operands only (activations and weights rounded where they enter a tensor-core GEMM; everything else fp32)
   B  A + the residual stream / LayerNorm outputs stored as bf16 between kernels (what the CUDA path does)
Prints max-abs / mean-abs error of the last hidden state against the pure fp32 oracle.  `python scripts/bf16_error_budget.py`"""
import sys, tempfile, os
sys.path.insert(0, ".")
import numpy as np
from kjarni_b200 import synth
from oracle import kjarni_oracle as ko


def bf16(x):
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def run(mode, m, ids, mask):
    lin0, ln0 = ko.linear, ko.layer_norm
    if mode in "AB":
        ko.linear = lambda x, w, b: lin0(bf16(x), bf16(w), b)
    if mode == "B":
        ko.layer_norm = lambda x, g, b, eps: bf16(ln0(x, g, b, eps))
    try:
        return ko.encoder_forward(m, ids, mask, None, noalloc=False)
    finally:
        ko.linear, ko.layer_norm = lin0, ln0


with tempfile.TemporaryDirectory() as td:
    d = synth.write_model_dir(os.path.join(td, "m"), "minilm-l6")
    m = ko.load_model_dir(d)
    ids, mask, _ = synth.synth_tokens(8, 64, 30522, regime="P", seed=7)
    ref = run("-", m, ids, mask)
    v = mask.astype(bool)
    for mode, what in (("A", "bf16 GEMM operands, fp32 residual stream"), ("B", "bf16 GEMM operands + bf16 residual stream")):
        h = run(mode, m, ids, mask)
        e = np.abs(h[v] - ref[v])
        print(f"{mode}: {what}: max-abs {e.max():.4f}  mean-abs {e.mean():.5f}  |hidden| max {np.abs(ref[v]).max():.2f}")
