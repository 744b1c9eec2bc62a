"""Single-kernel parity (through the C ABI debug hooks): tcgen05 GEMM + epilogues and the
fused attention kernel against fp32 numpy restatements on the same bf16-rounded inputs."""
import math

import numpy as np
import pytest

from kjarni_b200 import _native as N
from oracle import kjarni_oracle as ko
from kj_testutil import bf16_round, from_bf16_bits, ptr, to_bf16_bits

pytestmark = pytest.mark.gpu


def run_gemm(a, w, bias, res, epi, act=3, block_n=0):
    M, K = a.shape
    Nn = w.shape[0]
    ab, wb = to_bf16_bits(a), to_bf16_bits(w)
    f32out = epi in (2, 3)
    out = np.empty((M, Nn), np.float32 if f32out else np.uint16)
    N.check(N.lib().kjc_dbg_gemm(ptr(ab), ptr(wb), ptr(bias), ptr(res), M, Nn, K, epi, act, block_n, ptr(out)))
    return out if f32out else from_bf16_bits(out)


def ref_gemm(a, w, bias, res, epi, act):
    y = bf16_round(a).astype(np.float64) @ bf16_round(w).astype(np.float64).T
    if bias is not None:
        y = y + bias
    if epi == 1:
        y32 = y.astype(np.float32)
        y = {0: ko.gelu_erf, 1: ko.gelu_tanh, 2: lambda t: np.maximum(t, 0), 3: lambda t: t}[act](y32).astype(np.float64)
    if epi == 2:
        y = y + bf16_round(res)  # the residual stream is bf16 on device
    return y.astype(np.float32)


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (256, 384, 384), (300, 1152, 384), (1000, 384, 1536), (77, 96, 32),
                                    (4096, 1536, 384), (2048, 768, 3072), (130, 2304, 768), (130, 112, 64)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_matches_fp32(M, Nn, K, epi):
    rng = np.random.default_rng(M * 7 + Nn + K + epi)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((Nn, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(Nn).astype(np.float32)
    res = rng.standard_normal((M, Nn)).astype(np.float32) if epi == 2 else None
    got = run_gemm(a, w, bias, res, epi, act=0)
    want = ref_gemm(a, w, bias, res, epi, 0)
    tol = 1e-4 if epi in (2, 3) else 2.0 ** -8  # fp32 out: accumulation order only; bf16 out: one rounding
    err = np.abs(got - want) / (1.0 + np.abs(want))
    assert np.isfinite(got).all()
    assert err.max() < tol, (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("block_n", [64, 128, 192, 256])
def test_gemm_every_block_n_and_tails(block_n):
    rng = np.random.default_rng(block_n)
    M, Nn, K = 333, 416, 200  # M tail, N tail for every BN, K tail (K % 64 != 0)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = rng.standard_normal((Nn, K)).astype(np.float32) * 0.1
    bias = rng.standard_normal(Nn).astype(np.float32)
    got = run_gemm(a, w, bias, None, 3, block_n=block_n)
    want = ref_gemm(a, w, bias, None, 3, 3)
    assert np.abs(got - want).max() < 1e-3


@pytest.mark.parametrize("M,Nn,K", [(256, 768, 768), (300, 1152, 384), (1000, 2304, 768), (4096, 3072, 768), (18944, 1152, 384), (130, 416, 200),
                                    (128, 512, 64)])
@pytest.mark.parametrize("block_n", [192, 256])
@pytest.mark.parametrize("epi", [0, 1])
def test_gemm_cta_pair_matches_one_cta(M, Nn, K, block_n, epi):
    """The CTA-pair form of the GEMM (block_n + 2000: tcgen05.mma.cta_group::2 over two row tiles, half a weight tile per CTA, deeper
    ring) issues the same MMA shapes in the same k order as the one-CTA kernel: same bits, for odd row-tile counts (one padding tile),
    N and K tails, a single row tile, and against the fp32 oracle."""
    rng = np.random.default_rng(M + Nn + K + block_n + epi)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((Nn, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(Nn).astype(np.float32)
    one = run_gemm(a, w, bias, None, epi, act=0, block_n=block_n)
    pair = run_gemm(a, w, bias, None, epi, act=0, block_n=block_n + 2000)
    assert np.array_equal(one, pair)
    want = ref_gemm(a, w, bias, None, epi, 0)
    assert (np.abs(pair - want) / (1.0 + np.abs(want))).max() < 2.0 ** -8


@pytest.mark.parametrize("block_n", [128, 192, 256])
@pytest.mark.parametrize("Nn", [1152, 416, 1536])
@pytest.mark.parametrize("epi", [0, 1])
def test_gemm_bf16_epilogue_n_tails(block_n, Nn, epi):
    """bf16 (TMA-store) epilogues with N not a multiple of the tile width: QKV 1152 on 256-column tiles (4.5 tiles), 416."""
    rng = np.random.default_rng(block_n + Nn + epi)
    M, K = 300, 384
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((Nn, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(Nn).astype(np.float32)
    got = run_gemm(a, w, bias, None, epi, act=0, block_n=block_n)
    want = ref_gemm(a, w, bias, None, epi, 0)
    err = np.abs(got - want) / (1.0 + np.abs(want))
    assert np.isfinite(got).all() and err.max() < 2.0 ** -8, err.max()


@pytest.mark.parametrize("M,Nn,K,bn", [(256, 1152, 384, 192), (18944, 1152, 384, 192), (300, 1536, 384, 256), (130, 384, 192, 128),
                                       (4096 + 77, 1536, 384, 256), (100, 416, 320, 256)])
@pytest.mark.parametrize("epi", [0, 1])
def test_pair_gemm_matches_fp32(M, Nn, K, bn, epi):
    """CTA-pair (cta_group::2) A-resident GEMM (gemm_pair.cuh): M tails inside a 256-row pair tile, N tails, several pair counts."""
    if not N.lib().kjc_dbg_experimental_kernels():
        pytest.skip("experimental kernels (slower than the default path) are built only with make EXTRA=-DKJ_EXPERIMENTAL_KERNELS")
    rng = np.random.default_rng(M + Nn + K + epi)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((Nn, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(Nn).astype(np.float32)
    got = run_gemm(a, w, bias, None, epi, act=0, block_n=1000 + bn)
    want = ref_gemm(a, w, bias, None, epi, 0)
    err = np.abs(got - want) / (1.0 + np.abs(want))
    assert np.isfinite(got).all()
    assert err.max() < 2.0 ** -8, (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_activations(act):
    rng = np.random.default_rng(act)
    a = rng.standard_normal((256, 128)).astype(np.float32) * 2
    w = rng.standard_normal((256, 128)).astype(np.float32) * 0.3
    bias = rng.standard_normal(256).astype(np.float32)
    got = run_gemm(a, w, bias, None, 1, act=act)
    want = ref_gemm(a, w, bias, None, 1, act)
    tol = 3e-3 if act == 1 else 2.0 ** -8  # tanh.approx has ~5e-4 absolute error
    assert (np.abs(got - want) / (1.0 + np.abs(want))).max() < 2 * tol


@pytest.mark.parametrize("M,K,H", [(128, 384, 384), (1000, 384, 384), (300, 1536, 384), (18944, 384, 384), (77, 64, 384),
                                   (37965, 384, 384), (56900, 128, 384), (19000, 1536, 384),  # more row tiles than SMs: the ring <-> epilogue hand-back
                                   # hidden 768: a CTA pair per row tile, row statistics exchanged through distributed shared memory
                                   (128, 768, 768), (77, 64, 768), (1000, 3072, 768), (18944, 768, 768), (19000, 128, 768), (40000, 768, 768)])
def test_fused_gemm_residual_layernorm(M, K, H):
    """out-proj / FFN-down + bias + residual + LayerNorm in one kernel (hidden 384 / 768) vs the oracle's LayerNorm on fp32 sums."""
    import ctypes as C

    rng = np.random.default_rng(M + K + H)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((H, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(H).astype(np.float32) * 0.1
    gamma = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(H)).astype(np.float32)
    res = rng.standard_normal((M, H)).astype(np.float32)
    out = np.empty((M, H), np.uint16)
    us = C.c_float()
    N.check(N.lib().kjc_dbg_gemm_ln_h(ptr(to_bf16_bits(a)), ptr(to_bf16_bits(w)), ptr(bias), ptr(gamma), ptr(beta), 1e-12,
                                      ptr(to_bf16_bits(res)), M, H, K, ptr(out), 0, C.byref(us)))
    got = from_bf16_bits(out)
    y = (bf16_round(a).astype(np.float64) @ bf16_round(w).astype(np.float64).T + bias + bf16_round(res)).astype(np.float32)
    want = ko.layer_norm(y, gamma, beta, 1e-12)
    assert np.isfinite(got).all()
    assert (np.abs(got - want) / (1.0 + np.abs(want))).max() < 2.0 ** -7


@pytest.mark.parametrize("M,I", [(128, 1536), (300, 1536), (18944, 1536), (77, 128), (1000, 3072)])
def test_fused_ffn_layernorm(M, I):
    """FFN-up + erf-GELU + FFN-down + residual + LayerNorm in one kernel (hidden 384) vs the oracle chain in fp32, with the
    intermediate rounded to bf16 where the kernel rounds it."""
    if not N.lib().kjc_dbg_experimental_kernels():
        pytest.skip("experimental kernels (slower than the default path) are built only with make EXTRA=-DKJ_EXPERIMENTAL_KERNELS")
    import ctypes as C

    rng = np.random.default_rng(M + I)
    H = 384
    x = rng.standard_normal((M, H)).astype(np.float32)
    w1 = (rng.standard_normal((I, H)) / math.sqrt(H)).astype(np.float32)
    w2 = (rng.standard_normal((H, I)) / math.sqrt(I)).astype(np.float32)
    b1 = rng.standard_normal(I).astype(np.float32) * 0.1
    b2 = rng.standard_normal(H).astype(np.float32) * 0.1
    gamma = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(H)).astype(np.float32)
    out = np.empty((M, H), np.uint16)
    us = C.c_float()
    N.check(N.lib().kjc_dbg_ffn_ln(ptr(to_bf16_bits(x)), ptr(to_bf16_bits(w1)), ptr(b1), ptr(to_bf16_bits(w2)), ptr(b2), ptr(gamma), ptr(beta),
                                   1e-12, M, I, 0, ptr(out), 0, C.byref(us)))
    got = from_bf16_bits(out)
    xb = bf16_round(x).astype(np.float64)
    h = ko.gelu_erf((xb @ bf16_round(w1).astype(np.float64).T + b1).astype(np.float32))
    y = (bf16_round(h).astype(np.float64) @ bf16_round(w2).astype(np.float64).T + b2 + xb).astype(np.float32)
    want = ko.layer_norm(y, gamma, beta, 1e-12)
    assert np.isfinite(got).all()
    assert (np.abs(got - want) / (1.0 + np.abs(want))).max() < 2.0 ** -6


def ref_attention(qkv, mask, B, S, H, heads, noalloc):
    d = H // heads
    x = bf16_round(qkv).reshape(B, S, 3, heads, d)
    q, k, v = (x[:, :, i].transpose(0, 2, 1, 3) for i in range(3))
    s = (q @ k.transpose(0, 1, 3, 2)) * np.float32(1.0 / math.sqrt(d))
    fill = np.float32(-np.inf) if noalloc else ko.MASK_VALUE
    s = np.where((mask == 0)[:, None, None, :], fill, s).astype(np.float32)
    p = ko.softmax_rows(s)
    return (p @ v).transpose(0, 2, 1, 3).reshape(B * S, H)


@pytest.mark.parametrize("B,S,H,heads", [(3, 16, 64, 4), (2, 128, 384, 12), (3, 100, 128, 2), (2, 512, 128, 2), (2, 256, 64, 2), (5, 7, 32, 2), (2, 128, 128, 2), (4, 64, 768, 12), (3, 33, 64, 2), (150, 128, 384, 12)])
def test_attention_matches_fp32(B, S, H, heads):
    rng = np.random.default_rng(S + H)
    qkv = rng.standard_normal((B * S, 3 * H)).astype(np.float32)
    mask = np.ones((B, S), np.float32)
    mask[0, S // 2:] = 0  # padded tail
    if B > 1:
        mask[1, ::3] = 0  # holes, including position 0
    out = np.empty((B * S, H), np.uint16)
    N.check(N.lib().kjc_dbg_attention(ptr(to_bf16_bits(qkv)), ptr(mask), B, S, H, heads, 0, ptr(out)))
    got = from_bf16_bits(out)
    want = ref_attention(qkv, mask, B, S, H, heads, noalloc=False)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() < 2e-2  # P and the output are rounded to bf16


def test_attention_fully_padded_sequence_conventions():
    B, S, H, heads = 2, 32, 64, 2
    rng = np.random.default_rng(0)
    qkv = rng.standard_normal((B * S, 3 * H)).astype(np.float32)
    mask = np.ones((B, S), np.float32)
    mask[1, :] = 0
    out = np.empty((B * S, H), np.uint16)
    # alloc convention (-1e9): a fully padded sequence attends uniformly over all S keys
    N.check(N.lib().kjc_dbg_attention(ptr(to_bf16_bits(qkv)), ptr(mask), B, S, H, heads, 0, ptr(out)))
    got = from_bf16_bits(out)
    want = ref_attention(qkv, mask, B, S, H, heads, noalloc=False)
    assert np.abs(got - want).max() < 2e-2
    # no-alloc convention (-inf): the reference produces NaN for that sequence, finite elsewhere
    N.check(N.lib().kjc_dbg_attention(ptr(to_bf16_bits(qkv)), ptr(mask), B, S, H, heads, 1, ptr(out)))
    got = from_bf16_bits(out)
    assert np.isnan(got[S:]).all() and np.isfinite(got[:S]).all()


@pytest.mark.parametrize("M,K1,N2,epi2", [(128, 384, 1536, 1), (1000, 384, 1536, 1), (300, 1536, 1152, 0), (18944, 384, 1536, 1),
                                          (18944, 1536, 1152, 0), (77, 64, 208, 0)])
def test_chained_gemm_ln_gemm(M, K1, N2, epi2):
    """GEMM + residual + LayerNorm chained with the next projection in one launch (gemm_ln_gemm.cuh): x' must equal the fused
    GEMM+LN kernel bit for bit, the second projection must equal the stand-alone GEMM on x' bit for bit (same MMA shapes and k
    order), and both must match the fp32 oracle formulas."""
    import ctypes as C

    rng = np.random.default_rng(M + K1 + N2)
    H = 384
    a = rng.standard_normal((M, K1)).astype(np.float32)
    w1 = (rng.standard_normal((H, K1)) / math.sqrt(K1)).astype(np.float32)
    b1 = rng.standard_normal(H).astype(np.float32) * 0.1
    gamma = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(H)).astype(np.float32)
    res = rng.standard_normal((M, H)).astype(np.float32)
    w2 = (rng.standard_normal((N2, H)) / math.sqrt(H)).astype(np.float32)
    b2 = rng.standard_normal(N2).astype(np.float32) * 0.1
    out_x = np.empty((M, H), np.uint16)
    out2 = np.empty((M, N2), np.uint16)
    us = C.c_float()
    N.check(N.lib().kjc_dbg_gemm_ln_gemm(ptr(to_bf16_bits(a)), ptr(to_bf16_bits(w1)), ptr(b1), ptr(gamma), ptr(beta), 1e-12, ptr(to_bf16_bits(res)),
                                         M, K1, ptr(to_bf16_bits(w2)), ptr(b2), N2, epi2, 0, ptr(out_x), ptr(out2), 0, C.byref(us)))
    # the two-kernel path
    ref_x = np.empty((M, H), np.uint16)
    N.check(N.lib().kjc_dbg_gemm_ln(ptr(to_bf16_bits(a)), ptr(to_bf16_bits(w1)), ptr(b1), ptr(gamma), ptr(beta), 1e-12, ptr(to_bf16_bits(res)), M, K1,
                                    ptr(ref_x), 0, C.byref(us)))
    assert np.array_equal(out_x, ref_x)
    ref2 = np.empty((M, N2), np.uint16)
    N.check(N.lib().kjc_dbg_gemm(ptr(ref_x), ptr(to_bf16_bits(w2)), ptr(b2), None, M, N2, H, epi2, 0, 192, ptr(ref2)))
    assert np.array_equal(out2, ref2)
    # the CTA-pair variant (epi2 + 16: two CTAs per cluster share every weight tile through tcgen05.mma.cta_group::2, odd tile counts
    # leave one padding tile): the same bits again
    px = np.empty((M, H), np.uint16)
    p2 = np.empty((M, N2), np.uint16)
    N.check(N.lib().kjc_dbg_gemm_ln_gemm(ptr(to_bf16_bits(a)), ptr(to_bf16_bits(w1)), ptr(b1), ptr(gamma), ptr(beta), 1e-12, ptr(to_bf16_bits(res)),
                                         M, K1, ptr(to_bf16_bits(w2)), ptr(b2), N2, epi2 + 16, 0, ptr(px), ptr(p2), 0, C.byref(us)))
    assert np.array_equal(px, out_x)
    assert np.array_equal(p2, out2)
    # x' as the phase-2 A operand in tensor memory (epi2 + 32: TS-form tcgen05.mma, 128-column phase-2 tiles): the same bits again
    tx = np.empty((M, H), np.uint16)
    t2 = np.empty((M, N2), np.uint16)
    N.check(N.lib().kjc_dbg_gemm_ln_gemm(ptr(to_bf16_bits(a)), ptr(to_bf16_bits(w1)), ptr(b1), ptr(gamma), ptr(beta), 1e-12, ptr(to_bf16_bits(res)),
                                         M, K1, ptr(to_bf16_bits(w2)), ptr(b2), N2, epi2 + 32, 0, ptr(tx), ptr(t2), 0, C.byref(us)))
    assert np.array_equal(tx, out_x)
    assert np.array_equal(t2, out2)
    # and the oracle
    y = (bf16_round(a).astype(np.float64) @ bf16_round(w1).astype(np.float64).T + b1 + bf16_round(res)).astype(np.float32)
    want_x = ko.layer_norm(y, gamma, beta, 1e-12)
    got_x = from_bf16_bits(out_x)
    assert (np.abs(got_x - want_x) / (1.0 + np.abs(want_x))).max() < 2.0 ** -7
    want2 = ref_gemm(got_x, w2, b2, None, epi2, 0)
    assert (np.abs(from_bf16_bits(out2) - want2) / (1.0 + np.abs(want2))).max() < 2.0 ** -8
