"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol the headers declare, and its
host-only entry points (errors, cosine, softmax) match the reference's known answers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("kjarni_cuda.h", "kjarni_cuda_debug.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(kjc_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(N.SIGNATURES), decl ^ set(N.SIGNATURES)  # the ctypes table covers the whole ABI


def test_error_codes_and_names_match_reference():
    lib = N.lib()
    # KjarniErrorCode, kjarni-ffi/src/error.rs:14-40
    want = {0: b"Ok", 1: b"NullPointer", 2: b"InvalidUtf8", 3: b"ModelNotFound", 4: b"LoadFailed", 5: b"InferenceFailed",
            6: b"GpuUnavailable", 7: b"InvalidConfig", 8: b"Cancelled", 9: b"Timeout", 10: b"StreamEnded", 255: b"Unknown"}
    for code, name in want.items():
        assert lib.kjc_error_name(code) == name
    assert b"sm_100a" in lib.kjc_version()


def test_null_pointers_and_thread_local_error():
    lib = N.lib()
    lib.kjc_clear_error()
    assert lib.kjc_last_error_message() is None
    assert lib.kjc_encoder_create(None, 0, None) == 1
    h = C.c_void_p(123)
    assert lib.kjc_encoder_create(None, 0, C.byref(h)) == 1 and not h.value  # out-param cleared on error
    assert b"null pointer" in lib.kjc_last_error_message()
    assert lib.kjc_index_search(None, None, 1, 1, 0, None, None, None) == 1
    lib.kjc_clear_error()
    assert lib.kjc_last_error_message() is None
    if lib.kjc_device_count() <= 0:  # product path fails loudly without a GPU: no CPU fallback
        with pytest.raises(N.KjarniCudaError) as e:
            api.EncoderModel("/tmp")
        assert e.value.status == 6
        with pytest.raises(N.KjarniCudaError) as e:
            api.IndexShard(384, 10)
        assert e.value.status == 6


def test_cosine_similarity_kats(kats):
    # kjarni-search/src/vector.rs:386-398 and kjarni-ffi/src/lib.rs:177-188 semantics
    assert api.cosine_similarity([1, 0, 0], [1, 0, 0]) == pytest.approx(1.0, abs=1e-6)
    assert api.cosine_similarity([1, 0, 0], [0, 1, 0]) == pytest.approx(0.0, abs=1e-6)
    assert api.cosine_similarity([1, 0, 0], [-1, 0, 0]) == pytest.approx(-1.0, abs=1e-6)
    assert api.cosine_similarity([0, 0, 0], [1, 2, 3]) == 0.0  # denominator clamps to 1e-9
    assert api.cosine_similarity([1, 2], [1, 2, 3]) == 0.0
    assert N.lib().kjc_cosine_similarity(None, None, 3) == 0.0
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(384).astype(np.float32), rng.standard_normal(384).astype(np.float32)
    from oracle import kjarni_oracle as ko

    assert api.cosine_similarity(a, b) == pytest.approx(ko.cosine_similarity(a, b), abs=1e-6)


def test_softmax_rows_matches_reference_goldens(kats):
    from oracle import kjarni_oracle as ko

    x = np.array([[1.0, 2.0, 3.0], [1000.0, 1000.0, 1000.0], [-1e9, 0.0, -1e9]], np.float32)
    got = x.copy()
    N.lib().kjc_softmax_rows(got.ctypes.data_as(C.c_void_p), 3, 3)
    assert np.allclose(got, ko.softmax_rows(x), atol=1e-7)
    assert np.allclose(got.sum(1), 1, atol=1e-6)


def test_api_scores_to_top_k_is_stable():
    # ties resolve to the lowest label index (kjarni-models/src/models/sequence_classifier/mod.rs:369-372)
    r = api.scores_to_top_k(np.array([0.25, 0.5, 0.25, 0.0], np.float32), ["a", "b", "c", "d"], 3)
    assert [n for n, _ in r] == ["b", "a", "c"]


def test_public_ffi_symbols_are_exported():
    """Every kjarni_* entry point include/kjarni_ffi.h declares is exported under the reference's names."""
    txt = open(os.path.join(ROOT, "include", "kjarni_ffi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(kjarni_[a-z0-9_]+)\s*\(", txt))
    assert len(names) >= 40
    lib = C.CDLL(os.path.join(os.path.dirname(N.LIB_PATH), "libkjarni_ffi.so"))
    for name in sorted(names):
        assert hasattr(lib, name), name
    lib.kjarni_error_name.restype = C.c_char_p
    assert lib.kjarni_error_name(6) == b"GpuUnavailable" and lib.kjarni_init() == 0
    lib.kjarni_cosine_similarity.restype = C.c_float
    a = (C.c_float * 3)(1, 0, 0)
    assert lib.kjarni_cosine_similarity(a, a, 3) == pytest.approx(1.0)
    # frees accept NULL and empty results (pointer form, as the Rust source declares them)
    for fn in ("kjarni_float_array_free", "kjarni_float_2d_array_free", "kjarni_string_array_free", "kjarni_class_results_free",
               "kjarni_rerank_results_free", "kjarni_search_results_free", "kjarni_string_free"):
        getattr(lib, fn)(None)


def test_headers_are_valid_c99_and_link(tmp_path):
    """The three headers are consumed from C (cgo / P/Invoke generators / a Rust bindgen run): they must parse as plain C99 and a C
    program must link against the library by name."""
    import subprocess

    src = tmp_path / "use.c"
    src.write_text(
        '#include "kjarni_cuda.h"\n#include "kjarni_cuda_debug.h"\n#include "kjarni_ffi.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  KjarniEmbedderConfig c = kjarni_embedder_config_default();\n"
        "  KjcIndexDirInfo info; KjcForwardOptions o = {KJC_OUT_POOLED, KJC_POOL_MEAN, 1, KJC_MASK_AUTO};\n"
        "  (void)info; (void)o;\n"
        '  printf("%s|%s|%d|%d\\n", kjc_version(), kjarni_error_name(KJARNI_ERROR_GPU_UNAVAILABLE), (int)c.device, (int)c.normalize);\n'
        "  return kjc_index_dir_info(NULL, NULL) == KJC_NULL_POINTER ? 0 : 1;\n}\n")
    exe = tmp_path / "use"
    libdir = os.path.dirname(N.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lkjarni_cuda", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "sm_100a" in r.stdout and "GpuUnavailable|0|1" in r.stdout, (r.stdout, r.stderr)
