#!/usr/bin/env python
"""Writes tests/golden/ffi_symbols.json: the NAMES of the C-ABI exports of the reference's kjarni-ffi crate and of the symbols each
of its bindings resolves (Python ctypes, Go purego, C# P/Invoke).  Run in the build container (reads /root/reference):
    python scripts/gen_ffi_symbols.py
tests/test_abi_cpu.py checks that libkjarni_ffi.so exports every one of them, so that a binding which resolves its symbols eagerly
loads against this library."""
import glob
import json
import os
import re

REF = "/root/reference/crates/kjarni-ffi"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rust = set()
    for f in glob.glob(os.path.join(REF, "src", "*.rs")):
        rust |= set(re.findall(r'pub\s+(?:unsafe\s+)?extern\s+"C"\s+fn\s+(kjarni_\w+)', open(f).read()))
    py = set(re.findall(r"_lib\.(kjarni_\w+)", open(os.path.join(REF, "bindings/python/kjarni/_ffi.py")).read()))
    go = set(re.findall(r'"(kjarni_\w+)"', open(os.path.join(REF, "bindings/go/ffi.go")).read()))
    cs = set(re.findall(r'\b(kjarni_\w+)\s*\(', open(os.path.join(REF, "bindings/csharp/Kjarni/Native.cs")).read()))
    out = {"source": "olafurjohannsson/kjarni crates/kjarni-ffi (src/*.rs, bindings/{python,go,csharp})",
           "rust_exports": sorted(rust), "python_binding": sorted(py), "go_binding": sorted(go), "csharp_binding": sorted(cs & rust)}
    with open(os.path.join(ROOT, "tests", "golden", "ffi_symbols.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
