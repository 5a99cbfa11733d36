"""The Rust forwarding crate under shim/  cannot be compiled here (no
Rust toolchain); what can be checked is that its `extern "C"` block is exactly the header:
ffi.rs is generated from include/phnsw.h, and every phnsw_* call in the hand-written wrapper
names a declared function with the declared number of arguments."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_ffi_rs_is_the_header():
    import gen_rust_ffi
    text, names = gen_rust_ffi.generate()
    assert open(gen_rust_ffi.OUT).read() == text, "run python tools/gen_rust_ffi.py"
    lib = os.path.join(ROOT, "parallel_hnsw_b200", "libphnsw.so")
    if os.path.exists(lib):
        out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
        exported = {l.split()[-1] for l in out.splitlines() if " T phnsw_" in l}
        assert exported == set(names)


def _call_args(src, start):
    depth, i, args, cur = 0, start, [], ""
    while True:
        ch = src[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                if cur.strip():
                    args.append(cur)
                return args
        if ch == "," and depth == 1:
            args.append(cur)
            cur = ""
        elif not (ch == "(" and depth == 1 and not cur):
            cur += ch
        i += 1


def test_wrapper_calls_match_the_declarations():
    import gen_rust_ffi
    _, _, _, funcs = gen_rust_ffi.parse(open(gen_rust_ffi.HEADER).read())
    arity = {name: len(params) for _, name, params in funcs}
    src = open(os.path.join(ROOT, "shim", "src", "lib.rs")).read()
    src = re.sub(r"//.*", "", src)
    calls = list(re.finditer(r"\b(phnsw_\w+)\s*\(", src))
    assert len(calls) >= 20
    for m in calls:
        name = m.group(1)
        assert name in arity, name
        args = _call_args(src, m.end() - 1)
        assert len(args) == arity[name], (name, len(args), arity[name])
    assert src.count("{") == src.count("}") and src.count("(") == src.count(")")
