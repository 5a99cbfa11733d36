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


_RUST_SIZES = {"u64": (8, 8), "u32": (4, 4), "u16": (2, 2), "u8": (1, 1), "c_int": (4, 4), "f32": (4, 4),
               "f64": (8, 8), "usize": (8, 8)}


def _rust_structs():
    """{name: [(field, rust type)]} of every `#[repr(C)] pub struct` in the generated ffi.rs;
    a struct without repr(C) is a failure (Rust may reorder its fields)."""
    src = open(os.path.join(ROOT, "shim", "src", "ffi.rs")).read()
    out = {}
    for m in re.finditer(r"((?:#\[[^\]]*\]\s*)*)pub struct (\w+)\s*\{(.*?)\}", src, flags=re.S):
        attrs, name, body = m.group(1), m.group(2), m.group(3)
        fields = [(f.group(1), f.group(2).strip()) for f in re.finditer(r"pub (\w+):\s*([^,\n]+),", body)]
        if fields:  # opaque handles are `{ _private: [u8; 0] }`
            assert "#[repr(C)]" in attrs, "%s has fields but no #[repr(C)]" % name
            out[name] = fields
    return out


def _rust_layout(name, structs, memo):
    """(size, align, [(field, offset)]) under the repr(C) rules (natural alignment, declaration
    order) -- what rustc lays out for the generated declarations."""
    if name in memo:
        return memo[name]
    off, align, fields = 0, 1, []
    for fname, t in structs[name]:
        if t.startswith("*") or t.startswith("Option<") or t == "phnsw_progress_fn":
            sz, al = 8, 8
        elif t in _RUST_SIZES:
            sz, al = _RUST_SIZES[t]
        else:
            sz, al, _ = _rust_layout(t, structs, memo)
        off = (off + al - 1) // al * al
        fields.append((fname, off))
        off += sz
        align = max(align, al)
    size = (off + align - 1) // align * align
    memo[name] = (size, align, fields)
    return memo[name]


def test_struct_layouts_match_the_c_header(tmp_path):
    """Field names, order, offsets and sizes of every struct that crosses the ABI: the C side is
    asked (gcc, offsetof / sizeof on include/phnsw.h), the Rust side is computed from the
    #[repr(C)] declarations in ffi.rs."""
    import gen_rust_ffi
    _, c_structs, _, _ = gen_rust_ffi.parse(open(gen_rust_ffi.HEADER).read())
    rs = _rust_structs()
    assert {n for n, _ in c_structs} == set(rs), (sorted(n for n, _ in c_structs), sorted(rs))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "phnsw.h"', "int main(void) {"]
    for name, fields in c_structs:
        assert [f for _, f in fields] == [f for f, _ in rs[name]], name   # names and order
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (name, name))
        for _, f in fields:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    c_out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    want = {}
    for ln in c_out:
        p = ln.split()
        if len(p) == 3:
            want[(p[0], p[1])] = int(p[2])
    memo = {}
    for name in rs:
        size, _, fields = _rust_layout(name, rs, memo)
        assert want[(name, "size")] == size, (name, want[(name, "size")], size)
        for f, off in fields:
            assert want[(name, f)] == off, (name, f, want[(name, f)], off)
