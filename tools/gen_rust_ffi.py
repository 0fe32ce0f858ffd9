#!/usr/bin/env python3
"""Regenerate the `extern "C"` block of rust/anemoi_b200_shim.rs from include/anemoi_b200.h (between the
`// BEGIN GENERATED FFI` / `// END GENERATED FFI` markers). tests/test_rust_shim_abi.py parses both files independently
and fails on any drift, so after touching the header run:  python tools/gen_rust_ffi.py"""
import os
import re

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
C_TO_RUST = {"int": "c_int", "size_t": "usize", "uint64_t": "u64", "uint8_t": "u8", "char": "c_char", "void": "c_void",
             "double": "f64"}
RENAME = {"in": "input"}  # Rust keywords


def parse_header(text):
    """[(name, [(arg_name, c_type)], c_return_type)] for every anemoi_b200_* prototype."""
    src = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for ret, name, args in re.findall(r"((?:const\s+)?\w+\s*\**)\s*(anemoi_b200_\w+)\s*\(([^)]*)\)\s*;", src):
        params = []
        args = args.strip()
        if args not in ("", "void"):
            for a in args.split(","):
                a = " ".join(a.split())
                m = re.match(r"(.*?)(\w+)$", a)
                params.append((m.group(2), " ".join(m.group(1).split())))
        out.append((name, params, " ".join(ret.split())))
    return out


def rust_type(c_type):
    t = c_type.replace(" *", "*").strip()
    const = t.startswith("const ")
    base = t.replace("const ", "")
    stars = base.count("*")
    prim = C_TO_RUST[base.replace("*", "").strip()]
    for level in range(stars):
        prim = ("*const " if (const and level == 0) else "*mut ") + prim
    return prim


def extern_block(header_text):
    lines = []
    for name, params, ret in parse_header(header_text):
        args = ", ".join("%s: %s" % (RENAME.get(n, n), rust_type(t)) for n, t in params)
        lines.append("    pub fn %s(%s) -> %s;" % (name, args, rust_type(ret)))
    return "\n".join(lines)


def main():
    header = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    path = os.path.join(ROOT, "rust", "anemoi_b200_shim.rs")
    text = open(path).read()
    begin, end = "    // BEGIN GENERATED FFI (tools/gen_rust_ffi.py)\n", "    // END GENERATED FFI\n"
    i, j = text.index(begin) + len(begin), text.index(end)
    text = text[:i] + extern_block(header) + "\n" + text[j:]
    with open(path, "w") as f:
        f.write(text)
    print("wrote", path)


if __name__ == "__main__":
    main()
