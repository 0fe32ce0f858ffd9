"""The Rust shim (rust/anemoi_b200_shim.rs) cannot be compiled here (no cargo), so its `extern "C"` block is checked
textually: every function include/anemoi_b200.h declares must be bound, with the same number of arguments and with
Rust types that are ABI-equivalent to the C ones; nothing else may be bound. Any drift between the two files fails."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ABI = {  # C type (spaces normalised) -> the Rust FFI type it must be bound as
    "int": "c_int", "size_t": "usize", "double": "f64",
    "const char*": "*const c_char", "void*": "*mut c_void", "void**": "*mut *mut c_void",
    "uint64_t*": "*mut u64", "const uint64_t*": "*const u64", "uint8_t*": "*mut u8", "const uint8_t*": "*const u8",
    "int*": "*mut c_int", "double*": "*mut f64",
}


def c_prototypes():
    src = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"((?:const\s+)?\w+\s*\**)\s*(anemoi_b200_\w+)\s*\(([^)]*)\)\s*;", src):
        params = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = " ".join(a.split())
                ty = re.match(r"(.*?)(\w+)$", a).group(1)
                params.append(re.sub(r"\s*\*", "*", ty).strip())
        protos[name] = (params, re.sub(r"\s*\*", "*", " ".join(ret.split())).strip())
    return protos


def rust_bindings():
    src = open(os.path.join(ROOT, "rust", "anemoi_b200_shim.rs")).read()
    block = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    out = {}
    for name, args, ret in re.findall(r"fn\s+(\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        params = []
        for a in [x for x in args.split(",") if x.strip()]:
            params.append(" ".join(a.split(":", 1)[1].split()))
        out[name] = (params, " ".join((ret or "()").split()))
    return out


def test_extern_block_matches_header():
    c, r = c_prototypes(), rust_bindings()
    assert len(c) >= 45
    assert sorted(c) == sorted(r), "bound but not declared: %s; declared but not bound: %s" % (
        sorted(set(r) - set(c)), sorted(set(c) - set(r)))
    for name, (params, ret) in c.items():
        rp, rr = r[name]
        assert len(rp) == len(params), "%s: arity %d in the header, %d in Rust" % (name, len(params), len(rp))
        for i, (ct, rt) in enumerate(zip(params, rp)):
            assert ABI[ct] == rt, "%s arg %d: C `%s` must be bound as `%s`, found `%s`" % (name, i, ct, ABI[ct], rt)
        assert ABI[ret] == rr, "%s: return C `%s` must be `%s`, found `%s`" % (name, ret, ABI[ret], rr)


def test_generator_is_idempotent():
    """tools/gen_rust_ffi.py reproduces the committed block (the block was not hand-edited)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_ffi

    header = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    text = open(os.path.join(ROOT, "rust", "anemoi_b200_shim.rs")).read()
    assert gen_rust_ffi.extern_block(header) in text


def test_shim_surface_mirrors_the_reference_traits():
    """the batched trait speaks the reference's types -- Digest in/out where `Sponge` does
    (src/traits.rs:8-20, digest.rs:13-53), ragged `&[&[F]]`, and one impl per reference marker type."""
    src = open(os.path.join(ROOT, "rust", "anemoi_b200_shim.rs")).read()
    for sig in ("fn hash_field_batch(msgs: &[&[F]]) -> Vec<Self::Digest>", "fn hash_batch(msgs: &[&[u8]]) -> Vec<Self::Digest>",
                "fn merge_batch(pairs: &[[Self::Digest; 2]]) -> Vec<Self::Digest>", "fn compress_k_batch(elems: &[F], k: usize) -> Vec<F>",
                "fn compress_batch(elems: &[F]) -> Vec<F>", "fn permutation_batch(states: &mut [F])",
                "fn merkle_root(leaves: &[Self::Digest], n_gpus: usize) -> Self::Digest", "fn digests_to_bytes(digests: &[Self::Digest]) -> Vec<u8>",
                "fn merkle_open(", "fn merkle_verify(", "pub unsafe fn merkle_root_sharded<"):
        assert sig in src, sig
    impls = re.findall(r"impl_b200!\(crate::(\w+)::anemoi_(\d_\d), (\w+),", src)
    assert len(impls) == 14
    import json

    params = json.load(open(os.path.join(ROOT, "tests", "golden", "params.json")))
    assert sorted({f for f, _, _ in impls}) == sorted(params)
