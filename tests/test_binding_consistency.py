"""The bindings that cannot be compiled or that are hand-written against the header — the Rust shim (no Rust toolchain here) and the
ctypes table in hotstuff_b200/_lib.py — are checked against include/hs_crypto.h: same functions, same parameter count, compatible
parameter types, same struct sizes.  CPU only."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def header_functions():
    """name -> (return type, [parameter types]) with names and array suffixes removed ('const uint8_t *', 'size_t', ...)."""
    src = _strip_comments(open(os.path.join(ROOT, "include", "hs_crypto.h")).read())
    out = {}
    for m in re.finditer(r"\b([A-Za-z_][\w ]*?[\s\*]+)(hs_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3)
        types = []
        for p in [q.strip() for q in params.split(",") if q.strip() and q.strip() != "void"]:
            arr = "[" in p
            p = re.sub(r"\[.*?\]", "", p).strip()
            mm = re.match(r"^(.*?)(\w+)$", p)  # drop the parameter name
            t = (mm.group(1) if mm else p).strip()
            if arr:
                t += " *"
            types.append(re.sub(r"\s+", " ", t).replace(" *", "*").strip())
        out[name] = (re.sub(r"\s+", " ", ret).replace(" *", "*"), types)
    return out


RUST_TO_C = {
    "c_int": "int", "u32": "uint32_t", "u64": "uint64_t", "usize": "size_t",
    "*const u8": "const uint8_t*", "*mut u8": "uint8_t*", "*const u32": "const uint32_t*", "*mut u32": "uint32_t*",
    "*const u64": "const uint64_t*", "*mut c_int": "int*", "*mut HsCtx": "hs_ctx*", "*const HsCtx": "const hs_ctx*",
    "*mut *mut HsCtx": "hs_ctx**", "*const HsRec128": "const hs_rec128*", "*const HsVote": "const hs_vote*",
    "*const std::os::raw::c_char": "const char*", "*mut HsFrameInfo": "hs_frame_info*", "*mut HsIngestOut": "hs_ingest_out*",
}
RUST_SIZES = {"u8": 1, "u32": 4, "u64": 8, "usize": 8, "*mut u8": 8, "*mut u32": 8, "*mut u64": 8}
C_SIZES = {"uint8_t": 1, "uint32_t": 4, "uint64_t": 8, "size_t": 8, "uint8_t*": 8, "uint32_t*": 8, "uint64_t*": 8}


def _c_struct_fields(hdr, name):
    body = re.search(r"typedef struct\s*(?:\w+\s*)?\{([^}]*)\}\s*%s\s*;" % name, hdr).group(1)
    fields = []
    for decl in [d.strip() for d in body.split(";") if d.strip()]:
        m = re.match(r"^(\w+)\s+(.*)$", decl)
        for var in m.group(2).split(","):
            var = var.strip()
            fields.append((var.lstrip("*"), m.group(1) + ("*" if var.startswith("*") else "")))
    return fields


def _rust_struct_fields(src, name):
    body = re.search(r"pub struct %s\s*\{(.*?)\}" % name, src, flags=re.S).group(1)
    return [(m.group(1), re.sub(r"\s+", " ", m.group(2).strip())) for m in re.finditer(r"pub (\w+):\s*([^,}]+)", body)]


def test_rust_shim_ingest_structs_match_the_header_field_by_field():
    src = _strip_comments(open(os.path.join(ROOT, "rust", "crypto_gpu_shim.rs")).read())
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "hs_crypto.h")).read())
    for rust_name, c_name, size in (("HsFrameInfo", "hs_frame_info", 48), ("HsIngestOut", "hs_ingest_out", 13 * 8)):
        rf, cf = _rust_struct_fields(src, rust_name), _c_struct_fields(hdr, c_name)
        assert [f for f, _ in rf] == [f for f, _ in cf], (rust_name, rf, cf)
        assert [RUST_SIZES[t] for _, t in rf] == [C_SIZES[t] for _, t in cf], (rust_name, rf, cf)
        assert sum(RUST_SIZES[t] for _, t in rf) == size       # no implicit padding: fields are laid out in non-increasing alignment groups
    assert re.search(r"#define HS_FRAME_MALFORMED 255", hdr) and "pub const HS_FRAME_MALFORMED: u8 = 255;" in src
    assert re.search(r"#define HS_ERR_NOMEM 3", hdr) and "pub const HS_ERR_NOMEM: c_int = 3;" in src


def test_header_parser_sees_the_whole_abi():
    fns = header_functions()
    assert len(fns) >= 43 and "hs_verify_strict_batch" in fns and "hs_ingest_consensus_frames" in fns
    assert fns["hs_verify_strict_batch"] == ("int", ["hs_ctx*", "const hs_rec128*", "size_t", "uint32_t*"])
    assert fns["hs_verify_batch_shared_msg"][1][1] == "const uint8_t*"   # 'const uint8_t digest[32]'


def test_rust_shim_extern_block_matches_the_header():
    src = _strip_comments(open(os.path.join(ROOT, "rust", "crypto_gpu_shim.rs")).read())
    block = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S).group(1)
    fns = header_functions()
    seen = 0
    for m in re.finditer(r"fn\s+(hs_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, params, ret = m.group(1), m.group(2), (m.group(3) or "").strip()
        assert name in fns, "%s is not declared in include/hs_crypto.h" % name
        c_ret, c_types = fns[name]
        r_types = [re.sub(r"\s+", " ", p.split(":", 1)[1].strip()) for p in params.split(",") if p.strip()]
        assert len(r_types) == len(c_types), "%s: %d parameters in the shim, %d in the header" % (name, len(r_types), len(c_types))
        for k, (r, c) in enumerate(zip(r_types, c_types)):
            assert r in RUST_TO_C, "%s: unmapped Rust type %r" % (name, r)
            assert RUST_TO_C[r] == c, "%s parameter %d: shim %r vs header %r" % (name, k, r, c)
        assert RUST_TO_C[ret] == c_ret, "%s: return type" % name
        seen += 1
    assert seen >= 11
    # #[repr(C)] structs: field sizes add up to the C structs' sizes
    assert "pub struct HsRec128 { pub sig: [u8; 64], pub pk: [u8; 32], pub msg: [u8; 32] }" in src      # hs_rec128: 128 bytes
    assert "pub struct HsVote   { pub pk: [u8; 32], pub sig: [u8; 64] }" in src                          # hs_vote: 96 bytes
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "hs_crypto.h")).read())
    rec = re.search(r"typedef struct\s*(?:\w+\s*)?\{([^}]*)\}\s*hs_rec128\s*;", hdr).group(1)
    assert [int(x) for x in re.findall(r"\[(\d+)\]", rec)] == [64, 32, 32]
    vote = re.search(r"typedef struct\s*(?:\w+\s*)?\{([^}]*)\}\s*hs_vote\s*;", hdr).group(1)
    assert [int(x) for x in re.findall(r"\[(\d+)\]", vote)] == [32, 64]


def test_rust_shim_calls_only_what_it_declares_and_uses_the_cutover():
    src = _strip_comments(open(os.path.join(ROOT, "rust", "crypto_gpu_shim.rs")).read())
    block = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S).group(1)
    declared = set(re.findall(r"fn\s+(hs_\w+)", block))
    called = set(re.findall(r"\b(hs_\w+)\s*\(", src.replace(block, "")))
    assert called <= declared, called - declared
    assert declared - called == set(), "declared but never called: %s" % (declared - called)
    # every wrapper that returns verdicts treats rc != HS_OK as a rejection or an error, never as an accept
    for body in re.findall(r"pub fn \w+.*?\n\}", src, flags=re.S):
        if "unsafe { hs_" in body and "hs_ctx_create" not in body:
            assert "HS_OK" in body, body[:80]
    assert re.search(r"recs\.len\(\) < GPU_MIN_SIGS \{ return None; \}", src) and re.search(r"msgs\.len\(\) < GPU_MIN_DIGEST_MSGS", src)


C_TO_CTYPES = {
    "int": (ctypes.c_int,), "uint32_t": (ctypes.c_uint32,), "size_t": (ctypes.c_size_t,), "double": (ctypes.c_double,),
}


def test_ctypes_table_matches_the_header():
    """hotstuff_b200/_lib.py declares argtypes by hand: the count and the scalar / pointer kind of every parameter must match the header."""
    from hotstuff_b200 import _lib
    lib = _lib.load()
    fns = header_functions()
    checked = 0
    for name, (c_ret, c_types) in fns.items():
        f = getattr(lib, name)
        if f.argtypes is None:
            continue
        assert len(f.argtypes) == len(c_types), "%s: %d argtypes, header has %d parameters" % (name, len(f.argtypes), len(c_types))
        for k, (a, c) in enumerate(zip(f.argtypes, c_types)):
            if c.endswith("*"):
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, ctypes._Pointer), "%s parameter %d: header %s, ctypes %r" % (name, k, c, a)
            else:
                assert a in C_TO_CTYPES[c], "%s parameter %d: header %s, ctypes %r" % (name, k, c, a)
        checked += 1
    assert checked >= 40, checked
