#!/usr/bin/env python
"""Generates rust/dunk-b200-sys/src/lib.rs — the raw `extern "C"` binding of EVERY symbol, struct, enum and constant
that include/dunk_b200.h declares (so the sys crate cannot drift from the header; tests/test_rust_shim.py checks it).
    python tools/gen_rust_sys.py            # rewrite the file
    python tools/gen_rust_sys.py --check    # exit 1 if the committed file is stale"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dunk_b200.h")
OUT = os.path.join(ROOT, "rust", "dunk-b200-sys", "src", "lib.rs")

SCALARS = {"int": "c_int", "int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8",
           "size_t": "usize", "float": "f32", "double": "f64", "char": "c_char", "void": "c_void"}
OPAQUE = {"dunk_ctx": "DunkCtx", "dunk_db": "DunkDb", "dunk_elevation": "DunkElevation", "dunk_shard_group": "DunkShardGroup"}


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def rust_type(ctype):
    """'const uint8_t*' -> '*const u8'; 'dunk_ctx**' -> '*mut *mut DunkCtx'"""
    t = ctype.strip()
    const = t.startswith("const ")
    t = t[6:].strip() if const else t
    stars = t.count("*")
    base = t.replace("*", "").strip()
    base = base[7:].strip() if base.startswith("struct ") else base
    r = SCALARS.get(base) or OPAQUE.get(base) or base            # Dunk* structs keep their names
    for i in range(stars):
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


def parse(src):
    src = strip_comments(src)
    consts = re.findall(r"#define\s+(DUNK_[A-Z0-9_]+)\s+(\(?-?[0-9][^\n]*|\(\(1 << [A-Z_]+\) - 1\))", src)
    enums = []
    for name, body in re.findall(r"enum\s+(\w+)\s*\{([^}]*)\}", src):
        enums.append((name, [(k.strip(), int(v)) for k, v in re.findall(r"(\w+)\s*=\s*(-?\d+)", body)]))
    structs = []
    for body, name in re.findall(r"typedef struct \w+\s*\{([^}]*)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"^((?:const\s+)?(?:struct\s+)?\w+\s*\**)\s*(.*)$", decl)
            ctype, names = m.group(1), m.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                if arr:
                    fields.append((arr.group(1), f"[{rust_type(ctype)}; {arr.group(2)}]"))
                else:
                    fields.append((nm.lstrip("* "), rust_type(ctype + "*" * nm.count("*"))))
        structs.append((name, fields))
    funcs = []
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[\w]+\s*\**)\s*(dunk_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        ret = " ".join(ret.split())
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.+?)(\w+)$", a)
                params.append((m.group(2), rust_type(m.group(1))))
        funcs.append((name, params, None if ret == "void" else rust_type(ret)))
    return consts, enums, structs, funcs


RESERVED = {"type", "ref", "in", "fn", "mod", "use", "loop", "match", "box", "impl"}


def emit():
    consts, enums, structs, funcs = parse(open(HEADER).read())
    o = ["//! Raw FFI of `libdunk_b200.so` — GENERATED from include/dunk_b200.h by tools/gen_rust_sys.py; do not edit.",
         "//! One item per declaration of the header: every `extern \"C\"` entry point, `#[repr(C)]` struct, enum value and",
         "//! constant.  The safe wrappers live in the `feature_extraction`, `homographier` and `feature_database` shim crates.",
         "#![allow(non_camel_case_types, non_upper_case_globals, non_snake_case, clippy::too_many_arguments)]",
         "use std::os::raw::{c_char, c_int, c_void};", ""]
    for name, val in consts:
        v = val.strip()
        if "<<" in v:
            v = "(1 << DUNK_MAX_POINTS_SHIFT) - 1"
        o.append(f"pub const {name}: c_int = {v.strip('()') if v.startswith('(-') else v};")
    o.append("")
    for name, items in enums:
        o.append(f"/// `enum {name}`")
        for k, v in items:
            o.append(f"pub const {k}: c_int = {v};")
        o.append("")
    for rs in OPAQUE.values():
        o.append(f"#[repr(C)] pub struct {rs} {{ _private: [u8; 0] }}")
    o.append("")
    for name, fields in structs:
        o.append("#[repr(C)]\n#[derive(Clone, Copy, Debug)]")
        o.append(f"pub struct {name} {{")
        for f, t in fields:
            f = "r#" + f if f in RESERVED else f
            o.append(f"    pub {f}: {t},")
        o.append("}\n")
    o.append('#[link(name = "dunk_b200")]\nextern "C" {')
    for name, params, ret in funcs:
        ps = ", ".join(f"{('r#' + p) if p in RESERVED else p}: {t}" for p, t in params)
        o.append(f"    pub fn {name}({ps})" + (f" -> {ret}" if ret else "") + ";")
    o.append("}\n")
    o.append(HELPERS)
    return "\n".join(o), [f[0] for f in funcs]


HELPERS = '''/// message of the calling thread's last failed call (`dunk_last_error`, thread-local in the library)
pub fn last_error() -> String {
    // SAFETY: the library returns a NUL-terminated string that stays valid until the thread's next failing call
    unsafe { std::ffi::CStr::from_ptr(dunk_last_error()) }.to_string_lossy().into_owned()
}

/// Process-wide context on the GPU named by `DUNK_DEVICE` (default 0) with 8 stream / workspace slots: the reference's
/// callers are rayon workers (preprocessor/src/main.rs:233-243), concurrent calls take distinct slots.
/// There is no CPU fallback: without an sm_100 device this panics with the library's message.
pub fn ctx() -> *mut DunkCtx {
    use std::sync::OnceLock;
    struct P(*mut DunkCtx);
    // SAFETY: the context is internally synchronised (slot pool under a mutex); the pointer itself is immutable
    unsafe impl Send for P {}
    unsafe impl Sync for P {}
    static CTX: OnceLock<P> = OnceLock::new();
    CTX.get_or_init(|| {
        let device = std::env::var("DUNK_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut c = std::ptr::null_mut();
        // SAFETY: `c` is a valid out pointer
        let rc = unsafe { dunk_ctx_create(device, 8, &mut c) };
        assert_eq!(rc, 0, "dunk_ctx_create failed ({rc}): {}", last_error());
        P(c)
    })
    .0
}
'''


if __name__ == "__main__":
    text, names = emit()
    if "--check" in sys.argv:
        ok = os.path.exists(OUT) and open(OUT).read() == text
        print("up to date" if ok else "STALE: run python tools/gen_rust_sys.py")
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: {len(names)} functions")
