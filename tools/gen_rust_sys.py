#!/usr/bin/env python
"""Writes rust/ox_b200-sys/src/lib.rs from include/ox_b200.h: constants, opaque handles, repr(C) structs and one `extern "C"`
declaration per OX_API prototype - a mechanical transcription, so the FFI crate cannot drift from the header
(tests/test_rust_sources.py regenerates it and compares). The image has no Rust toolchain: the output is never compiled here."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TMAP = {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "double": "c_double", "void": "c_void", "char": "c_char",
        "ox_status": "ox_status", "ox_model_tables": "c_void"}


def rtype(t):
    t = t.replace("const ", "").strip()
    if t == "void":
        return None
    if t.endswith("*"):
        base = t[:-1].strip()
        if base == "char":
            return "*const c_char"
        if base == "ox_model_tables":
            return "*const c_void"
        return "*mut " + TMAP.get(base, base)
    return TMAP.get(t, t)


def argtype(a):
    a = a.strip()
    if a in ("", "void"):
        return None
    m = re.match(r"(.*?)([A-Za-z_0-9]+)$", a)
    t, name = m.group(1).strip(), m.group(2)
    const = t.startswith("const ")
    core = t.replace("const", "").strip()
    stars, base = core.count("*"), core.replace("*", "").strip()
    rb = TMAP.get(base, base)
    if stars == 0:
        ty = rb
    elif stars == 1:
        ty = ("*const " if const else "*mut ") + rb
    elif base == "void" and "const*" in t.replace(" ", ""):
        ty = "*const *mut c_void"
    else:
        ty = "*mut " + ("*const " if const else "*mut ") + rb
    if name in ("in", "type", "ref", "box"):
        name += "_"
    return f"{name}: {ty}"


def generate():
    hdr = open(os.path.join(ROOT, "include", "ox_b200.h")).read()
    out = ["//! Raw bindings to libox_b200.so - GENERATED from include/ox_b200.h by tools/gen_rust_sys.py, do not edit.",
           "//! NOT COMPILED in the build environment (no cargo / rustc there); tests/test_rust_sources.py keeps it in sync with the header.",
           "#![allow(non_camel_case_types, non_upper_case_globals)]", "use std::os::raw::{c_char, c_double, c_void};", "",
           "pub type ox_status = i32;"]
    # anonymous enums: every NAME = value (values may reference 1 << n)
    for body in re.findall(r"enum\s*\{(.*?)\}\s*;", hdr, re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [x.strip() for x in item.split("=")]
                val = eval(expr)  # integer literals and shifts only
            else:
                name, val = item, val + 1
            out.append(f"pub const {name}: i32 = {val};")
    out += ["", "pub const OX_MAXVAL: f64 = 1e10;", "pub const OX_MINVAL: f64 = 1e-15;", ""]
    for h in ("ox_model", "ox_batch", "ox_env", "ox_group"):
        out.append(f"#[repr(C)] pub struct {h} {{ _private: [u8; 0] }}")
    # plain structs of the ABI
    for name, body in re.findall(r"typedef struct (ox_[a-z_]+) \{(.*?)\} \1;", hdr, re.S):
        if name == "ox_model_tables":
            continue  # read through ox_model_int_table / ox_model_real_table / ox_model_size
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?([a-z_0-9]+)(\s*\*)?\s+(.*)$", decl)
            const, base, ptr, names = m.group(1), m.group(2), m.group(3), m.group(4)
            for nm in [x.strip() for x in names.split(",")]:
                ty = TMAP.get(base, base)
                if ptr:
                    ty = ("*const " if const else "*mut ") + ty
                fields.append(f"pub {nm}: {ty}")
        out.append(f"#[repr(C)] #[derive(Clone, Copy)] pub struct {name} {{ {', '.join(fields)} }}")
    out += ["", '#[link(name = "ox_b200")]', 'extern "C" {']
    for ret, name, args in re.findall(r"OX_API\s+([^;]+?)\s*\b(ox_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        al = [argtype(a) for a in re.sub(r"/\*.*?\*/", "", args, flags=re.S).split(",")]
        r = rtype(ret)
        out.append(f"    pub fn {name}({', '.join(a for a in al if a)})" + (f" -> {r}" if r else "") + ";")
    out.append("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = generate()
    path = os.path.join(ROOT, "rust", "ox_b200-sys", "src", "lib.rs")
    if len(sys.argv) > 1 and sys.argv[1] == "--check":
        sys.exit(0 if open(path).read() == text else 1)
    open(path, "w").write(text)
    print(f"wrote {path}: {text.count('pub fn ')} functions, {text.count('pub const ')} constants")
