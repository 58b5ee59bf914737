#!/usr/bin/env python3
"""glsl2cpp.py <shader dir> <shader file> <driver.inc> <out.cpp>

Turns ONE of the reference's GLSL shaders — read from where it lies under /root/reference/shaders — into a C++20
translation unit that compiles against oracle/ref_rig/glsl_shim.h.  The output goes to oracle/_ref/ only (git-ignored);
no reference source is stored in the repository.  TEST INFRASTRUCTURE (see glsl_shim.h).

The shader BODY is not rewritten: every expression, branch, constant and statement order is the reference's.  Only
the parts of GLSL that are not C++ are mapped, mechanically:
  * `#pragma include "x"` is resolved like GLHelper::preprocessShader does (src/Graphics/GLHelper.cpp:265-283), `#version` /
    `#extension` lines are dropped and the text is run through the C preprocessor (g++ -E), so `#if USE_RGBA16F`, `#ifdef
    NORMAL_MAP`, `#define POINT_LIGHT 0` … select and expand exactly what a GLSL compiler would see;
  * interface declarations become namespace-level variables the driver sets (uniforms shared, `in`/`out` and built-ins
    thread_local: the drivers run the invocations of a dispatch on all host threads): `layout(...) uniform T x [= v];`, `in`/`out`
    variables, `in NAME {...} inst;`, `layout(...) buffer/uniform NAME {...};` (an unsized `T a[];` member becomes
    ssbo_array<T>), `layout(local_size...) in;` is dropped;
  * array constructors `T[](a, b, c)` -> `{a, b, c}`;  `x.length()` -> glsl_length(x);
  * `out` / `inout` parameters -> references; `discard` -> throw Discard(); `flat`, and `layout(..) coherent volatile` on image
    parameters, are dropped;
  * unsized per-vertex arrays of tessellation stages get the patch size 3; `T x = x(...)` calls the function x (GLSL scoping);
  * floating literals get an `f` suffix (GLSL literals are fp32; C++ would evaluate in double);
  * `imageLoad` on a `uimage3D` / `iimage3D` -> imageLoadU / imageLoadI (uvec4 / ivec4);  `void main()` -> `void shader_main()`.
"""
import os
import re
import subprocess
import sys


def resolve_includes(shader_dir, name, depth=0):
    out = []
    for line in open(os.path.join(shader_dir, name), encoding="utf-8", errors="replace").read().splitlines():
        m = re.match(r'\s*#\s*pragma\s+include\s+"([^"]+)"', line)
        if m and depth < 8:
            out.extend(resolve_includes(shader_dir, m.group(1), depth + 1))
        elif re.match(r"\s*#\s*(version|extension)\b", line):
            continue
        else:
            out.append(line)
    return out


def match_paren(text, open_at):
    depth = 0
    for i in range(open_at, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced parenthesis")


TYPES = r"(?:float|int|uint|bool|vec[234]|ivec[234]|uvec[234]|mat[34])"


def array_constructors(text):
    pat = re.compile(TYPES + r"\s*\[\s*\w*\s*\]\s*\(")
    while True:
        m = pat.search(text)
        if not m:
            return text
        close = match_paren(text, m.end() - 1)
        text = text[:m.start()] + "{" + text[m.end():close] + "}" + text[close + 1:]


def interface_blocks(text):
    def block(m):
        body = m.group("body")
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            a = re.match(r"(\w+)\s+(\w+)\s*\[\s*\]$", decl)
            out.append(f"static ssbo_array<{a.group(1)}> {a.group(2)};" if a else f"static {decl};")
        return "\n".join(out)
    text = re.sub(r"(?:layout\s*\([^)]*\)\s*)?(?:buffer|uniform)\s+\w+\s*\{(?P<body>[^}]*)\}\s*;", block, text)
    text = re.sub(r"(?:flat\s+)?\b(?:in|out)\s+(\w+)\s*\{([^}]*)\}\s*(\w+)\s*\[\s*\]\s*;", r"static thread_local struct \1 {\2} \3[3];", text)   # gs_in[]
    text = re.sub(r"(?:flat\s+)?\b(?:in|out)\s+(\w+)\s*\{([^}]*)\}\s*(\w+)\s*;", r"static thread_local struct \1 {\2} \3;", text)
    return text


def global_declarations(text):
    out, depth = [], 0
    qual = re.compile(r"^\s*(?:layout\s*\([^)]*\)\s*)?((?:(?:uniform|readonly|writeonly|coherent|restrict|flat|smooth|in|out)\s+)+)")
    for line in text.splitlines():
        if depth == 0:
            if re.match(r"^\s*layout\s*\([^)]*\)\s*(?:in|out)\s*;", line):
                line = ""
            else:
                m = qual.match(line)
                if m:
                    # uniforms are shared by all invocations; `in` / `out` variables are per-invocation state (one copy per host thread)
                    per_invocation = re.search(r"\b(?:in|out)\b", m.group(1)) is not None
                    line = ("static thread_local " if per_invocation else "static ") + line[m.end():]
                    line = re.sub(r"\[\s*\]\s*;", "[3];", line)                  # per-vertex arrays of a 3-vertex patch
        depth += line.count("{") - line.count("}")
        out.append(line)
    return "\n".join(out)


def float_suffix(text):
    lit = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
    return lit.sub(lambda m: m.group(1) + "f", text)


def translate(shader_dir, name, ns):
    src = "\n".join(resolve_includes(shader_dir, name)) + "\n"
    pre = subprocess.run(["g++", "-E", "-P", "-undef", "-x", "c++", "-"], input=src, capture_output=True, text=True, check=True).stdout
    pre = re.sub(r"\bflat\s+", "", pre)                                            # interpolation qualifier inside interface blocks
    t = interface_blocks(pre)
    t = global_declarations(t)
    t = array_constructors(t)
    # `T name = name(...)`: in GLSL the variable is not yet in scope in its own initialiser and the call finds the function
    t = re.sub(r"\b(\w+)\s+(\w+)\s*=\s*\2\s*\(", lambda m: f"{m.group(1)} {m.group(2)} = {ns}::{m.group(2)}(", t)
    t = re.sub(r"\b(\w+)\.length\(\)", r"glsl_length(\1)", t)
    t = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", t)
    t = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", t)
    t = re.sub(r"\blayout\s*\([^)]*\)\s*", "", t)                                   # what is left: qualifiers on image parameters
    t = re.sub(r"\b(?:coherent|volatile)\s+", "", t)
    t = re.sub(r"\bdiscard\s*;", "throw Discard();", t)
    t = float_suffix(t)
    for u in re.findall(r"\buimage3D\s+(\w+)\s*;", t):
        t = re.sub(r"\bimageLoad\s*\(\s*" + u + r"\b", "imageLoadU(" + u, t)
    for u in re.findall(r"\biimage3D\s+(\w+)\s*;", t):
        t = re.sub(r"\bimageLoad\s*\(\s*" + u + r"\b", "imageLoadI(" + u, t)
    t = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", t)
    return t


def main():
    shader_dir, name, driver, out = sys.argv[1:5]
    ns = re.sub(r"\W", "_", name)
    body = translate(shader_dir, name, ns)
    # a driver that re-binds some uniforms per invocation (texture units, the material block) lists them:
    #   // glsl2cpp: per-invocation diffuseMap material ...
    for names in re.findall(r"//\s*glsl2cpp:\s*per-invocation\s+([^\n]*)", open(driver).read()):
        for n in names.split():
            body, k = re.subn(r"^static (?!thread_local)([^;=\n]*\b" + n + r"\b)", r"static thread_local \1", body, flags=re.M)
            if k != 1:
                raise SystemExit(f"glsl2cpp: per-invocation name {n} matched {k} declarations in {name}")
    with open(out, "w") as f:
        f.write(f"// GENERATED by oracle/ref_rig/glsl2cpp.py from {os.path.join(shader_dir, name)} — do not commit.\n")
        f.write('#include "glsl_shim.h"\n#include <vector>\n#include <cstdio>\n')
        f.write("namespace glsl {\nnamespace " + ns + " {\n")
        f.write("// built-in per-invocation variables: one copy per host thread (the drivers run invocations under OpenMP)\n"
                "static thread_local ivec3 gl_GlobalInvocationID;   // uvec3 in GLSL; ids stay far below 2^31\n"
                "static thread_local vec4 gl_FragCoord; static thread_local int gl_Layer;\n"
                "static thread_local int gl_InvocationID; static thread_local float gl_TessLevelInner[2], gl_TessLevelOuter[4];\n"
                "static thread_local vec3 gl_TessCoord; static thread_local vec4 gl_Position; static thread_local int gl_InstanceID;\n"
                "struct gl_PerVertex { vec4 gl_Position; }; static thread_local gl_PerVertex gl_in[3];\n"
                "static void EmitVertex(); static void EndPrimitive();   // geometry stage: defined by the driver\n")
        f.write(body)
        f.write("\n// ---- driver (this repository's code)\n")
        f.write(open(driver).read())
        f.write("\n}  // namespace " + ns + "\n}  // namespace glsl\n")


if __name__ == "__main__":
    main()
