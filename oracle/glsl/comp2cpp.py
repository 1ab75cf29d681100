#!/usr/bin/env python3
"""ORACLE — test infrastructure.  turns one of the reference's compute shaders into a C++ translation unit for glsl_shim.h.

    comp2cpp.py <reference>/src/pipe/modules <module> <kernel> > _ref/shader_<module>_<kernel>.cpp

The shader is read where it lies under /root/reference; the output goes to oracle/_ref/ (git-ignored), never into the repo.
What happens to the source: `#version` / `#extension` lines dropped, includes and `#if` resolved by cpp, interface blocks
(`layout(...) uniform`) become plain structs / image_t globals bound by the entry point below, `out` / `inout` parameters become
references, array constructors become brace lists, expressions of two floating literals are folded in double (what glslang's
constant folding does: 0.18 + 0.01 is ONE fp32 constant), floating literals get an `f` (a GLSL literal is fp32), and functions that
main() never reaches are left out (so that a shader only needs the parts of shared.glsl it uses to compile).  the function
bodies - the arithmetic - are untouched."""
import os
import re
import subprocess
import sys

SCALAR = {"float": 4, "int": 4, "uint": 4, "vec2": 8, "vec3": 12, "vec4": 16, "ivec2": 8, "ivec4": 16, "uvec4": 16, "mat3": 48}
ALIGN = {"float": 4, "int": 4, "uint": 4, "vec2": 8, "vec3": 16, "vec4": 16, "ivec2": 8, "ivec4": 16, "uvec4": 16, "mat3": 16}


def inline_includes(path, search, seen):
    """textual #include resolution (the shaders use GL_GOOGLE_include_directive), dropping #version / #extension lines"""
    out = []
    for ln in open(path).read().splitlines():
        if re.match(r"\s*#\s*(version|extension)\b", ln):
            continue
        m = re.match(r'\s*#\s*include\s+"([^"]+)"', ln)
        if m:
            for d in [os.path.dirname(path)] + search:
                cand = os.path.join(d, m.group(1))
                if os.path.exists(cand):
                    if ("#pragma once" in open(cand).read() or cand.endswith(".h")) and cand in seen:
                        break
                    seen.add(cand)
                    out.append(inline_includes(cand, search, seen))
                    break
            else:
                sys.exit("include not found: " + m.group(1))
            continue
        if re.match(r"\s*#\s*pragma\s+once", ln):
            continue
        out.append(ln)
    return "\n".join(out)


def preprocess(moddir, module, kernel):
    src = inline_includes(os.path.join(moddir, module, kernel + ".comp"), [os.path.join(moddir, module), moddir, os.path.dirname(moddir)], set())
    # GL_core_profile: predefined by every GLSL compiler (matrices.h keys its column major mat3() form on it)
    r = subprocess.run(["cpp", "-P", "-undef", "-nostdinc", "-DGL_core_profile=1", "-x", "c", "-"], input=src, capture_output=True, text=True)
    if r.returncode:
        sys.exit("cpp failed: " + r.stderr[:2000])
    return r.stdout


def split_toplevel(text):
    """top level declarations: text up to a ';' or a balanced {...} (plus a trailing instance name and ';') at depth 0"""
    out, depth, start, i, n = [], 0, 0, 0, len(text)
    while i < n:
        c = text[i]
        if c in "({[":
            depth += 1
        elif c in ")}]":
            depth -= 1
            if c == "}" and depth == 0:
                j = i + 1
                m = re.match(r"\s*\w*\s*(\[\s*\d*\s*\])?\s*;", text[j:])     # struct / interface block instance
                head = text[start:i + 1]
                if m and not re.search(r"\)\s*\{", head.split("{", 1)[0] + "{"):
                    i = j + m.end() - 1
                out.append(text[start:i + 1].strip())
                start = i + 1
        elif c == ";" and depth == 0:
            out.append(text[start:i + 1].strip())
            start = i + 1
        i += 1
    return [d for d in out if d]


def std140(members):
    off, res = 0, []
    for typ, name, count in members:
        size, align = SCALAR[typ], ALIGN[typ]
        if count:
            align = 16
            stride = (size + 15) // 16 * 16
            off = (off + align - 1) // align * align
            res.append((typ, name, count, off, stride))
            off += stride * count
        else:
            off = (off + align - 1) // align * align
            res.append((typ, name, 0, off, size))
            off += size
    return res, off


FLIT = r"(?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+"


def fold_literals(code):
    """glslang folds constant expressions in double and rounds the result to fp32 once (1.0/2.2, 0.18 + 0.01): do the same for
    expressions of two floating literals, where operator precedence allows it, before the literals get their `f`."""
    mul = re.compile(r"(?<![\w.])(%s)\s*([*/])\s*(%s)(?![\w.])" % (FLIT, FLIT))
    add = re.compile(r"(?<![\w.])(%s)\s*([-+])\s*(%s)(?![\w.])" % (FLIT, FLIT))

    def prev_char(code, i):
        while i > 0 and code[i - 1].isspace():
            i -= 1
        return code[i - 1] if i > 0 else "("

    def next_char(code, i):
        while i < len(code) and code[i].isspace():
            i += 1
        return code[i] if i < len(code) else ";"

    changed = True
    while changed:
        changed = False
        for pat, is_mul in ((mul, True), (add, False)):
            pos = 0
            while True:
                m = pat.search(code, pos)
                if not m:
                    break
                before, after = prev_char(code, m.start()), next_char(code, m.end())
                ok = before not in "*/" if is_mul else (before in "(,=?:<>[{" and after not in "*/")
                if is_mul and before in "-+" and prev_char(code, m.start() - 1 - (len(code[:m.start()]) - len(code[:m.start()].rstrip()))) in "eE":
                    ok = False
                if not ok:
                    pos = m.end()
                    continue
                a, b = float(m.group(1)), float(m.group(3))
                v = {"*": a * b, "/": a / b if b != 0.0 else None, "+": a + b, "-": a - b}[m.group(2)]
                if v is None:
                    pos = m.end()
                    continue
                lit = repr(v)
                if "." not in lit and "e" not in lit and "inf" not in lit and "nan" not in lit:
                    lit += ".0"
                if v < 0:
                    lit = "(" + lit + ")"
                code = code[:m.start()] + lit + code[m.end():]
                changed = True
                pos = m.start() + len(lit)
    return code


def subst_const_floats(code):
    """`const float a = 1.0993; ... (a - 1)`: for glslang `a` is a constant kept in double, and `a - 1` one fp32 constant.  replace
    such names by their literal within their block, fold again (constant arguments of log / exp / sqrt included), repeat."""
    import math
    lit1 = r"\(?\s*-?\s*(?:%s)\s*\)?" % FLIT
    fn = re.compile(r"\b(log|exp|sqrt|log2|exp2)\(\s*(%s)\s*\)" % FLIT)
    mixed_a = re.compile(r"(?<![\w.])(%s)\s*([-+*/])\s*(\d+)(?![\w.])" % FLIT)
    mixed_b = re.compile(r"(?<![\w.])(\d+)\s*([-+*/])\s*(%s)(?![\w.])" % FLIT)
    done = set()
    for _ in range(200):
        code = fold_literals(code)
        code2 = fn.sub(lambda m: repr({"log": math.log, "exp": math.exp, "sqrt": math.sqrt, "log2": math.log2, "exp2": lambda x: 2.0 ** x}[m.group(1)](float(m.group(2)))), code)
        # an integer literal next to a floating one converts: make it floating so that the pair folds (only inside parentheses / after = ,)
        code2 = re.sub(r"([(=,]\s*)(%s)(\s*[-+]\s*)(\d+)(\s*[),;])" % FLIT, lambda m: m.group(1) + m.group(2) + m.group(3) + m.group(4) + ".0" + m.group(5), code2)
        if code2 != code:
            code = code2
            continue
        m = None
        for mm in re.finditer(r"\bconst\s+float\s+([^;]*);", code):
            parts = [x.strip() for x in mm.group(1).split(",")]
            for part in parts:
                pm = re.match(r"^(\w+)\s*=\s*(%s)$" % lit1, part)
                if pm and (mm.start(), pm.group(1)) not in done:
                    m = (mm, pm)
                    break
            if m:
                break
        if not m:
            break
        mm, pm = m
        done.add((mm.start(), pm.group(1)))
        name, lit = pm.group(1), pm.group(2).strip()
        if not lit.startswith("("):
            lit = "(" + lit + ")" if lit.startswith("-") else lit
        # scope: from the end of this declarator list to the end of the enclosing block
        depth, j = 0, mm.end()
        while j < len(code):
            if code[j] == "{": depth += 1
            elif code[j] == "}":
                if depth == 0: break
                depth -= 1
            j += 1
        # later declarators of the same statement may use the name too: substitute there as well
        head, decl, body, tail = code[:mm.start()], code[mm.start():mm.end()], code[mm.end():j], code[j:]
        k = decl.index(pm.group(0)) + len(pm.group(0))
        decl = decl[:k] + re.sub(r"(?<![\w.])%s(?![\w.(])" % name, lit, decl[k:])
        body = re.sub(r"(?<![\w.])%s(?![\w.(])" % name, lit, body)
        code = head + decl + body + tail
    return code


def fix_body(code):
    code = re.sub(r"\b(?:in\s+)?(?:inout|out)\s+(\w+)\s+(\w+)", r"\1 &\2", code)           # out / inout parameters
    while True:                                                                                # array constructors: float[](a, b) -> {a, b}
        m = re.search(r"\b(\w+)\s*\[\s*\d*\s*\]\s*\(", code)
        if not m or not (m.group(1) in SCALAR or m.group(1) in ("mat2", "mat3")):
            break
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(code[j], 0)
            j += 1
        code = code[:m.start()] + "{" + code[m.end():j - 1] + "}" + code[j:]
    code = code.replace("^^", "!=")                                                          # logical xor of two bools
    code = fold_literals(code)
    code = subst_const_floats(code)   # glslang keeps a `const float` in double while it folds the expressions that use it
    code = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", code)   # fp32 literals
    code = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?)[lL][fF]\b", r"\1f", code)
    return code


def main():
    moddir, module, kernel = sys.argv[1:4]
    text = preprocess(moddir, module, kernel)
    decls = split_toplevel(text)
    blocks, images, funcs, others, local = [], [], {}, [], {"x": 8, "y": 8}
    for d in decls:
        m = re.match(r"layout\s*\(([^)]*)\)\s*in\s*;", d)
        if m:
            for axis in "xy":
                mm = re.search(r"local_size_%s\s*=\s*\(?\s*(\d+)" % axis, m.group(1))
                if mm:
                    local[axis] = int(mm.group(1))
            continue
        m = re.match(r"shared\s+(.*)", d, re.S)
        if m:                                   # workgroup shared memory: one workgroup runs at a time, a plain global does
            others.append("static " + m.group(1))
            continue
        m = re.match(r"layout\s*\(([^)]*)\)\s*uniform\s+(\w+)\s*\{(.*)\}\s*(\w+)\s*;", d, re.S)
        if m:
            members = []
            for mem in m.group(3).split(";"):
                mem = mem.strip()
                if not mem:
                    continue
                typ, rest = mem.split(None, 1)
                for one in rest.split(","):
                    mm = re.match(r"(\w+)\s*(?:\[\s*(\d+)\s*\])?$", one.strip())
                    if not mm or typ not in SCALAR:
                        sys.exit("unsupported block member: " + mem)
                    members.append((typ, mm.group(1), int(mm.group(2) or 0)))
            blocks.append(("push" if "push_constant" in m.group(1) else "params", m.group(2), m.group(4), members))
            continue
        m = re.match(r"layout\s*\(([^)]*)\)\s*uniform\s+(?:writeonly\s+|readonly\s+|coherent\s+)*(sampler2D|image2D)\s+(\w+)\s*(\[\s*\])?\s*;", d)
        if m:
            b = re.search(r"binding\s*=\s*(\d+)", m.group(1))
            s = re.search(r"set\s*=\s*(\d+)", m.group(1))
            if not s or int(s.group(1)) != 1:
                sys.exit("unsupported descriptor set: " + d)
            images.append((int(b.group(1)), m.group(3), bool(m.group(4))))
            continue
        if d.startswith("layout"):
            sys.exit("unsupported interface: " + d[:120])
        m = re.match(r"(?:const\s+)?[\w]+\s+(\w+)\s*\(([^)]*)\)\s*\{", d, re.S)
        if m and d.endswith("}"):
            funcs[m.group(1)] = (funcs[m.group(1)] + "\n" if m.group(1) in funcs else "") + d      # overloads travel together
        else:
            others.append(d)
    if "main" not in funcs:
        sys.exit("no main()")
    # functions reachable from main()
    keep, todo = set(), ["main"]
    while todo:
        f = todo.pop()
        if f in keep:
            continue
        keep.add(f)
        for name in set(re.findall(r"\b(\w+)\s*\(", funcs[f])):
            if name in funcs and name not in keep:
                todo.append(name)
    # a swizzle handed to a function with a single inout parameter: C++ cannot bind the proxy to a reference, go through a temporary
    for f in list(keep):
        m = re.match(r"[\w\s]*?\b(\w+)\s*\(\s*inout\s+(\w+)\s+\w+\s*\)", funcs[f])
        if m:
            name, typ = m.group(1), m.group(2)
            for g in keep:
                funcs[g] = re.sub(r"\b%s\s*\(\s*(\w+\.[xyzwrgba]{2,4})\s*\)\s*;" % name, r"{ %s _t = \1; %s(_t); \1 = _t; }" % (typ, name), funcs[g])
    used_text = "\n".join(funcs[f] for f in keep)
    ns = "shader_%s_%s" % (re.sub(r"\W", "_", module), re.sub(r"\W", "_", kernel))
    o = ['// generated by oracle/glsl/comp2cpp.py from %s/%s.comp of the reference: do not commit' % (module, kernel), '#include "glsl_shim.h"',
         "namespace glsl { namespace %s {" % ns, "static thread_local uvec3 gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID;"]
    for kind, tname, inst, members in blocks:
        o.append("struct %s { %s };" % (tname, " ".join("%s %s%s;" % (t, n, "[%d]" % c if c else "") for t, n, c in members)))
        o.append("static %s %s;" % (tname, inst))
    images.sort()
    for b, name, arr in images:
        o.append("static image_t %s%s;" % (name, "[64]" if arr else ""))
    for d in others:   # global constants and structs, only those the kept functions mention
        m = re.match(r"(?:static\s+)?(?:const\s+)?(?:struct\s+)?\w+\s+(\w+)", d)
        if m and re.search(r"\b%s\b" % re.escape(m.group(1)), used_text + "\n".join(x for x in others if x is not d)):
            o.append(fix_body(d))
    order = [f for f in funcs if f in keep]    # source order: callees are defined before their callers in GLSL
    for f in order:
        o.append(re.sub(r"void\s+main\s*\(", "static void shader_main(", fix_body(funcs[f]), count=1) if f == "main" else fix_body(funcs[f]))
    o.append("}}")
    o.append('extern "C" int %s(const void *params_blob, int params_size, const void *push_blob, int push_size, const glsl::image_t *imgs, const int *counts, int nbind, int wd, int ht, int dp)' % ns)
    o.append("{ using namespace glsl; using namespace glsl::%s;" % ns)
    for kind, tname, inst, members in blocks:
        lay, total = std140(members)
        src, size = ("push_blob", "push_size") if kind == "push" else ("params_blob", "params_size")
        o.append("  if(%s < %d) return -1;" % (size, max(off + (stride * cnt if cnt else sz) for _, _, cnt, off, stride in lay for sz in [stride])))
        for typ, name, cnt, off, stride in lay:
            if cnt:
                o.append("  for(int k = 0; k < %d; k++) memcpy(&%s.%s[k], (const char *)%s + %d + %d * k, %d);" % (cnt, inst, name, src, off, stride, SCALAR[typ]))
            elif typ == "mat3":
                o.append("  for(int k = 0; k < 3; k++) memcpy(&%s.%s[k], (const char *)%s + %d + 16 * k, 12);" % (inst, name, src, off))
            else:
                o.append("  memcpy(&%s.%s, (const char *)%s + %d, %d);" % (inst, name, src, off, SCALAR[typ]))
    o.append("  if(nbind != %d) return -2;" % len(images))
    o.append("  int at = 0;")
    for i, (b, name, arr) in enumerate(images):
        if arr:
            o.append("  if(counts[%d] > 64) return -3; for(int k = 0; k < counts[%d]; k++) %s[k] = imgs[at + k]; at += counts[%d];" % (i, i, name, i))
        else:
            o.append("  %s = imgs[at]; at += counts[%d];" % (name, i))
    if re.search(r"\bbarrier\s*\(", used_text):
        # the reference dispatches ceil(wd / 8) x ceil(ht / 8) workgroups whatever the shader's local size is (record-cmd.h:377-380):
        # nodes with other local sizes pre-scale wd / ht (demosaic/main.c:125-131)
        o.append("  const int lx = %d, ly = %d;" % (local["x"], local["y"]))
        o.append("  for(int gy = 0; gy < (ht + 7) / 8; gy++) for(int gx = 0; gx < (wd + 7) / 8; gx++)")
        o.append("    run_workgroup(lx, ly, [gx, gy, lx, ly](int x, int y) { gl_WorkGroupID = uvec3(gx, gy, 0); gl_LocalInvocationID = uvec3(x, y, 0); gl_GlobalInvocationID = uvec3(gx * lx + x, gy * ly + y, 0); shader_main(); });")
    else:
        o.append("  for(int z = 0; z < dp; z++) for(int y = 0; y < ht; y++) for(int x = 0; x < wd; x++) { gl_GlobalInvocationID = uvec3(x, y, z); shader_main(); }")
    o.append("  return 0;\n}")
    print("\n".join(o))


if __name__ == "__main__":
    main()
