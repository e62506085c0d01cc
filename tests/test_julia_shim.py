"""The Julia shim (rxinfer.jl_b200/julia/RxGaussB200.jl) cannot be executed here (no Julia in the image); what CAN be
checked on the CPU: it binds EVERY export of include/rxgauss.h with the right arity, its block structure is balanced
(a cheap stand-in for `julia/check_syntax.jl`), and the fallback list covers SURVEY.md appendix C."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "rxinfer.jl_b200", "julia", "RxGaussB200.jl")
HDR = os.path.join(ROOT, "include", "rxgauss.h")


def _strip(src):
    src = re.sub(r'"""(.|\n)*?"""', '""', src)            # docstrings
    src = re.sub(r'"(\\.|[^"\\\n])*"', '""', src)         # strings
    return "\n".join(l.split("#")[0] for l in src.splitlines())


def test_every_export_is_bound_with_the_right_arity():
    hdr = open(HDR).read()
    jl = open(JL).read()
    decls = re.findall(r"^\s*(?:int|long long|double|const char\*)\s+(rxg_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, re.M | re.S)
    assert len(decls) >= 60
    for name, args in decls:
        m = re.search(r"ccall\(\(:%s, LIB\),\s*\w+,\s*(\(\)|\(.*?\)|\w+)\s*[,)]" % name, jl, re.S)
        assert m, f"{name} is not bound in RxGaussB200.jl"
        nargs = 0 if args.strip() in ("", "void") else len([a for a in re.sub(r"\[[^\]]*\]", "", args).split(",")])
        sig = m.group(1)
        if sig.startswith("("):
            inner = sig[1:-1].strip().rstrip(",")
            depth, parts, cur = 0, [], ""
            for ch in inner:
                if ch in "{(":
                    depth += 1
                if ch in "})":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur); cur = ""
                else:
                    cur += ch
            if cur.strip():
                parts.append(cur)
            assert len(parts) == nargs, f"{name}: header has {nargs} arguments, the ccall signature {len(parts)}"
        else:                                                       # a named tuple constant (R9, P8, ...)
            const = re.search(r"const %s = \((.*?)\)\n" % sig, jl, re.S)
            assert const, (name, sig)
            assert len([p for p in const.group(1).split(",") if p.strip()]) - const.group(1).count("Ptr{Cvoid}, ") * 0 >= 1
            n = len(re.findall(r"Ptr\{[A-Za-z0-9]+\}|F32P|Int64|Cint|Cuint|Cfloat|Csize_t|Clonglong", const.group(1)))
            assert n == nargs, f"{name}: header has {nargs} arguments, {sig} has {n}"


def test_block_structure_is_balanced():
    """Block openers and `end`s balance.  Tokens inside brackets are not block syntax: `for` / `if` inside [...] or (...)
    belong to comprehensions / generators, `end` inside [...] is the last-index keyword."""
    src = _strip(open(JL).read())
    opener = re.compile(r"(?<![\w.:!@])(module|function|struct|if|for|while|let|do|try|begin|macro|quote)(?![\w!])")
    ender = re.compile(r"(?<![\w.:!])end(?![\w!])")
    depth = 0
    for ln, line in enumerate(src.splitlines(), 1):
        sq = par = 0
        i = 0
        while i < len(line):
            ch = line[i]
            if ch == "[":
                sq += 1
            elif ch == "]":
                sq -= 1
            elif ch == "(":
                par += 1
            elif ch == ")":
                par -= 1
            else:
                m = opener.match(line, i)
                if m:
                    word = m.group(1)
                    if not ((word in ("for", "if")) and (sq > 0 or par > 0)):
                        depth += 1
                    i = m.end()
                    continue
                m = ender.match(line, i)
                if m:
                    if sq == 0:
                        depth -= 1
                    i = m.end()
                    continue
            i += 1
        assert depth >= 0, f"unbalanced `end` at line {ln}"
    assert depth == 0, depth
    for a, b in ("()", "[]", "{}"):
        assert src.count(a) == src.count(b), (a, src.count(a), src.count(b))


def test_fallback_list_covers_appendix_c():
    jl = open(JL).read()
    for kw in ("callbacks", "annotations", "predictvars", "meta", "options", "events", "uselock"):
        assert f":{kw}" in jl
    assert "stock()" in jl and "RxInfer.infer(" in jl
