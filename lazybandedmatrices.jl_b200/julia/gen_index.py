#!/usr/bin/env python
"""Prints the dispatch index of LBMB200.jl for INTEGRATION.md: every method / wrapper of the shim with its line number and the
C-ABI entry points its body calls (direct `c_lbm_*` calls and the `$var` interpolations of the @eval loops).
   python julia/gen_index.py            -> markdown table rows on stdout
   python julia/gen_index.py --write    -> rewrites the block between the INDEX markers in INTEGRATION.md"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = open(os.path.join(HERE, "LBMB200.jl"), encoding="utf-8").read().splitlines()


def oneliner(text):
    """`name(args...) [where ...] = body` (short-form method): returns the signature, else None."""
    m = re.match(r"^\s*([A-Za-z_][\w\.!]*(?:\{[^}]*\})?)\(", text)
    if not m or text.lstrip().startswith(("if ", "for ", "while ", "return ", "check(", "mcheck(", "throw(")):
        return None
    depth, j = 0, m.end() - 1
    while j < len(text):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                break
        j += 1
    if depth != 0:
        return None
    rest = text[j + 1:]
    w = re.match(r"^(\s+where\s+[\w{},<: ]+?)?\s*=(?!=)", rest)
    if not w:
        return None
    return text[:j + 1].strip() + (w.group(1) or "")


def rows():
    out = []
    var2name = {}
    i = 0
    defs = []   # (line, signature text, body lines)
    start_re = re.compile(r"^\s*(function\s+(.+)|([A-Za-z_][\w\.!:\*\+\-]*\(.*\)(\s+where\s+.+?)?)\s*=\s*(?!=).*)$")
    for ln, text in enumerate(SRC, 1):
        for var, suffix in re.findall(r"(\w+) = Symbol\(:c_lbm_, p, :([a-z0-9_]+)\)", text):
            var2name[var] = "lbm_?" + suffix
    n = len(SRC)
    ln = 0
    while ln < n:
        text = SRC[ln]
        m = re.match(r"^\s*function\s+(.+)$", text)
        one = None if m else oneliner(text)
        if m or one:
            sig = (m.group(1) if m else one).strip()
            body = [text]
            j = ln + 1
            if m:
                depth = 1
                while j < n and depth > 0:
                    t = SRC[j]
                    if re.match(r"^\s*(function|for|if|while|let|begin|try|struct|mutable struct|do)\b", t) or re.search(r"\bdo\s*$", t):
                        depth += 1
                    if re.match(r"^\s*end\b", t):
                        depth -= 1
                    body.append(t)
                    j += 1
            else:
                while j < n and SRC[j].startswith("    ") and SRC[j - 1].rstrip().endswith(("=", ",", "(", "||", "&&")):
                    body.append(SRC[j])
                    j += 1
            btxt = "\n".join(body)
            calls = set(re.findall(r"\bc_(lbm_[a-z0-9_]+)\(", btxt))
            for var in re.findall(r"\$(\w+)\(", btxt):
                if var in var2name:
                    calls.add(var2name[var])
            if calls:
                defs.append((ln + 1, sig, sorted(calls)))
            ln = j if m else ln + 1
        else:
            ln += 1
    for line, sig, calls in defs:
        sig = re.sub(r"\s+", " ", sig)
        if len(sig) > 150:
            sig = sig[:147] + "..."
        out.append(f"| `LBMB200.jl:{line}` | `{sig.replace('|', chr(92) + '|')}` | " + ", ".join(f"`{c}`" for c in calls) + " |")
    return out


if __name__ == "__main__":
    r = rows()
    if "--write" in sys.argv:
        p = os.path.join(ROOT, "INTEGRATION.md")
        s = open(p, encoding="utf-8").read()
        a, b = s.index("<!-- BEGIN INDEX -->"), s.index("<!-- END INDEX -->")
        s = s[:a] + "<!-- BEGIN INDEX -->\n| shim line | method / wrapper (as written in the shim) | C-ABI entries its body calls |\n|---|---|---|\n" + "\n".join(r) + "\n" + s[b:]
        open(p, "w", encoding="utf-8").write(s)
        print(f"wrote {len(r)} rows")
    else:
        print("\n".join(r))
