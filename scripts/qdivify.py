#!/usr/bin/env python
"""Rewrite the FP64 divisions `L / R` of selected regions of a CUDA source into `QDIV(L, R)` calls.

Used once on uvic2.9_b200/csrc/k_mobi.cu (the MOBI kernels hold ~60 divides per Euler sub-step; the compiler expands each
into the IEEE sequence plus an exponent-range test and a call to an out-of-line slow path that these operands never
need -- see qdiv() in k_mobi.cu).  The transformation is purely syntactic and keeps C's evaluation order:
`a * b / c * d` -> `QDIV(a * b, c) * d`, `a / b / c` -> `QDIV(QDIV(a, b), c)`.  Divisions whose two operands are numeric
literals are left for the compiler to fold; preprocessor lines and comments are not touched.

    python scripts/qdivify.py FILE START_MARKER END_MARKER [START END ...]   # rewrites FILE in place
"""
import re
import sys

TOK = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<pp>\#(?:\\\n|[^\n])*)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFuUlL]*)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<op>->|\+\+|--|<<=|>>=|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|::|[-+*/%<>=!&|^~?:;,.(){}\[\]])
""", re.X | re.S | re.M)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = TOK.match(src, pos)
        if not m:
            raise SystemExit(f"cannot tokenize at {pos}: {src[pos:pos + 40]!r}")
        out.append((m.lastgroup, m.group()))
        pos = m.end()
    return out


SKIP = ("ws", "lc", "bc", "pp")
LOW = {"+", "-", "<", ">", "<=", ">=", "==", "!=", "&&", "||", "=", "+=", "-=", "*=", "/=", ",", "?", ":", ";", "{", "}",
       "<<", ">>", "&", "|", "^", "%="}


def prev_sig(toks, i):
    i -= 1
    while i >= 0 and toks[i][0] in SKIP:
        i -= 1
    return i


def next_sig(toks, i):
    i += 1
    while i < len(toks) and toks[i][0] in SKIP:
        i += 1
    return i


def operand_end(tok):
    return tok[0] in ("num", "id") or tok[1] in (")", "]")


def left_start(toks, i):
    """index of the first token of the multiplicative chain that ends right before toks[i] ('/')"""
    depth, j = 0, i
    start = i
    while True:
        j = prev_sig(toks, j)
        if j < 0:
            return start
        k, t = toks[j]
        if t in (")", "]"):
            depth += 1
        elif t in ("(", "["):
            if depth == 0:
                return start
            depth -= 1
        elif depth == 0:
            if k == "id" and t in ("return", "else", "case", "const", "double"):
                return start
            if t in LOW:
                if t in ("+", "-"):
                    p = prev_sig(toks, j)
                    unary = p < 0 or not operand_end(toks[p])
                    if not unary:
                        return start
                    # unary sign: belongs to the operand only when it directly starts the chain (e.g. `x * -b / c`, `(-b / c)`)
                else:
                    return start
        start = j


def right_end(toks, i):
    """index of the last token of the primary expression that starts after toks[i] ('/')"""
    j = next_sig(toks, i)
    while toks[j][1] in ("+", "-", "!", "~"):   # unary prefix
        j = next_sig(toks, j)
    k, t = toks[j]
    if t == "(":
        j = match_close(toks, j)
    elif k not in ("num", "id"):
        raise SystemExit(f"unexpected right operand {t!r}")
    # postfix: calls, subscripts, member access
    while True:
        n = next_sig(toks, j)
        if n >= len(toks):
            return j
        t = toks[n][1]
        if t in ("(", "["):
            j = match_close(toks, n)
        elif t in ("->", "."):
            j = next_sig(toks, n)
        else:
            return j


def match_close(toks, j):
    opener = toks[j][1]
    closer = {"(": ")", "[": "]"}[opener]
    depth = 0
    while True:
        t = toks[j][1]
        if toks[j][0] == "op":
            if t == opener:
                depth += 1
            elif t == closer:
                depth -= 1
                if depth == 0:
                    return j
        j += 1


def is_literal(toks, a, b):
    sig = [t for t in toks[a:b + 1] if t[0] not in SKIP]
    const = ("RN15STD", "RC13STD", "RC14STD", "TRCMIN")   # literal macros of k_mobi.cu
    return all(k == "num" or t in const or t in ("(", ")", "+", "-", "*") for k, t in sig) and \
        any(k == "num" or t in const for k, t in sig)


def integer_context(toks, i):
    """the division sits in an integer declaration / assignment to an int, or inside a subscript"""
    depth = 0
    j = i
    while True:
        j = prev_sig(toks, j)
        if j < 0:
            return False
        t = toks[j][1]
        if t == "]":
            depth += 1
        elif t == "[":
            if depth == 0:
                return True
            depth -= 1
        elif t in (";", "{", "}"):
            n = next_sig(toks, j)
            words = []
            while toks[n][0] == "id" and len(words) < 3:
                words.append(toks[n][1])
                n = next_sig(toks, n)
            return any(w in ("int", "long", "unsigned", "size_t", "bool") for w in words)


def rewrite(region):
    n = 0
    start_at = 0
    while True:
        toks = tokenize(region)
        idx = None
        seen = 0
        for i, (k, t) in enumerate(toks):
            if k == "op" and t == "/":
                if seen >= start_at:
                    idx = i
                    break
                seen += 1
        if idx is None:
            return region, n
        a = left_start(toks, idx)
        b = right_end(toks, idx)
        if integer_context(toks, idx) or (is_literal(toks, a, idx - 1) and is_literal(toks, idx + 1, b)):
            start_at += 1       # leave it to constant folding
            continue
        left = "".join(t for k_, t in toks[a:idx] if k_ not in ("lc", "bc")).strip()
        right = "".join(t for k_, t in toks[idx + 1:b + 1] if k_ not in ("lc", "bc")).strip()
        left = re.sub(r"\s*\n\s*", " ", left)
        right = re.sub(r"\s*\n\s*", " ", right)
        new = "".join(t for _, t in toks[:a]) + f"QDIV({left}, {right})" + "".join(t for _, t in toks[b + 1:])
        region = new
        n += 1


def main():
    path = sys.argv[1]
    marks = sys.argv[2:]
    src = open(path).read()
    total = 0
    for s_mark, e_mark in zip(marks[0::2], marks[1::2]):
        a = src.index(s_mark)
        b = src.index(e_mark, a)
        new, n = rewrite(src[a:b])
        src = src[:a] + new + src[b:]
        total += n
        print(f"{s_mark[:50]!r}: {n} divisions rewritten")
    open(path, "w").write(src)
    print("total", total)


if __name__ == "__main__":
    main()
