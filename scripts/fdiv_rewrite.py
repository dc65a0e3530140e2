#!/usr/bin/env python3
"""One-time source rewrite: every fp64 `a / b` in the device kernels -> or_div(a, b), sqrt( -> or_sqrt(.
Token based; left operand = the whole multiplicative chain to the left (C precedence: (a*b)/c)."""
import re, sys

TOK = re.compile(r'''
  (?P<ws>\s+) | (?P<lc>//[^\n]*) | (?P<bc>/\*.*?\*/) | (?P<pp>\#(?:[^\n\\]|\\\n|\\.)*) |
  (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFuUlL]*|0[xX][0-9a-fA-F]+[uUlL]*) |
  (?P<id>[A-Za-z_][A-Za-z_0-9]*) | (?P<str>"(?:\\.|[^"\\])*") |
  (?P<op>->|\+\+|--|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||[-+*/%]=|::|[-+*/%=<>!&|^~?:;,.(){}\[\]])
''', re.X | re.S | re.M)

def tokenize(s):
    out = []; i = 0
    while i < len(s):
        m = TOK.match(s, i)
        if not m: raise SystemExit(f"tokenize failed at {i}: {s[i:i+40]!r}")
        out.append((m.lastgroup, m.group())); i = m.end()
    return out

SKIP = ('ws', 'lc', 'bc')
def prev_sig(t, i):
    i -= 1
    while i >= 0 and t[i][0] in SKIP: i -= 1
    return i
def next_sig(t, i):
    i += 1
    while i < len(t) and t[i][0] in SKIP: i += 1
    return i
def match_fwd(t, i):      # t[i] is ( or [ -> index of matching closer
    o = t[i][1]; c = {'(': ')', '[': ']'}[o]; d = 0
    while True:
        if t[i][1] == o: d += 1
        elif t[i][1] == c:
            d -= 1
            if d == 0: return i
        i += 1
def match_bwd(t, i):
    c = t[i][1]; o = {')': '(', ']': '['}[c]; d = 0
    while True:
        if t[i][1] == c: d += 1
        elif t[i][1] == o:
            d -= 1
            if d == 0: return i
        i -= 1

def primary_fwd(t, i):    # i = first token of a unary/primary -> last index
    while t[i][1] in ('-', '+', '!'): i = next_sig(t, i)
    if t[i][1] == '(': j = match_fwd(t, i)
    elif t[i][0] in ('id', 'num'): j = i
    else: raise SystemExit(f"bad right operand at {t[i]}")
    while True:
        k = next_sig(t, j)
        if k < len(t) and t[k][1] in ('(', '['): j = match_fwd(t, k)
        elif k < len(t) and t[k][1] == '.' and t[next_sig(t, k)][0] == 'id': j = next_sig(t, k)
        else: return j

def primary_bwd(t, i):    # i = last token of a primary -> first index
    j = i
    while True:
        if t[j][1] in (')', ']'):
            j = match_bwd(t, j)
            k = prev_sig(t, j)
            if t[k][0] == 'id' and t[k][1] not in ('return', 'if', 'while', 'for'): j = k
            elif t[k][1] in (')', ']'): j = k; continue      # a[i][j]
            else: return j
        elif t[j][0] in ('id', 'num'): pass
        else: raise SystemExit(f"bad left operand at {t[j]}")
        k = prev_sig(t, j)
        if t[k][1] == '.': j = prev_sig(t, k); continue
        return j

def rewrite(src):
    n = 0
    while True:
        t = tokenize(src)
        idx = None
        for i, (ty, tx) in enumerate(t):
            if ty == 'op' and tx == '/':
                a = t[prev_sig(t, i)]; b = t[next_sig(t, i)]
                if a[0] == 'num' and b[0] == 'num': continue          # constant folding stays with the compiler
                idx = i; break
        if idx is None: return src, n
        i = idx
        r0 = next_sig(t, i); r1 = primary_fwd(t, r0)
        l1 = prev_sig(t, i); l0 = primary_bwd(t, l1)
        while True:                                                   # extend over the multiplicative chain
            k = prev_sig(t, l0)
            if t[k][1] == '*': l0 = primary_bwd(t, prev_sig(t, k))
            else: break
        k = prev_sig(t, l0)
        if t[k][1] == ')' :                                           # a cast such as (double)x would be split
            raise SystemExit("cast before division chain: " + ''.join(x[1] for x in t[max(0, k - 6):r1 + 1]))
        left = ''.join(x[1] for x in t[l0:l1 + 1]); right = ''.join(x[1] for x in t[r0:r1 + 1])
        new = t[:l0] + [('id', f'or_div({left}, {right})')] + t[r1 + 1:]
        src = ''.join(x[1] for x in new); n += 1

if __name__ == '__main__':
    for path in sys.argv[1:]:
        s = open(path).read()
        # protect numeric-literal divisions like 1.0 / 3.0 (handled in rewrite) ; do the rewrite on the whole file
        out, n = rewrite(s)
        out2 = re.sub(r'(?<![A-Za-z_0-9])sqrt\(', 'or_sqrt(', out)
        open(path, 'w').write(out2)
        print(path, n, 'divisions,', out2.count('or_sqrt(') , 'sqrt')
