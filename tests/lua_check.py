"""A small static checker for the LuaJIT shims in depth-estimation_b200/lua/ (test infrastructure).

The build image has no Lua interpreter, so the shims cannot be run here.  This module does what
can be done without one: it tokenises Lua 5.1 source (comments, short and long strings, numbers,
names, operators), checks that blocks and brackets nest and close (function / if / for / while /
do ... end, repeat ... until, (), [], {}), and extracts call sites with their top-level argument
count so that every C.dm_*(...) call can be compared with the prototype in include/depthmatch.h.
It is not a Lua parser: it accepts some programs a real parser would reject.
"""
import re

KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local", "nil",
            "not", "or", "repeat", "return", "then", "true", "until", "while"}
_OPS = ["...", "..", "==", "~=", "<=", ">=", "::"] + list("+-*/%^#<>=(){}[];:,.")


class LuaSyntaxError(Exception):
    pass


def tokenize(src):
    """-> list of (kind, text, line); kind in name, keyword, number, string, op."""
    toks, i, n, line = [], 0, len(src), 1
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1
            i += 1
        elif c in " \t\r":
            i += 1
        elif src.startswith("--", i):
            m = re.match(r"--\[(=*)\[", src[i:])
            if m:
                close = "]" + m.group(1) + "]"
                j = src.find(close, i + len(m.group(0)))
                if j < 0:
                    raise LuaSyntaxError("unterminated long comment at line %d" % line)
                line += src.count("\n", i, j)
                i = j + len(close)
            else:
                j = src.find("\n", i)
                i = n if j < 0 else j
        elif c in "\"'":
            j = i + 1
            while j < n and src[j] != c:
                if src[j] == "\\":
                    j += 1
                if j < n and src[j] == "\n":
                    raise LuaSyntaxError("unterminated string at line %d" % line)
                j += 1
            if j >= n:
                raise LuaSyntaxError("unterminated string at line %d" % line)
            toks.append(("string", src[i:j + 1], line))
            i = j + 1
        elif re.match(r"\[=*\[", src[i:]):
            m = re.match(r"\[(=*)\[", src[i:])
            close = "]" + m.group(1) + "]"
            j = src.find(close, i + len(m.group(0)))
            if j < 0:
                raise LuaSyntaxError("unterminated long string at line %d" % line)
            toks.append(("string", src[i:j + len(close)], line))
            line += src.count("\n", i, j)
            i = j + len(close)
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = re.match(r"0[xX][0-9a-fA-F]+|\d*\.?\d+(?:[eE][+-]?\d+)?|\d+\.", src[i:])
            toks.append(("number", m.group(0), line))
            i += len(m.group(0))
        elif c.isalpha() or c == "_":
            m = re.match(r"[A-Za-z_][A-Za-z_0-9]*", src[i:])
            w = m.group(0)
            toks.append(("keyword" if w in KEYWORDS else "name", w, line))
            i += len(w)
        else:
            for op in _OPS:
                if src.startswith(op, i):
                    toks.append(("op", op, line))
                    i += len(op)
                    break
            else:
                raise LuaSyntaxError("unexpected character %r at line %d" % (c, line))
    return toks


def check_structure(src):
    """Raises LuaSyntaxError when blocks or brackets do not nest; returns the token list."""
    toks = tokenize(src)
    stack = []   # (opener, line)
    pairs = {")": "(", "]": "[", "}": "{"}
    for k, (kind, text, line) in enumerate(toks):
        if kind == "op" and text in "([{":
            stack.append((text, line))
        elif kind == "op" and text in ")]}":
            if not stack or stack[-1][0] != pairs[text]:
                raise LuaSyntaxError("unbalanced %r at line %d" % (text, line))
            stack.pop()
        elif kind == "keyword":
            if text in ("function", "if", "repeat"):
                stack.append((text, line))
            elif text in ("for", "while"):
                stack.append((text, line))          # its `do` is consumed below
            elif text == "do":
                if stack and stack[-1][0] in ("for", "while"):
                    stack[-1] = ("do", stack[-1][1])
                else:
                    stack.append(("do", line))
            elif text == "end":
                if not stack or stack[-1][0] not in ("function", "if", "do"):
                    raise LuaSyntaxError("'end' without an open block at line %d (open: %r)" % (line, stack[-1:] or None))
                stack.pop()
            elif text == "until":
                if not stack or stack[-1][0] != "repeat":
                    raise LuaSyntaxError("'until' without 'repeat' at line %d" % line)
                stack.pop()
            elif text in ("then", "else", "elseif"):
                if not stack or stack[-1][0] != "if":
                    raise LuaSyntaxError("%r outside an if at line %d" % (text, line))
    if stack:
        raise LuaSyntaxError("unclosed %r opened at line %d" % stack[-1])
    return toks


def calls(toks, prefix):
    """Call sites NAME(...) whose dotted name starts with `prefix` (e.g. 'C.dm_'): -> [(name, nargs, line)]."""
    out, k = [], 0
    while k < len(toks):
        kind, text, line = toks[k]
        if kind == "name":
            name, j = text, k + 1
            while j + 1 < len(toks) and toks[j][1] == "." and toks[j + 1][0] == "name":
                name += "." + toks[j + 1][1]
                j += 2
            if name.startswith(prefix) and j < len(toks) and toks[j][1] == "(":
                depth, nargs, empty, i = 0, 1, True, j
                while True:
                    t = toks[i][1] if toks[i][0] == "op" else None
                    if t in ("(", "[", "{"):
                        depth += 1
                    elif t in (")", "]", "}"):
                        depth -= 1
                        if depth == 0:
                            break
                    elif depth == 1 and t == ",":
                        nargs += 1
                    elif depth >= 1 and i > j:
                        empty = False
                    if i > j and depth >= 1 and t not in (",",):
                        empty = False
                    i += 1
                out.append((name, 0 if empty else nargs, line))
            k = j
        else:
            k += 1
    return out


def header_arg_counts(header_text):
    """{function name: number of parameters} from the C prototypes of include/depthmatch.h."""
    text = re.sub(r"/\*.*?\*/", " ", header_text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(dm_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def table_fields(toks, var):
    """Fields assigned in `local <var> = { name = expr, ... }`: {field: first token of expr}."""
    for k in range(len(toks) - 3):
        if toks[k][1] == var and toks[k + 1][1] == "=" and toks[k + 2][1] == "{":
            fields, depth, i = {}, 0, k + 2
            while i < len(toks):
                t = toks[i][1] if toks[i][0] == "op" else None
                if t in ("{", "(", "["):
                    depth += 1
                elif t in ("}", ")", "]"):
                    depth -= 1
                    if depth == 0:
                        return fields
                elif depth == 1 and toks[i][0] == "name" and toks[i + 1][1] == "=" and toks[i + 2][1] != "=":
                    fields[toks[i][1]] = toks[i + 2][1]
                i += 1
    return None
