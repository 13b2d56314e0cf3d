"""A small parser for the Go expression subset used by Kueue's table tests
(builder call chains, composite literals, slices, maps, basic arithmetic).

Used ONLY by the tools/transcribe_*.py scripts to turn the reference's test tables into
JSON fixtures; the reference is parsed, never executed (no Go toolchain in this image).

AST nodes are tuples:
  ('id', name) ('str', s) ('num', n) ('sel', x, name) ('call', fn, [args])
  ('index', x, i) ('unary', op, x) ('bin', op, a, b)
  ('comp', type_or_None, [(key_or_None, value), ...]) ('type', text) ('func', None)
"""
from __future__ import annotations

import re

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<str>"(?:\\.|[^"\\])*"|`[^`]*`)
  | (?P<num>\d[\d_]*(?:\.\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>:=|\.\.\.|[-+*/&|!<>=]=?|[(){}\[\],:.;])
""", re.X | re.S)


def tokenize(src: str):
    out, pos = [], 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"bad char {src[pos]!r} at {pos}: {src[pos-30:pos+30]!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            if "\n" in m.group():
                out.append(("nl", "\n"))
            continue
        out.append((k, m.group()))
    out.append(("eof", ""))
    return out


class Parser:
    def __init__(self, src: str):
        self.t = tokenize(src)
        self.i = 0

    def peek(self, skip_nl=True):
        j = self.i
        while skip_nl and self.t[j][0] == "nl":
            j += 1
        return self.t[j]

    def next(self, skip_nl=True):
        while skip_nl and self.t[self.i][0] == "nl":
            self.i += 1
        tok = self.t[self.i]
        self.i += 1
        return tok

    def _skip_nl(self, j):
        while self.t[j][0] == "nl":
            j += 1
        return j

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            ctx = " ".join(x[1] for x in self.t[max(0, self.i - 12):self.i + 5])
            raise SyntaxError(f"expected {val!r} got {tok!r} near: {ctx}")
        return tok

    # ---- types (only what appears before composite literals) ----
    def parse_type(self):
        tok = self.peek()
        if tok[1] == "[":
            self.next(); self.expect("]")
            return ("type", "[]" + self.type_text(self.parse_type()))
        if tok[1] == "*":
            self.next()
            return ("type", "*" + self.type_text(self.parse_type()))
        if tok[1] == "map":
            self.next(); self.expect("[")
            k = self.parse_type(); self.expect("]")
            v = self.parse_type()
            return ("type", f"map[{self.type_text(k)}]{self.type_text(v)}")
        if tok[1] == "struct":
            self.next(); self.expect("{")
            depth = 1
            while depth:
                t = self.next()
                depth += t[1] == "{"
                depth -= t[1] == "}"
            return ("type", "struct")
        if tok[0] == "id":
            name = self.next()[1]
            while self.peek(False)[1] == ".":
                self.next(); name += "." + self.next()[1]
            if self.peek(False)[1] == "[" and self.t[self.i + 1][0] == "id" and self.t[self.i + 2][1] == "]":
                pass  # generic instantiation handled by postfix
            return ("type", name)
        raise SyntaxError(f"type expected at {tok}")

    @staticmethod
    def type_text(t):
        return t[1]

    # ---- expressions ----
    def parse_expr(self):
        return self.parse_bin(0)

    PREC = {"||": 1, "&&": 2, "==": 3, "!=": 3, "<": 3, ">": 3, "<=": 3, ">=": 3, "+": 4, "-": 4, "*": 5, "/": 5}

    def parse_bin(self, minp):
        left = self.parse_unary()
        while True:
            tok = self.peek(False)
            op = tok[1]
            if tok[0] != "op" or op not in self.PREC or self.PREC[op] < minp:
                return left
            self.next()
            right = self.parse_bin(self.PREC[op] + 1)
            left = ("bin", op, left, right)

    def parse_unary(self):
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("*", "&", "-", "!"):
            self.next()
            return ("unary", tok[1], self.parse_unary())
        return self.parse_postfix()

    def parse_comp_body(self, typ):
        self.expect("{")
        elems = []
        while self.peek()[1] != "}":
            if self.peek()[1] == "{":
                v = self.parse_comp_body(None)
                k = None
            else:
                v = self.parse_expr()
                k = None
            if self.peek()[1] == ":":
                self.next()
                k = v
                v = self.parse_comp_body(None) if self.peek()[1] == "{" else self.parse_expr()
            elems.append((k, v))
            if self.peek()[1] == ",":
                self.next()
        self.expect("}")
        return ("comp", typ, elems)

    def parse_primary(self):
        tok = self.peek()
        if tok[0] == "str":
            self.next()
            s = tok[1]
            return ("str", s[1:-1] if s[0] == "`" else bytes(s[1:-1], "utf-8").decode("unicode_escape"))
        if tok[0] == "num":
            self.next()
            txt = tok[1].replace("_", "")
            return ("num", float(txt) if "." in txt else int(txt))
        if tok[1] == "(":
            self.next(); e = self.parse_expr(); self.expect(")")
            return e
        if tok[1] in ("[", "map", "struct"):
            typ = self.parse_type()
            if self.peek(False)[1] == "{":
                return self.parse_comp_body(typ)
            return typ
        if tok[1] == "func":
            # function literal: keep `name := expr` / `return expr` statements of the body (enough for the
            # immediately-invoked table helpers); anything else makes the body opaque
            self.next()
            depth = 0
            while True:  # skip the signature up to the body's opening brace
                t = self.peek()
                if t[1] == "{" and depth == 0:
                    break
                depth += t[1] in ("(", "[")
                depth -= t[1] in (")", "]")
                self.next()
            start = self.i
            try:
                self.expect("{")
                stmts = []
                while self.peek()[1] != "}":
                    t = self.peek()
                    if t[1] == "return":
                        self.next()
                        stmts.append(("return", None, self.parse_expr()))
                    elif t[0] == "id" and self.t[self._skip_nl(self.i) + 1][1] in (":=", "="):
                        name = self.next()[1]
                        self.next()
                        stmts.append(("assign", name, self.parse_expr()))
                    else:
                        raise SyntaxError("opaque statement")
                self.expect("}")
                return ("func", stmts)
            except SyntaxError:
                self.i = start
                depth = 0
                while True:
                    t = self.next()
                    if t[1] == "{":
                        depth += 1
                    elif t[1] == "}":
                        depth -= 1
                        if depth == 0:
                            break
                return ("func", None)
        if tok[0] == "id":
            self.next()
            return ("id", tok[1])
        raise SyntaxError(f"unexpected token {tok} near {' '.join(x[1] for x in self.t[max(0,self.i-10):self.i+5])}")

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tok = self.peek(False)
            if tok[0] == "nl":
                # a selector may continue on the next line: "foo.\n  Bar()"  (dot ends the line)
                return e
            if tok[1] == ".":
                self.next(False)
                name = self.next()  # newline allowed after the dot
                if name[1] == "(":  # type assertion — not used
                    raise SyntaxError("type assertion")
                e = ("sel", e, name[1])
            elif tok[1] == "(":
                self.next()
                args = []
                while self.peek()[1] != ")":
                    args.append(self.parse_expr())
                    if self.peek()[1] == "...":
                        self.next()
                    if self.peek()[1] == ",":
                        self.next()
                self.expect(")")
                e = ("call", e, args)
            elif tok[1] == "[":
                self.next()
                idx = self.parse_type() if self.peek()[0] == "id" and self.t[self.i + 1][1] == "]" and self._looks_like_type() else self.parse_expr()
                self.expect("]")
                e = ("index", e, idx)
            elif tok[1] == "{" and self._is_typeish(e):
                e = self.parse_comp_body(("type", self._type_name(e)))
            else:
                return e

    def _looks_like_type(self):
        return self.peek()[1] in ("int32", "int64", "int", "string", "bool", "float64")

    def _is_typeish(self, e):
        # composite literal after a (qualified) type name: Foo{...} / pkg.Foo{...}
        if e[0] == "id":
            return e[1][0].isupper() or e[1] in ("struct", "rfMap")  # rfMap: local map type alias in flavorassigner_test.go
        if e[0] == "sel":
            return e[1][0] == "id" and e[2][0].isupper()
        return False

    def _type_name(self, e):
        return e[1] if e[0] == "id" else f"{e[1][1]}.{e[2]}"


def find_matching(src: str, open_pos: int) -> int:
    """Index of the brace matching src[open_pos] == '{', skipping strings and comments."""
    depth, i, n = 0, open_pos, len(src)
    while i < n:
        c = src[i]
        if c == '"':
            i += 1
            while src[i] != '"':
                i += 2 if src[i] == "\\" else 1
        elif c == "`":
            i = src.index("`", i + 1)
        elif src.startswith("//", i):
            i = src.index("\n", i)
            continue
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def parse_expression(src: str):
    p = Parser(src)
    e = p.parse_expr()
    return e


def parse_composite_body(src: str):
    """src starts with '{': parse it as an (elided-type) composite literal."""
    p = Parser(src)
    return p.parse_comp_body(None)
