"""ORACLE (test infrastructure, NOT product code) -- a CPU restatement of the Rego evaluation that
the reference delegates to `rego.Driver.Query`.

The evaluator itself is NOT in /root/reference: it lives in the un-vendored Go modules
  github.com/open-policy-agent/frameworks/constraint v0.0.0-20260223174506-488c888fd079 (go.mod:18)
  github.com/open-policy-agent/opa v1.13.2                                               (go.mod:19)
so this file restates OPA's *published* Rego semantics (topdown evaluation of partial-set rules,
undefined propagation, negation-as-failure, comprehensions, set algebra, `sprintf("%v")` rendering) for
the language subset used by every in-tree ConstraintTemplate.  Parity is pinned
on the reference's own golden results (tests/golden/*.json: test/gator/test/test.bats:73-249, the TestTest table of
pkg/gator/test/test_test.go:85-268, test/gator/verify/suite.yaml, the psp-all-violations suite,
pkg/target/target_integration_test.go:164-520, website/docs/audit.md:52).
PARITY UNPINNED for what the reference holds no expectation for: the exact violation sets / messages of K8sAllowedRepos,
K8sContainerLimits, K8sBannedImageTags, agilebank K8sRequiredLabels and the K8sPSP* templates beyond the all-violating suite,
and `sprintf("%v")` of composite values -- there this file is an "oracle-by-restatement" (SURVEY.md section 8(c)).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Value model: null=None, bool, int (arbitrary precision) / float, str, array=tuple, object=RObj (hashable
dict), set=frozenset.  `UNDEF` is never a value; undefined is represented by "no solution".
"""
from __future__ import annotations

import re
from fractions import Fraction

# --------------------------------------------------------------------------------------------------
# values


class RObj(dict):
    """Immutable, hashable Rego object."""

    __slots__ = ("_h",)

    def __hash__(self):  # type: ignore[override]
        try:
            return self._h
        except AttributeError:
            self._h = hash(frozenset(self.items()))
            return self._h


def from_json(v):
    """JSON (python json.loads output) -> Rego value."""
    if isinstance(v, dict):
        return RObj((k, from_json(x)) for k, x in v.items())
    if isinstance(v, (list, tuple)):
        return tuple(from_json(x) for x in v)
    if isinstance(v, float) and v == int(v) and abs(v) < 2**53:
        return int(v)
    return v


def to_json(v):
    """Rego value -> plain JSON-able python (sets become sorted lists, like OPA's JSON encoding)."""
    if isinstance(v, RObj):
        return {(k if isinstance(k, str) else fmt_value(k)): to_json(x) for k, x in sort_items(v)}
    if isinstance(v, tuple):
        return [to_json(x) for x in v]
    if isinstance(v, frozenset):
        return [to_json(x) for x in sorted_values(v)]
    return v


def type_rank(v):
    # OPA total order across types: null < boolean < number < string < array < object < set
    if v is None:
        return 0
    if isinstance(v, bool):
        return 1
    if isinstance(v, (int, float, Fraction)):
        return 2
    if isinstance(v, str):
        return 3
    if isinstance(v, tuple):
        return 4
    if isinstance(v, RObj):
        return 5
    if isinstance(v, frozenset):
        return 6
    raise TypeError(f"not a rego value: {v!r}")


def compare(a, b):
    """Three-way compare under OPA's total order."""
    ra, rb = type_rank(a), type_rank(b)
    if ra != rb:
        return -1 if ra < rb else 1
    if ra == 0:
        return 0
    if ra in (1, 2, 3):
        return -1 if a < b else (1 if a > b else 0)
    if ra == 4:
        for x, y in zip(a, b):
            c = compare(x, y)
            if c:
                return c
        return (len(a) > len(b)) - (len(a) < len(b))
    if ra == 5:
        ia, ib = sort_items(a), sort_items(b)
        for (ka, va), (kb, vb) in zip(ia, ib):
            c = compare(ka, kb)
            if c:
                return c
            c = compare(va, vb)
            if c:
                return c
        return (len(ia) > len(ib)) - (len(ia) < len(ib))
    sa, sb = sorted_values(a), sorted_values(b)
    for x, y in zip(sa, sb):
        c = compare(x, y)
        if c:
            return c
    return (len(sa) > len(sb)) - (len(sa) < len(sb))


class _Key:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __lt__(self, o):
        return compare(self.v, o.v) < 0


def sorted_values(vs):
    return sorted(vs, key=_Key)


def sort_items(o):
    return sorted(o.items(), key=lambda kv: _Key(kv[0]))


def equal(a, b):
    return compare(a, b) == 0


def fmt_number(n):
    if isinstance(n, bool):
        return "true" if n else "false"
    if isinstance(n, int):
        return str(n)
    if isinstance(n, Fraction):
        if n.denominator == 1:
            return str(n.numerator)
        n = float(n)
    if n == int(n) and abs(n) < 1e21:
        return str(int(n))
    return repr(n)


def fmt_string(s):
    """JSON-style quoting used by OPA when a string is nested in a composite."""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20:
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def fmt_value(v, top=False):
    """OPA `%v` rendering (ast.Value.String()): top-level strings raw, nested strings quoted, sets and
    object keys sorted, ', ' separators; empty set is `set()`.  Pinned in-tree for sets of strings only:
    test/gator/test/test.bats:233 (`{"geo"}`), website/docs/audit.md:52 (`{"gatekeeper"}`)."""
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float, Fraction)):
        return fmt_number(v)
    if isinstance(v, str):
        return v if top else fmt_string(v)
    if isinstance(v, tuple):
        return "[" + ", ".join(fmt_value(x) for x in v) + "]"
    if isinstance(v, RObj):
        return "{" + ", ".join(fmt_value(k) + ": " + fmt_value(x) for k, x in sort_items(v)) + "}"
    if isinstance(v, frozenset):
        if not v:
            return "set()"
        return "{" + ", ".join(fmt_value(x) for x in sorted_values(v)) + "}"
    raise TypeError(v)


# --------------------------------------------------------------------------------------------------
# lexer


class RegoError(Exception):
    pass


TOKEN_RE = re.compile(
    r"""
    (?P<ws>[ \t\r]+)
  | (?P<nl>\n)
  | (?P<comment>\#[^\n]*)
  | (?P<num>-?(?:\d+\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+|\d+))
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<raw>`[^`]*`)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>:=|==|!=|<=|>=|[=<>+\-*/%&|\[\]{}().,;:])
""",
    re.X,
)


class Tok:
    __slots__ = ("kind", "val", "line", "pos", "end")

    def __init__(self, kind, val, line, pos, end):
        self.kind, self.val, self.line, self.pos, self.end = kind, val, line, pos, end

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


def _unescape(s):
    body = s[1:-1]
    out = []
    i = 0
    while i < len(body):
        c = body[i]
        if c == "\\":
            i += 1
            e = body[i]
            if e == "n":
                out.append("\n")
            elif e == "t":
                out.append("\t")
            elif e == "r":
                out.append("\r")
            elif e == "b":
                out.append("\b")
            elif e == "f":
                out.append("\f")
            elif e == "u":
                out.append(chr(int(body[i + 1 : i + 5], 16)))
                i += 4
            else:
                out.append(e)
        else:
            out.append(c)
        i += 1
    return "".join(out)


def lex(src):
    toks = []
    line = 1
    i = 0
    n = len(src)
    while i < n:
        m = TOKEN_RE.match(src, i)
        if not m:
            raise RegoError(f"rego_parse_error: illegal token at line {line}: {src[i:i+10]!r}")
        kind = m.lastgroup
        txt = m.group()
        if kind == "nl":
            line += 1
        elif kind in ("ws", "comment"):
            pass
        elif kind == "num":
            # a leading '-' is only part of the number when it cannot be a binary minus
            if txt[0] == "-" and toks and toks[-1].line == line and (
                toks[-1].kind in ("num", "str", "id") or toks[-1].val in (")", "]", "}")
            ):
                toks.append(Tok("op", "-", line, i, i + 1))
                i += 1
                continue
            v = int(txt) if re.fullmatch(r"-?\d+", txt) else float(txt)
            toks.append(Tok("num", v, line, i, m.end()))
        elif kind == "str":
            toks.append(Tok("str", _unescape(txt), line, i, m.end()))
        elif kind == "raw":
            toks.append(Tok("str", txt[1:-1], line, i, m.end()))
            line += txt.count("\n")
        else:
            toks.append(Tok(kind, txt, line, i, m.end()))
        i = m.end()
    toks.append(Tok("eof", None, line, n, n))
    return toks


# --------------------------------------------------------------------------------------------------
# AST (plain tuples): ("scalar", v) ("var", name) ("ref", head, [args]) ("call", name, [args])
#   ("array", [t]) ("object", [(k,v)]) ("set", [t]) ("acompr", t, body) ("scompr", t, body)
#   ("ocompr", k, v, body)
# statements: ("expr", term) ("not", term) ("some", [names]) ("somein", k, v, coll) ("assign", lhs, rhs)
#   ("unify", lhs, rhs)   -- infix ops are ("call", opname, [l, r])

KEYWORDS = {"not", "some", "default", "package", "import", "true", "false", "null", "else", "with", "as"}
INFIX = {
    "==": "equal", "!=": "neq", "<": "lt", "<=": "lte", ">": "gt", ">=": "gte",
    "+": "plus", "-": "minus", "*": "mul", "/": "div", "%": "rem", "&": "and", "|": "or",
}
PREC = {"==": 1, "!=": 1, "<": 1, "<=": 1, ">": 1, ">=": 1, "in": 1, "|": 2, "&": 3, "+": 4, "-": 4, "*": 5, "/": 5, "%": 5}


class Rule:
    __slots__ = ("name", "kind", "args", "key", "value", "body", "default", "els", "line")

    def __init__(self, name, kind, args, key, value, body, default=False, line=0):
        self.name, self.kind, self.args, self.key, self.value, self.body = name, kind, args, key, value, body
        self.default = default
        self.els = []
        self.line = line


class Parser:
    def __init__(self, src):
        self.toks = lex(src)
        self.i = 0
        self.wild = 0

    # token helpers
    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def at(self, val):
        t = self.peek()
        return t.kind in ("op", "id") and t.val == val

    def accept(self, val):
        if self.at(val):
            return self.next()
        return None

    def expect(self, val):
        t = self.next()
        if t.val != val or t.kind not in ("op", "id"):
            raise RegoError(f"rego_parse_error: expected {val!r} got {t.val!r} at line {t.line}")
        return t

    def same_line(self):
        return self.i > 0 and self.peek().line == self.toks[self.i - 1].line

    # module
    def parse_module(self):
        pkg = None
        rules = []
        self.imports = {}     # alias -> "lib.<...>" (template libs; frameworks: templates.Target.Libs)
        while self.peek().kind != "eof":
            if self.at("package"):
                self.next()
                parts = [self.next().val]
                while self.accept("."):
                    parts.append(self.next().val)
                pkg = ".".join(parts)
            elif self.at("import"):
                self.next()
                parts = [self.next().val]
                while self.same_line() and self.accept("."):
                    parts.append(self.next().val)
                alias = parts[-1]
                if self.same_line() and self.accept("as"):
                    alias = self.next().val
                if parts[0] in ("future", "rego"):
                    pass
                elif parts[:2] == ["data", "lib"] and len(parts) > 2:
                    self.imports[alias] = ".".join(parts[1:])
                else:
                    raise RegoError(f"rego_unsupported: import {'.'.join(parts)}")
            else:
                rules.append(self.parse_rule())
        if pkg is None:
            raise RegoError("rego_parse_error: package expected")
        return pkg, rules

    def parse_rule(self):
        default = bool(self.accept("default"))
        t = self.next()
        if t.kind != "id" or t.val in KEYWORDS:
            raise RegoError(f"rego_parse_error: unexpected {t.val!r} at line {t.line}")
        name, line = t.val, t.line
        args = key = value = None
        kind = "complete"
        if self.same_line() and self.at("("):
            self.next()
            args = []
            while not self.at(")"):
                args.append(self.parse_term())
                if not self.accept(","):
                    break
            self.expect(")")
            kind = "func"
        elif self.same_line() and self.at("["):
            self.next()
            key = self.parse_term()
            self.expect("]")
            kind = "set"
        if self.at("contains"):
            self.next()
            key = self.parse_term()
            kind = "set"
        if self.at("=") or self.at(":="):
            self.next()
            value = self.parse_term()
            if kind == "set":
                kind = "object"
        rule = Rule(name, kind, args, key, value, None, default, line)
        has_if = bool(self.accept("if"))
        if self.at("{"):
            self.next()
            rule.body = self.parse_body("}")
            self.expect("}")
        elif has_if:
            rule.body = [self.parse_stmt()]
        else:
            rule.body = []
        while self.at("else"):
            self.next()
            ev = None
            if self.at("=") or self.at(":="):
                self.next()
                ev = self.parse_term()
            self.accept("if")
            eb = []
            if self.at("{"):
                self.next()
                eb = self.parse_body("}")
                self.expect("}")
            rule.els.append((ev, eb))
        if rule.value is None and kind in ("complete", "func"):
            rule.value = ("scalar", True)
        return rule

    def parse_body(self, closer):
        stmts = []
        while not self.at(closer):
            if self.peek().kind == "eof":
                raise RegoError("rego_parse_error: unexpected eof in body")
            stmts.append(self.parse_stmt())
            if self.accept(";"):
                continue
            if not self.at(closer) and self.same_line():
                t = self.peek()
                raise RegoError(f"rego_parse_error: unexpected {t.val!r} at line {t.line}")
        return stmts

    def parse_stmt(self):
        if self.at("some"):
            self.next()
            first = self.parse_term(no_in=True)
            names = [first]
            while self.accept(","):
                names.append(self.parse_term(no_in=True))
            if self.accept("in"):
                coll = self.parse_term()
                if len(names) == 1:
                    return ("somein", None, names[0], coll)
                return ("somein", names[0], names[1], coll)
            return ("some", [n[1] for n in names])
        if self.at("every") and self.toks[self.i + 1].kind == "id":
            # every [k,] v in coll { body }  (OPA v1 keyword): the body holds for every element; an empty domain is true
            self.next()
            first = self.parse_term(no_in=True)
            second = self.parse_term(no_in=True) if self.accept(",") else None
            self.expect("in")
            coll = self.parse_term()
            self.expect("{")
            body = self.parse_body("}")
            self.expect("}")
            return ("every", first if second is not None else None, second if second is not None else first, coll, body)
        if self.at("not"):
            self.next()
            return ("not", self.parse_expr())
        e = self.parse_expr()
        if self.same_line() and self.at("with"):
            raise RegoError("rego_unsupported: `with` modifier")
        return e

    def parse_expr(self):
        lhs = self.parse_term()
        if self.same_line() and (self.at(":=") or self.at("=")):
            op = self.next().val
            rhs = self.parse_term()
            return ("assign" if op == ":=" else "unify", lhs, rhs)
        return ("expr", lhs)

    def parse_term(self, prec=0, no_in=False, no_bar=False):
        lhs = self.parse_unary()
        while True:
            t = self.peek()
            if not self.same_line():
                break
            if t.kind == "op" and t.val in PREC and not (no_bar and t.val == "|"):
                op = t.val
            elif t.kind == "id" and t.val == "in" and not no_in:
                op = "in"
            else:
                break
            p = PREC[op]
            if p <= prec:
                break
            self.next()
            rhs = self.parse_term(p, no_in)
            if op == "in":
                lhs = ("call", "internal.member_2", [lhs, rhs])
            else:
                lhs = ("call", INFIX[op], [lhs, rhs])
        return lhs

    def parse_unary(self):
        t = self.peek()
        if t.kind == "op" and t.val == "-":
            self.next()
            x = self.parse_unary()
            if x[0] == "scalar" and isinstance(x[1], (int, float)):
                return ("scalar", -x[1])
            return ("call", "minus", [("scalar", 0), x])
        return self.parse_postfix(self.parse_primary())

    def parse_postfix(self, base):
        while self.same_line():
            if self.at("."):
                self.next()
                f = self.next()
                if f.kind != "id":
                    raise RegoError(f"rego_parse_error: bad ref at line {f.line}")
                base = self._ref(base, ("scalar", f.val))
            elif self.at("[") and self.peek().pos == self.toks[self.i - 1].end:
                self.next()
                idx = self.parse_term()
                self.expect("]")
                base = self._ref(base, idx)
            elif self.at("(") and self.peek().pos == self.toks[self.i - 1].end and base[0] in ("var", "ref"):
                self.next()
                args = []
                while not self.at(")"):
                    args.append(self.parse_term())
                    if not self.accept(","):
                        break
                self.expect(")")
                base = ("call", self._dotted(base), args)
            else:
                break
        return base

    @staticmethod
    def _ref(base, idx):
        if base[0] == "ref":
            return ("ref", base[1], base[2] + [idx])
        return ("ref", base, [idx])

    @staticmethod
    def _dotted(t):
        if t[0] == "var":
            return t[1]
        parts = [t[1][1]]
        for a in t[2]:
            parts.append(a[1])
        return ".".join(parts)

    def parse_primary(self):
        t = self.next()
        if t.kind == "num" or t.kind == "str":
            return ("scalar", t.val)
        if t.kind == "id":
            if t.val == "true":
                return ("scalar", True)
            if t.val == "false":
                return ("scalar", False)
            if t.val == "null":
                return ("scalar", None)
            if t.val == "_":
                self.wild += 1
                return ("var", f"$w{self.wild}")
            if t.val == "set" and self.at("(") and self.peek(1).val == ")":
                self.next()
                self.next()
                return ("set", [])
            if t.val in ("not", "some", "default", "package", "import", "else", "with", "as"):
                raise RegoError(f"rego_parse_error: unexpected keyword {t.val!r} at line {t.line}")
            return ("var", t.val)
        if t.kind == "op":
            if t.val == "(":
                e = self.parse_term()
                self.expect(")")
                return e
            if t.val == "[":
                if self.at("]"):
                    self.next()
                    return ("array", [])
                first = self.parse_term(no_bar=True)
                if self.accept("|"):
                    body = self.parse_body("]")
                    self.expect("]")
                    return ("acompr", first, body)
                items = [first]
                while self.accept(","):
                    if self.at("]"):
                        break
                    items.append(self.parse_term())
                self.expect("]")
                return ("array", items)
            if t.val == "{":
                if self.at("}"):
                    self.next()
                    return ("object", [])
                first = self.parse_term(no_bar=True)
                if self.accept(":"):
                    val = self.parse_term(no_bar=True)
                    if self.accept("|"):
                        body = self.parse_body("}")
                        self.expect("}")
                        return ("ocompr", first, val, body)
                    items = [(first, val)]
                    while self.accept(","):
                        if self.at("}"):
                            break
                        k = self.parse_term()
                        self.expect(":")
                        items.append((k, self.parse_term()))
                    self.expect("}")
                    return ("object", items)
                if self.accept("|"):
                    body = self.parse_body("}")
                    self.expect("}")
                    return ("scompr", first, body)
                items = [first]
                while self.accept(","):
                    if self.at("}"):
                        break
                    items.append(self.parse_term())
                self.expect("}")
                return ("set", items)
        raise RegoError(f"rego_parse_error: unexpected token {t.val!r} at line {t.line}")


# --------------------------------------------------------------------------------------------------
# builtins  (OPA v1.13.2 builtin semantics; a type error => undefined, like OPA's non-strict mode)


class Undefined(Exception):
    pass


def _num(x):
    if isinstance(x, bool) or not isinstance(x, (int, float, Fraction)):
        raise Undefined
    return x


def _str(x):
    if not isinstance(x, str):
        raise Undefined
    return x


def _norm(x):
    if isinstance(x, Fraction):
        return x.numerator if x.denominator == 1 else x
    if isinstance(x, float) and x == int(x) and abs(x) < 2**63:
        return int(x)
    return x


def _arith(op, a, b):
    if op in ("minus", "and", "or") and isinstance(a, frozenset) and isinstance(b, frozenset):
        return {"minus": a - b, "and": a & b, "or": a | b}[op]
    a, b = _num(a), _num(b)
    if isinstance(a, float):
        a = Fraction(a)
    if isinstance(b, float):
        b = Fraction(b)
    if op == "plus":
        return _norm(a + b)
    if op == "minus":
        return _norm(a - b)
    if op == "mul":
        return _norm(a * b)
    if op == "div":
        if b == 0:
            raise Undefined
        return _norm(Fraction(a) / Fraction(b))
    if op == "rem":
        if not (isinstance(a, int) and isinstance(b, int)) or b == 0:
            raise Undefined
        r = abs(a) % abs(b)  # Go's big.Int.Rem: sign follows dividend
        return -r if a < 0 else r
    raise Undefined


def _count(x):
    if isinstance(x, str):
        return len(x)  # unicode code points, like OPA
    if isinstance(x, (tuple, RObj, frozenset)):
        return len(x)
    raise Undefined


def _to_number(x):
    if x is None:
        return 0
    if isinstance(x, bool):
        return 1 if x else 0
    if isinstance(x, (int, float, Fraction)):
        return x
    if isinstance(x, str):
        try:
            if re.fullmatch(r"[+-]?\d+", x):
                return int(x)
            if re.fullmatch(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", x):
                return _norm(float(x))
        except ValueError:
            pass
        raise Undefined
    raise Undefined


def _substring(s, off, ln):
    s = _str(s)
    off, ln = _num(off), _num(ln)
    if off < 0:
        raise Undefined
    if off >= len(s):
        return ""
    if ln < 0:
        return s[off:]
    return s[off : off + ln]


def _sprintf(fmt, args):
    """OPA sprintf: Go fmt verbs over ast.Values.  %v/%s -> fmt_value(top=True); %d on integers;
    %% literal.  Other verbs are outside the lowered subset."""
    fmt = _str(fmt)
    if not isinstance(args, tuple):
        raise Undefined
    out = []
    i = 0
    ai = 0
    while i < len(fmt):
        c = fmt[i]
        if c != "%":
            out.append(c)
            i += 1
            continue
        i += 1
        if i >= len(fmt):
            out.append("%!(NOVERB)")
            break
        v = fmt[i]
        i += 1
        if v == "%":
            out.append("%")
            continue
        if ai >= len(args):
            out.append("%!" + v + "(MISSING)")
            continue
        a = args[ai]
        ai += 1
        if v in ("v", "s"):
            out.append(fmt_value(a, top=True))
        elif v == "d":
            if isinstance(a, int) and not isinstance(a, bool):
                out.append(str(a))
            else:
                out.append("%!d(" + _go_type(a) + "=" + fmt_value(a, top=True) + ")")
        elif v == "q" and isinstance(a, str):
            out.append(fmt_string(a))
        else:
            raise RegoError(f"rego_unsupported: sprintf verb %{v}")
    if ai < len(args):
        out.append("%!(EXTRA " + ", ".join(_go_type(a) + "=" + fmt_value(a, top=True) for a in args[ai:]) + ")")
    return "".join(out)


def _go_type(a):
    return "string" if isinstance(a, str) else "ast.Value"


def _trim(s, cut):
    return _str(s).strip(_str(cut)) if cut else _str(s)


def _object_get(o, k, d):
    if not isinstance(o, RObj):
        raise Undefined
    if isinstance(k, tuple):  # path form
        cur = o
        for p in k:
            if isinstance(cur, RObj) and p in cur:
                cur = cur[p]
            elif isinstance(cur, tuple) and isinstance(p, int) and not isinstance(p, bool) and 0 <= p < len(cur):
                cur = cur[p]
            else:
                return d
        return cur
    return _lookup(o, k, d)


def _lookup(o, k, d):
    try:
        v = o[k]
    except (KeyError, TypeError):
        return d
    if isinstance(k, (bool, int, float)):  # python: True == 1 == 1.0; rego: distinct unless both numbers
        for kk, vv in o.items():
            if kk == k and type_rank(kk) == type_rank(k):
                return vv
        return d
    return v


def _any_prefix(search, base):
    ss = (search,) if isinstance(search, str) else search
    bb = (base,) if isinstance(base, str) else base
    if not isinstance(ss, (tuple, frozenset)) or not isinstance(bb, (tuple, frozenset)):
        raise Undefined
    for s in ss:
        _str(s)
    for b in bb:
        _str(b)
    return any(s.startswith(b) for s in ss for b in bb)


def _any_suffix(search, base):
    ss = (search,) if isinstance(search, str) else search
    bb = (base,) if isinstance(base, str) else base
    if not isinstance(ss, (tuple, frozenset)) or not isinstance(bb, (tuple, frozenset)):
        raise Undefined
    for s in ss:
        _str(s)
    for b in bb:
        _str(b)
    return any(s.endswith(b) for s in ss for b in bb)


def _re_match(pat, s):
    # RE2 syntax ~ python `re` for the patterns the fixtures use (anchors, classes, +, *, ?, |, groups)
    try:
        return re.search(_str(pat), _str(s)) is not None
    except re.error:
        raise Undefined


def _bool_coll(x):
    if not isinstance(x, (tuple, frozenset)):
        raise Undefined
    return x


def _concat(sep, xs):
    if not isinstance(xs, (tuple, frozenset)):
        raise Undefined
    items = list(xs) if isinstance(xs, tuple) else sorted_values(xs)
    return _str(sep).join(_str(x) for x in items)


def _member(x, coll):
    if isinstance(coll, RObj):
        return any(equal(x, v) for v in coll.values())
    if isinstance(coll, (tuple, frozenset)):
        return any(equal(x, v) for v in coll)
    return False


BUILTINS = {
    "equal": lambda a, b: equal(a, b),
    "neq": lambda a, b: not equal(a, b),
    "lt": lambda a, b: compare(a, b) < 0,
    "lte": lambda a, b: compare(a, b) <= 0,
    "gt": lambda a, b: compare(a, b) > 0,
    "gte": lambda a, b: compare(a, b) >= 0,
    "plus": lambda a, b: _arith("plus", a, b),
    "minus": lambda a, b: _arith("minus", a, b),
    "mul": lambda a, b: _arith("mul", a, b),
    "div": lambda a, b: _arith("div", a, b),
    "rem": lambda a, b: _arith("rem", a, b),
    "and": lambda a, b: _arith("and", a, b),
    "or": lambda a, b: _arith("or", a, b),
    "count": _count,
    "sprintf": _sprintf,
    "startswith": lambda s, p: _str(s).startswith(_str(p)),
    "endswith": lambda s, p: _str(s).endswith(_str(p)),
    "contains": lambda s, p: _str(p) in _str(s),
    "strings.any_prefix_match": _any_prefix,
    "strings.any_suffix_match": _any_suffix,
    "split": lambda s, d: tuple(_str(s).split(_str(d))) if d != "" else tuple(_str(s)),
    "trim": _trim,
    "trim_space": lambda s: _str(s).strip(" \t\n\r\v\f"),
    "trim_prefix": lambda s, p: _str(s)[len(p):] if _str(s).startswith(_str(p)) else s,
    "trim_suffix": lambda s, p: _str(s)[: len(s) - len(p)] if p and _str(s).endswith(_str(p)) else s,
    "trim_left": lambda s, c: _str(s).lstrip(_str(c)),
    "trim_right": lambda s, c: _str(s).rstrip(_str(c)),
    "replace": lambda s, o, n: _str(s).replace(_str(o), _str(n)),
    "substring": _substring,
    "lower": lambda s: _str(s).lower(),
    "upper": lambda s: _str(s).upper(),
    "concat": _concat,
    "indexof": lambda s, t: _str(s).find(_str(t)),
    "re_match": _re_match,
    "regex.match": _re_match,
    "to_number": _to_number,
    "is_number": lambda x: not isinstance(x, bool) and isinstance(x, (int, float, Fraction)),
    "is_string": lambda x: isinstance(x, str),
    "is_boolean": lambda x: isinstance(x, bool),
    "is_array": lambda x: isinstance(x, tuple),
    "is_object": lambda x: isinstance(x, RObj),
    "is_set": lambda x: isinstance(x, frozenset),
    "is_null": lambda x: x is None,
    "any": lambda xs: any(x is True for x in _bool_coll(xs)),
    "all": lambda xs: all(x is True for x in _bool_coll(xs)),
    "object.get": _object_get,
    "array.concat": lambda a, b: (a + b) if isinstance(a, tuple) and isinstance(b, tuple) else _raise_undef(),
    "abs": lambda x: abs(_num(x)),
    "max": lambda xs: sorted_values(_bool_coll(xs))[-1] if xs else _raise_undef(),
    "min": lambda xs: sorted_values(_bool_coll(xs))[0] if xs else _raise_undef(),
    "sum": lambda xs: _norm(sum((Fraction(_num(x)) for x in _bool_coll(xs)), Fraction(0))),
    "internal.member_2": _member,
    "print": lambda *a: True,
    "trace": lambda *a: True,
}


def _raise_undef():
    raise Undefined


# ---- further OPA builtins the gatekeeper policy library leans on (collections, objects, rounding); semantics as documented
# in OPA's policy reference (v1.x); type mismatches are undefined (non-strict builtin errors)
def _int_arg(x):
    x = _num(x)
    if isinstance(x, float) or (isinstance(x, Fraction) and x.denominator != 1):
        raise Undefined
    return int(x)


def _sort(xs):
    return tuple(sorted_values(_bool_coll(xs)))


def _obj(x):
    if not isinstance(x, RObj):
        raise Undefined
    return x


def _key_set(ks):
    if isinstance(ks, RObj):
        return list(ks.keys())
    return list(_bool_coll(ks))


def _has_key(keys, k):
    return any(equal(k, x) for x in keys)


def _object_union(a, b):
    a, b = _obj(a), _obj(b)
    out = dict(a)
    for k, v in b.items():
        if k in a and isinstance(a[k], RObj) and isinstance(v, RObj):
            out[k] = _object_union(a[k], v)
        else:
            out[k] = v
    return RObj(out)


def _numbers_range(a, b):
    a, b = _int_arg(a), _int_arg(b)
    return tuple(range(a, b + 1)) if a <= b else tuple(range(a, b - 1, -1))


def _array_slice(arr, lo, hi):
    if not isinstance(arr, tuple):
        raise Undefined
    lo, hi = _int_arg(lo), _int_arg(hi)
    lo = max(lo, 0)
    hi = min(hi, len(arr))
    return arr[lo:hi] if lo < hi else ()


def _round(x):
    x = _num(x)
    if isinstance(x, int):
        return x
    f = Fraction(x)
    import math
    return math.floor(f + Fraction(1, 2)) if f >= 0 else -math.floor(-f + Fraction(1, 2))      # half away from zero (Go math.Round)


def _floor(x):
    import math
    x = _num(x)
    return x if isinstance(x, int) else math.floor(Fraction(x))


def _ceil(x):
    import math
    x = _num(x)
    return x if isinstance(x, int) else math.ceil(Fraction(x))


def _format_int(x, base):
    base = _int_arg(base)
    if base not in (2, 8, 10, 16):
        raise Undefined
    n = _floor(x)
    digits = "0123456789abcdef"
    neg, n = n < 0, abs(n)
    out = ""
    while True:
        out = digits[n % base] + out
        n //= base
        if n == 0:
            break
    return ("-" if neg else "") + out


def _set_of_sets(xs):
    if not isinstance(xs, frozenset) or not all(isinstance(x, frozenset) for x in xs):
        raise Undefined
    return xs


def _union(xs):
    out = frozenset()
    for x in _set_of_sets(xs):
        out |= x
    return out


def _intersection(xs):
    xs = list(_set_of_sets(xs))
    if not xs:
        return frozenset()
    out = xs[0]
    for x in xs[1:]:
        out &= x
    return out


def _product(xs):
    acc = Fraction(1)
    for x in _bool_coll(xs):
        acc *= Fraction(_num(x))
    return _norm(acc)


def _type_name(x):
    return ["null", "boolean", "number", "string", "array", "object", "set"][type_rank(x)]


def _b64(x, enc):
    import base64
    import binascii
    x = _str(x)
    if enc:
        return base64.b64encode(x.encode()).decode()
    try:
        return base64.b64decode(x.encode(), validate=True).decode()
    except (binascii.Error, UnicodeDecodeError):
        raise Undefined


BUILTINS.update({
    "sort": _sort,
    "object.keys": lambda o: frozenset(_obj(o).keys()),
    "object.union": _object_union,
    "object.remove": lambda o, ks: (lambda keys: RObj({k: v for k, v in _obj(o).items() if not _has_key(keys, k)}))(_key_set(ks)),
    "object.filter": lambda o, ks: (lambda keys: RObj({k: v for k, v in _obj(o).items() if _has_key(keys, k)}))(_key_set(ks)),
    "numbers.range": _numbers_range,
    "array.slice": _array_slice,
    "array.reverse": lambda a: tuple(reversed(a)) if isinstance(a, tuple) else _raise_undef(),
    "strings.reverse": lambda x: _str(x)[::-1],
    "round": _round,
    "floor": _floor,
    "ceil": _ceil,
    "format_int": _format_int,
    "union": _union,
    "intersection": _intersection,
    "product": _product,
    "type_name": _type_name,
    "base64.encode": lambda x: _b64(x, True),
    "base64.decode": lambda x: _b64(x, False),
})


# --------------------------------------------------------------------------------------------------
# evaluator


# ---- scoping of comprehension locals: a variable declared with `:=` / `some` inside a comprehension is local to it and shadows
# an outer variable of the same name (OPA rewrites declared locals per closure).  The evaluator binds by name, so every
# such local is given a name of its own.
_scope_counter = [0]


def _pattern_vars(t, out):
    if not isinstance(t, tuple) or not t:
        return
    if t[0] == "var":
        out.append(t[1])
    elif t[0] in ("array", "set"):
        for x in t[1]:
            _pattern_vars(x, out)
    elif t[0] == "object":
        for _, v in t[1]:
            _pattern_vars(v, out)


def _rename(x, ren):
    if isinstance(x, tuple):
        if x and x[0] == "var" and len(x) == 2 and isinstance(x[1], str):
            return ("var", ren.get(x[1], x[1]))
        if x and x[0] == "some":
            return ("some", [ren.get(n, n) for n in x[1]])
        return tuple(_rename(y, ren) for y in x)
    if isinstance(x, list):
        return [_rename(y, ren) for y in x]
    return x


def _scope_fix(x):
    if isinstance(x, list):
        return [_scope_fix(y) for y in x]
    if not isinstance(x, tuple) or not x:
        return x
    if x[0] in ("scalar", "var", "some"):
        return x
    x = tuple(_scope_fix(y) for y in x)
    if x[0] == "every":
        decl = []
        _pattern_vars(x[1], decl)
        _pattern_vars(x[2], decl)
        for st in x[4]:
            if st[0] == "assign":
                _pattern_vars(st[1], decl)
            elif st[0] == "some":
                decl.extend(st[1])
        ren = {}
        for n in decl:
            if n and not n.startswith("$") and n not in ren:
                _scope_counter[0] += 1
                ren[n] = "%s$%d" % (n, _scope_counter[0])
        return ("every", _rename(x[1], ren), _rename(x[2], ren), x[3], _rename(x[4], ren))
    if x[0] in ("acompr", "scompr", "ocompr"):
        body = x[-1]
        decl = []
        for st in body:
            if st[0] == "assign":
                _pattern_vars(st[1], decl)
            elif st[0] == "some":
                decl.extend(st[1])
            elif st[0] == "somein":
                _pattern_vars(st[1], decl)
                _pattern_vars(st[2], decl)
        ren = {}
        for n in decl:
            if n and not n.startswith("$") and n not in ren:
                _scope_counter[0] += 1
                ren[n] = "%s$%d" % (n, _scope_counter[0])
        if ren:
            x = _rename(x, ren)
    return x


# ---- body ordering.  OPA's compiler reorders the expressions of a body so that every variable is bound before it is needed
# (ast/compile.go reorderBodyForSafety); templates rely on it, e.g. pkg/gator/fixtures/fixtures.go:461
#   selectors := [s | s = concat(":", [key, val]); val = obj.spec.selector[key]]
# The evaluator below runs a body left to right, so bodies are put into a safe order once, at load.  A body that is already safe
# in its written order is left alone.
def _vars_of(x, rules):
    out = set()
    _all_vars(x, rules, out, False)
    return out


def _is_plain_var(t, rules):
    return isinstance(t, tuple) and len(t) == 2 and t[0] == "var" and t[1] not in ("input", "data") and t[1] not in rules


def _all_vars(x, rules, out, deep=True):
    """deep=False: the variables a body itself names (the locals of its comprehensions are not visible outside them)."""
    if isinstance(x, tuple):
        if _is_plain_var(x, rules):
            out.add(x[1])
        elif x and x[0] == "some" and len(x) == 2 and isinstance(x[1], list):
            out.update(n for n in x[1] if isinstance(n, str))
        elif not deep and x and x[0] in ("acompr", "scompr", "ocompr"):
            return
        else:
            for y in x:
                _all_vars(y, rules, out, deep)
    elif isinstance(x, list):
        for y in x:
            _all_vars(y, rules, out, deep)


def _value_use(t, rules, scope, need, out):
    """Variables needed to evaluate term t as a value / variables that doing so binds (reference index positions)."""
    if not isinstance(t, tuple) or not t:
        return
    k = t[0]
    if k == "scalar":
        return
    if k == "var":
        if _is_plain_var(t, rules):
            need.add(t[1])
    elif k == "ref":
        if _is_plain_var(t[1], rules):
            need.add(t[1][1])
        else:
            _value_use(t[1], rules, scope, need, out)
        for a in t[2]:
            if _is_plain_var(a, rules):
                out.add(a[1])
            else:
                _value_use(a, rules, scope, need, out)
    elif k == "call":
        for a in t[2]:
            _value_use(a, rules, scope, need, out)
    elif k in ("array", "set"):
        for a in t[1]:
            _value_use(a, rules, scope, need, out)
    elif k == "object":
        for kk, vv in t[1]:
            _value_use(kk, rules, scope, need, out)
            _value_use(vv, rules, scope, need, out)
    elif k in ("acompr", "scompr", "ocompr"):
        inner = set()
        _all_vars(t[1:], rules, inner)
        need.update(v for v in inner if v in scope)      # its closure: variables of the enclosing bodies


def _pattern_use(t, rules, scope, need, out):
    if _is_plain_var(t, rules):
        out.add(t[1])
    elif isinstance(t, tuple) and t and t[0] == "array":
        for a in t[1]:
            _pattern_use(a, rules, scope, need, out)
    elif isinstance(t, tuple) and t and t[0] == "object":
        for kk, vv in t[1]:
            _value_use(kk, rules, scope, need, out)
            _pattern_use(vv, rules, scope, need, out)
    else:
        _value_use(t, rules, scope, need, out)


def _stmt_options(st, rules, scope):
    """[(need, binds)] -- the statement can run once ONE option's `need` is bound."""
    k = st[0]
    if k == "some":
        return [(set(), set())]
    if k in ("expr", "not"):
        need, out = set(), set()
        _value_use(st[1] if k == "expr" else (st[1][1] if isinstance(st[1], tuple) and st[1] and st[1][0] == "expr" else st[1]), rules, scope, need, out)
        if k == "not":
            if isinstance(st[1], tuple) and st[1] and st[1][0] in ("assign", "unify"):
                need, out = set(), set()
                _value_use(st[1][1], rules, scope, need, out)
                _value_use(st[1][2], rules, scope, need, out)
            need |= {v for v in out if not v.startswith("$")}
            out = set()
        return [(need, out)]
    if k in ("assign", "unify"):
        opts = []
        for lhs, rhs in ((st[1], st[2]), (st[2], st[1])):
            need, out = set(), set()
            _value_use(rhs, rules, scope, need, out)
            _pattern_use(lhs, rules, scope, need, out)
            opts.append((need - (out - need), out))
            if k == "assign":
                break
        return opts
    if k == "somein":
        need, out = set(), set()
        _value_use(st[3], rules, scope, need, out)
        for t in (st[1], st[2]):
            if t is not None:
                _pattern_use(t, rules, scope, need, out)
        return [(need, out)]
    if k == "every":
        need, out = set(), set()
        _value_use(st[3], rules, scope, need, out)
        inner = set()
        _all_vars(st[4], rules, inner)
        need.update(v for v in inner if v in scope)
        return [(need, set())]
    return [(set(), set())]


def _reorder_body(body, bound, rules, scope):
    scope = set(scope)
    _all_vars([st for st in body], rules, scope, False)      # closure candidates of nested comprehensions: everything named so far
    body = [_reorder_nested(st, rules, scope) for st in body]
    opts = [_stmt_options(st, rules, scope) for st in body]

    def runnable(i, b):
        for need, out in opts[i]:
            if need <= b:
                return out
        return None
    b = set(bound)
    ok = True
    for i in range(len(body)):
        out = runnable(i, b)
        if out is None:
            ok = False
            break
        b |= out
    if ok:
        return body
    b = set(bound)
    left = list(range(len(body)))
    order = []
    while left:
        for i in left:
            out = runnable(i, b)
            if out is not None:
                order.append(i)
                b |= out
                left.remove(i)
                break
        else:
            order.extend(left)          # nothing can run: keep what is left as written (the evaluator reports the unsafe variable)
            break
    return [body[i] for i in order]


def _reorder_nested(x, rules, scope):
    """Comprehension / every bodies inside a statement or term, innermost last."""
    if isinstance(x, list):
        return [_reorder_nested(y, rules, scope) for y in x]
    if not isinstance(x, tuple) or not x or x[0] in ("scalar", "var", "some"):
        return x
    if x[0] in ("acompr", "scompr", "ocompr"):
        head = tuple(_reorder_nested(y, rules, scope) for y in x[1:-1])
        return (x[0],) + head + (_reorder_body(x[-1], scope, rules, scope),)
    if x[0] == "every":
        inner = set(scope)
        _all_vars([x[1], x[2]], rules, inner)
        return ("every", x[1], x[2], _reorder_nested(x[3], rules, scope), _reorder_body(x[4], inner, rules, inner))
    return tuple(_reorder_nested(y, rules, scope) for y in x)


class Module:
    def __init__(self, src, libs=(), _registry=None):
        p = Parser(src)
        self.package, rules = p.parse_module()
        for r in rules:
            r.args = _scope_fix(r.args) if r.args is not None else None
            r.key, r.value, r.body = _scope_fix(r.key), _scope_fix(r.value), _scope_fix(r.body)
            r.els = [(_scope_fix(ev), _scope_fix(eb)) for ev, eb in r.els]
        self.imports = p.imports
        self.rules = {}
        for r in rules:
            self.rules.setdefault(r.name, []).append(r)
        # template libs: each is its own module under `package lib.<...>`; the entry point (and other libs) reach their
        # rules through `import data.lib.<...>` or the full data.lib path.  One registry per template.
        self.libs = _registry if _registry is not None else {}
        if _registry is None:
            for lsrc in libs:
                lm = Module(lsrc, _registry=self.libs)
                if lm.package != "lib" and not lm.package.startswith("lib."):
                    raise RegoError(f"rego_compile_error: lib package `{lm.package}` must begin with `lib`")
                self.libs[lm.package] = lm
            for mod in [self] + list(self.libs.values()):
                for alias, pkg in mod.imports.items():
                    if pkg not in self.libs:
                        raise RegoError(f"rego_compile_error: import data.{pkg}: the template has no lib with that package")
        self._check()
        names = set(self.rules) | set(self.imports)
        for rs in self.rules.values():
            for r in rs:
                bound = set()
                _all_vars(r.args or [], names, bound)
                r.body = _reorder_body(r.body, bound, names, bound)
                r.els = [(ev, _reorder_body(eb, bound, names, bound)) for ev, eb in r.els]
                for attr in ("key", "value"):
                    setattr(r, attr, _reorder_nested(getattr(r, attr), names, bound | _vars_of(r.body, names)))

    def lib_of_call(self, name):
        """`alias.fn` / `data.lib.<pkg>.fn` -> (lib module, fn) or None."""
        if "." not in name:
            return None
        first, rest = name.split(".", 1)
        if first in self.imports:
            return self.libs[self.imports[first]], rest
        if name.startswith("data.lib."):
            pkg, _, fn = name[5:].rpartition(".")
            if pkg in self.libs:
                return self.libs[pkg], fn
        return None

    def _check(self):
        # the compile-time checks the reference surfaces from AddTemplate as rego_* errors
        # (pkg/gator/fixtures/fixtures.go TemplateCompileError: body references undeclared `f`)
        for rs in self.rules.values():
            for r in rs:
                bound = set()
                for a in r.args or []:
                    _collect_vars(a, bound)
                self._check_body(r.body, set(bound), r)
                for _, eb in r.els:
                    self._check_body(eb, set(bound), r)

    def _check_body(self, body, bound, r):
        for st in body:
            if st[0] == "expr" and st[1][0] == "var":
                n = st[1][1]
                if n not in bound and n not in self.rules and not n.startswith("$") and n not in ("input", "data"):
                    raise RegoError(f"rego_unsafe_var_error: var {n} is unsafe (line {r.line})")
            for t in st[1:]:
                if isinstance(t, tuple):
                    _collect_vars(t, bound)
                elif isinstance(t, list):
                    bound.update(x for x in t if isinstance(x, str))


def _collect_vars(t, out):
    if not isinstance(t, tuple) or not t:
        return
    if t[0] == "var":
        out.add(t[1])
        return
    for x in t[1:]:
        if isinstance(x, tuple):
            _collect_vars(x, out)
        elif isinstance(x, list):
            for y in x:
                if isinstance(y, tuple) and y and isinstance(y[0], str):
                    _collect_vars(y, out)
                elif isinstance(y, tuple):
                    for z in y:
                        _collect_vars(z, out)


class Evaluator:
    """Top-down evaluation of one module against one `input` document (+ optional data.inventory)."""

    def __init__(self, module: Module, input_doc, data=None):
        self.m = module
        self.input = input_doc
        self.data = data if data is not None else RObj()
        self.cache = {}
        self.depth = 0
        self.subs = {}

    def sub(self, lib_module):
        ev = self.subs.get(id(lib_module))
        if ev is None:
            ev = self.subs[id(lib_module)] = Evaluator(lib_module, self.input, self.data)
        return ev

    # -- public ---------------------------------------------------------------------------------
    def rule_value(self, name):
        """Value of a complete rule / extent of a partial set or object rule.  Raises Undefined."""
        if name in self.cache:
            v = self.cache[name]
            if v is _UNDEF:
                raise Undefined
            return v
        rules = self.m.rules[name]
        kind = rules[0].kind
        try:
            if kind == "set":
                out = set()
                for r in rules:
                    for env in self.eval_body(r.body, {}):
                        for v, _ in self.eval_term(r.key, env):
                            out.add(v)
                val = frozenset(out)
            elif kind == "object":
                d = {}
                for r in rules:
                    for env in self.eval_body(r.body, {}):
                        for k, env2 in self.eval_term(r.key, env):
                            for v, _ in self.eval_term(r.value, env2):
                                d[k] = v
                val = RObj(d)
            elif kind == "complete":
                val = _UNDEF
                default = _UNDEF
                for r in rules:
                    if r.default:
                        default = next(self.eval_term(r.value, {}))[0]
                        continue
                    got = self._eval_rule_chain(r, {})
                    if got is not _UNDEF:
                        val = got
                        break
                if val is _UNDEF:
                    val = default
            else:
                raise RegoError(f"rego_type_error: {name} is a function")
        except RecursionError:
            raise RegoError("rego_recursion_error")
        self.cache[name] = val
        if val is _UNDEF:
            raise Undefined
        return val

    def _eval_rule_chain(self, r, env0):
        for env in self.eval_body(r.body, env0):
            for v, _ in self.eval_term(r.value, env):
                return v
        for ev, eb in r.els:
            for env in self.eval_body(eb, env0):
                if ev is None:
                    return True
                for v, _ in self.eval_term(ev, env):
                    return v
        return _UNDEF

    def call_function(self, name, argvals):
        rules = self.m.rules[name]
        for r in rules:
            if r.kind != "func" or len(r.args) != len(argvals):
                continue
            envs = [{}]
            for pat, v in zip(r.args, argvals):
                nxt = []
                for e in envs:
                    nxt.extend(self.unify_val(pat, v, e))
                envs = nxt
                if not envs:
                    break
            for e in envs:
                got = self._eval_rule_chain(r, e)
                if got is not _UNDEF:
                    return got
        raise Undefined

    # -- bodies ---------------------------------------------------------------------------------
    def eval_body(self, body, env, i=0):
        if i == len(body):
            yield env
            return
        st = body[i]
        k = st[0]
        if k == "some":
            yield from self.eval_body(body, env, i + 1)
        elif k == "not":
            if not any(True for _ in self.eval_stmt(st[1], env)):
                yield from self.eval_body(body, env, i + 1)
        else:
            for env2 in self.eval_stmt(st, env):
                yield from self.eval_body(body, env2, i + 1)

    def eval_stmt(self, st, env):
        k = st[0]
        if k == "expr":
            for v, env2 in self.eval_term(st[1], env):
                if v is not False:
                    yield env2
        elif k in ("assign", "unify"):
            yield from self.unify(st[1], st[2], env)
        elif k == "every":
            for coll, env2 in self.eval_term(st[3], env):
                if isinstance(coll, tuple):
                    items = list(enumerate(coll))
                elif isinstance(coll, RObj):
                    items = sort_items(coll)
                elif isinstance(coll, frozenset):
                    items = [(x, x) for x in sorted_values(coll)]
                else:
                    items = []        # a defined non-collection has no members (OPA generates them with `domain[k] = v`): vacuously true
                ok = True
                for kk, vv in items:
                    envs = [env2]
                    if st[1] is not None:
                        envs = list(self.unify_val(st[1], kk, env2))
                    envs = [e2 for e in envs for e2 in self.unify_val(st[2], vv, e)]
                    if not any(True for e in envs for _ in self.eval_body(st[4], e)):
                        ok = False
                        break
                if ok:
                    yield env2
        elif k == "somein":
            for coll, env2 in self.eval_term(st[3], env):
                if isinstance(coll, tuple):
                    items = list(enumerate(coll))
                elif isinstance(coll, RObj):
                    items = sort_items(coll)
                elif isinstance(coll, frozenset):
                    items = [(x, x) for x in sorted_values(coll)]
                else:
                    continue
                for kk, vv in items:
                    envs = [env2]
                    if st[1] is not None:
                        envs = list(self.unify_val(st[1], kk, env2))
                    for e in envs:
                        yield from self.unify_val(st[2], vv, e)
        else:
            raise RegoError(f"rego_unsupported: statement {k}")

    # -- unification ------------------------------------------------------------------------------
    def is_ground(self, t, env):
        k = t[0]
        if k == "scalar":
            return True
        if k == "var":
            n = t[1]
            return n in env or n in self.m.rules or n in ("input", "data")
        if k in ("array", "set"):
            return all(self.is_ground(x, env) for x in t[1])
        if k == "object":
            return all(self.is_ground(a, env) and self.is_ground(b, env) for a, b in t[1])
        return True  # refs/calls/comprehensions evaluate (refs may bind inner vars)

    def unify(self, a, b, env):
        ga, gb = self.is_ground(a, env), self.is_ground(b, env)
        if ga and gb:
            for va, e1 in self.eval_term(a, env):
                for vb, e2 in self.eval_term(b, e1):
                    if equal(va, vb):
                        yield e2
        elif gb:
            for vb, e1 in self.eval_term(b, env):
                yield from self.unify_val(a, vb, e1)
        elif ga:
            for va, e1 in self.eval_term(a, env):
                yield from self.unify_val(b, va, e1)
        else:
            if a[0] == "array" and b[0] == "array" and len(a[1]) == len(b[1]):
                envs = [env]
                for x, y in zip(a[1], b[1]):
                    envs = [e2 for e in envs for e2 in self.unify(x, y, e)]
                yield from envs
            else:
                raise RegoError("rego_unsafe_var_error: cannot unify two non-ground terms")

    def unify_val(self, pat, val, env):
        """Unify pattern term (may contain unbound vars) with a concrete value."""
        k = pat[0]
        if k == "var":
            n = pat[1]
            if n in env:
                if equal(env[n], val):
                    yield env
            elif n in self.m.rules or n in ("input", "data"):
                for v, e in self.eval_term(pat, env):
                    if equal(v, val):
                        yield e
            else:
                e = dict(env)
                e[n] = val
                yield e
        elif k == "array":
            if isinstance(val, tuple) and len(val) == len(pat[1]):
                envs = [env]
                for p, v in zip(pat[1], val):
                    envs = [e2 for e in envs for e2 in self.unify_val(p, v, e)]
                yield from envs
        elif k == "object":
            if isinstance(val, RObj) and len(val) == len(pat[1]):
                envs = [env]
                for pk, pv in pat[1]:
                    nxt = []
                    for e in envs:
                        for kv, e1 in self.eval_term(pk, e):
                            got = _lookup(val, kv, _UNDEF)
                            if got is not _UNDEF:
                                nxt.extend(self.unify_val(pv, got, e1))
                    envs = nxt
                yield from envs
        else:
            for v, e in self.eval_term(pat, env):
                if equal(v, val):
                    yield e

    # -- terms ----------------------------------------------------------------------------------
    def eval_term(self, t, env):
        """Yields (value, env') for every solution of term t."""
        k = t[0]
        if k == "scalar":
            yield t[1], env
        elif k == "var":
            n = t[1]
            if n in env:
                yield env[n], env
            elif n == "input":
                yield self.input, env
            elif n == "data":
                yield self.data, env
            elif n in self.m.rules:
                try:
                    yield self.rule_value(n), env
                except Undefined:
                    return
            else:
                raise RegoError(f"rego_unsafe_var_error: var {n} is unsafe")
        elif k == "ref":
            yield from self.eval_ref(t, env)
        elif k == "call":
            yield from self.eval_call(t, env)
        elif k == "array":
            yield from self._eval_seq(t[1], env, tuple)
        elif k == "set":
            yield from self._eval_seq(t[1], env, frozenset)
        elif k == "object":
            def rec(i, e, acc):
                if i == len(t[1]):
                    yield RObj(acc), e
                    return
                kt, vt = t[1][i]
                for kv, e1 in self.eval_term(kt, e):
                    for vv, e2 in self.eval_term(vt, e1):
                        yield from rec(i + 1, e2, acc + [(kv, vv)])
            yield from rec(0, env, [])
        elif k == "acompr":
            out = []
            for e in self.eval_body(t[2], env):
                for v, _ in self.eval_term(t[1], e):
                    out.append(v)
            yield tuple(out), env
        elif k == "scompr":
            out = set()
            for e in self.eval_body(t[2], env):
                for v, _ in self.eval_term(t[1], e):
                    out.add(v)
            yield frozenset(out), env
        elif k == "ocompr":
            d = {}
            for e in self.eval_body(t[3], env):
                for kv, e1 in self.eval_term(t[1], e):
                    for vv, _ in self.eval_term(t[2], e1):
                        d[kv] = vv
            yield RObj(d), env
        else:
            raise RegoError(f"rego_unsupported: term {k}")

    def _eval_seq(self, items, env, ctor):
        def rec(i, e, acc):
            if i == len(items):
                yield ctor(acc), e
                return
            for v, e1 in self.eval_term(items[i], e):
                yield from rec(i + 1, e1, acc + [v])
        yield from rec(0, env, [])

    def eval_ref(self, t, env):
        head, path = t[1], t[2]
        # <import alias>.<rule>...  and  data.lib.<pkg>.<rule>... : a rule of one of the template's libs
        if head[0] == "var" and head[1] not in env and path and path[0][0] == "scalar":
            lib = rest = None
            if head[1] in self.m.imports:
                lib, rest = self.m.libs[self.m.imports[head[1]]], path
            elif head[1] == "data" and path[0][1] == "lib":
                names = []
                for p in path:
                    if p[0] != "scalar" or not isinstance(p[1], str):
                        break
                    names.append(p[1])
                for n in range(len(names) - 1, 0, -1):
                    if ".".join(names[:n]) in self.m.libs:
                        lib, rest = self.m.libs[".".join(names[:n])], path[n:]
                        break
            if lib is not None:
                rn = rest[0][1]
                if rn not in lib.rules:
                    return
                if lib.rules[rn][0].kind == "func":
                    raise RegoError("rego_type_error: function used as ref")
                try:
                    base = self.sub(lib).rule_value(rn)
                except Undefined:
                    return
                yield from self._walk(base, rest[1:], 0, env)
                return
        # data.<...> refs: data.inventory... is the synced cache; data.lib unsupported
        if head[0] == "var" and head[1] not in env and head[1] in self.m.rules and self.m.rules[head[1]][0].kind == "func":
            raise RegoError("rego_type_error: function used as ref")
        for base, e in self.eval_term(head, env):
            yield from self._walk(base, path, 0, e)

    def _walk(self, cur, path, i, env):
        if i == len(path):
            yield cur, env
            return
        p = path[i]
        if p[0] == "var" and p[1] not in env and p[1] not in self.m.rules and p[1] not in ("input", "data"):
            # unbound var: iterate
            if isinstance(cur, tuple):
                items = list(enumerate(cur))
            elif isinstance(cur, RObj):
                items = sort_items(cur)
            elif isinstance(cur, frozenset):
                items = [(x, x) for x in sorted_values(cur)]
            else:
                return
            for kk, vv in items:
                e = dict(env)
                e[p[1]] = kk
                yield from self._walk(vv, path, i + 1, e)
            return
        if not self.is_ground(p, env):
            # pattern key, e.g. general_violation[{"msg": msg, "field": "containers"}]
            if isinstance(cur, frozenset):
                for x in sorted_values(cur):
                    for e in self.unify_val(p, x, env):
                        yield from self._walk(x, path, i + 1, e)
            elif isinstance(cur, RObj):
                for kk, vv in sort_items(cur):
                    for e in self.unify_val(p, kk, env):
                        yield from self._walk(vv, path, i + 1, e)
            return
        for kv, e in self.eval_term(p, env):
            if isinstance(cur, RObj):
                got = _lookup(cur, kv, _UNDEF)
                if got is not _UNDEF:
                    yield from self._walk(got, path, i + 1, e)
            elif isinstance(cur, tuple):
                if isinstance(kv, int) and not isinstance(kv, bool) and 0 <= kv < len(cur):
                    yield from self._walk(cur[kv], path, i + 1, e)
            elif isinstance(cur, frozenset):
                for x in cur:
                    if equal(x, kv):
                        yield from self._walk(x, path, i + 1, e)
                        break

    def eval_call(self, t, env):
        name, args = t[1], t[2]

        def rec(i, e, acc):
            if i == len(args):
                yield acc, e
                return
            for v, e1 in self.eval_term(args[i], e):
                yield from rec(i + 1, e1, acc + [v])

        target = self
        lib = self.m.lib_of_call(name)
        if lib is not None and lib[1] in lib[0].rules:
            target, name = self.sub(lib[0]), lib[1]
        user = name in target.m.rules and target.m.rules[name][0].kind == "func"
        nargs = len(target.m.rules[name][0].args) if user else None
        out_pat = None
        if user and len(args) == nargs + 1:
            out_pat, args = args[-1], args[:-1]
        for argvals, e in rec(0, env, []):
            try:
                if user:
                    self.depth += 1
                    if self.depth > 200:
                        raise RegoError("rego_recursion_error")
                    try:
                        v = target.call_function(name, argvals)
                    finally:
                        self.depth -= 1
                else:
                    fn = BUILTINS.get(name)
                    if fn is None:
                        raise RegoError(f"rego_type_error: undefined function {name}")
                    v = fn(*argvals)
            except Undefined:
                continue
            except (TypeError, ValueError, AttributeError):
                continue
            if out_pat is not None:
                yield from ((True, e2) for e2 in self.unify_val(out_pat, v, e))
            else:
                yield v, e


_UNDEF = object()


def eval_violations(module: Module, input_doc, data=None):
    """The reference's hook query reduced to one constraint: evaluate the template's `violation` partial
    set with `input = {"review": ..., "parameters": ...}` and return the SET of result objects.  Each
    element must be an object with a string "msg" (frameworks' rego driver rejects anything else) and an
    optional "details"."""
    ev = Evaluator(module, input_doc, data)
    if "violation" not in module.rules:
        raise RegoError("rego_compile_error: template has no `violation` rule")
    try:
        vs = ev.rule_value("violation")
    except Undefined:
        return []
    out = []
    for v in sorted_values(vs):
        if not isinstance(v, RObj) or not isinstance(v.get("msg"), str):
            raise RegoError("rego_type_error: violation element must be {\"msg\": string, ...}")
        out.append(v)
    return out
