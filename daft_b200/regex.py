"""Regular expressions for the device: a pattern is parsed here, on the host, and compiled into a small program for the Pike VM in
`daft_b200/csrc/jit_prelude.cuh` (`rx_exec`), which every row runs in O(len x instructions) with leftmost-first semantics -- the semantics of
the `regex` crate the reference uses (`src/daft-functions-utf8/src/{match_.rs,extract.rs,replace.rs}`) and of Python's `re`, which the oracle
uses as the restatement.

Supported: literals, `.`, classes `[a-z0-9_]` / `[^/]` (ASCII members; a negated class or `.` consumes one whole UTF-8 character), the escapes
`\\d \\w \\s \\D \\W \\S \\n \\t \\r` and escaped punctuation, greedy and lazy `* + ?` and `{m}` / `{m,}` / `{m,n}`, alternation, capturing and
non-capturing groups, `^` and `$` (whole value: no multi-line mode).  Refused with an error: back-references, look-around, named groups, flags,
Unicode classes / non-ASCII class members, and patterns that compile to more than 64 instructions."""
from __future__ import annotations

MAX_INST = 64
OP_MATCH, OP_CHAR, OP_ANY, OP_CLASS, OP_SPLIT, OP_JMP, OP_SAVE, OP_BOL, OP_EOL, OP_TAIL = range(10)

_D = [(48, 57)]
_W = [(48, 57), (65, 90), (95, 95), (97, 122)]
_S = [(9, 13), (32, 32)]


def _negate(ranges):
    out, lo = [], 0
    for a, b in sorted(ranges):
        if a > lo:
            out.append((lo, a - 1))
        lo = max(lo, b + 1)
    if lo <= 127:
        out.append((lo, 127))
    return out


class _Parser:
    def __init__(self, pattern: str):
        self.s, self.i, self.ngroups = pattern, 0, 0

    def fail(self, msg):
        raise ValueError(f"regex {self.s!r}: {msg}")

    def peek(self):
        return self.s[self.i] if self.i < len(self.s) else None

    def parse(self):
        node = self.alt()
        if self.i != len(self.s):
            self.fail("unbalanced ')'")
        return node

    def alt(self):
        branches = [self.concat()]
        while self.peek() == "|":
            self.i += 1
            branches.append(self.concat())
        node = branches[-1]
        for b in reversed(branches[:-1]):
            node = ("alt", b, node)
        return node

    def concat(self):
        items = []
        while self.peek() is not None and self.peek() not in "|)":
            items.append(self.repeat())
        return ("cat", items)

    def repeat(self):
        atom = self.atom()
        while self.peek() is not None and self.peek() in "*+?{":
            c = self.peek()
            if c == "{":
                j = self.s.find("}", self.i)
                body = self.s[self.i + 1:j] if j > 0 else ""
                parts = body.split(",")
                if j < 0 or not all(p.isdigit() or p == "" for p in parts) or not parts[0].isdigit() or len(parts) > 2:
                    self.fail("malformed {m,n}")
                lo = int(parts[0])
                hi = lo if len(parts) == 1 else (None if parts[1] == "" else int(parts[1]))
                self.i = j + 1
            else:
                lo, hi = {"*": (0, None), "+": (1, None), "?": (0, 1)}[c]
                self.i += 1
            greedy = True
            if self.peek() == "?":
                greedy = False
                self.i += 1
            if atom[0] in ("bol", "eol"):
                self.fail("a quantifier cannot follow an anchor")
            if hi is not None and (hi < lo or hi > 32):
                self.fail("repetition counts above 32 are not supported")
            atom = ("rep", atom, lo, hi, greedy)
        return atom

    def escape(self, in_class: bool):
        self.i += 1
        c = self.peek()
        if c is None:
            self.fail("trailing backslash")
        self.i += 1
        table = {"d": _D, "w": _W, "s": _S, "D": _negate(_D), "W": _negate(_W), "S": _negate(_S)}
        if c in table:
            return ("set", table[c], c in "DWS")     # third item: also matches every non-ASCII character
        if c in "ntr":
            return ("lit", {"n": "\n", "t": "\t", "r": "\r"}[c])
        if c.isalnum():
            self.fail(f"unsupported escape \\{c}")
        return ("lit", c)

    def atom(self):
        c = self.peek()
        if c == "(":
            self.i += 1
            idx = None
            if self.s.startswith("?:", self.i):
                self.i += 2
            elif self.peek() == "?":
                self.fail("only (?:...) groups are supported")
            else:
                self.ngroups += 1
                idx = self.ngroups
            inner = self.alt()
            if self.peek() != ")":
                self.fail("missing ')'")
            self.i += 1
            return ("group", idx, inner)
        if c == "[":
            return self.cls()
        if c == ".":
            self.i += 1
            return ("any",)
        if c == "^":
            self.i += 1
            return ("bol",)
        if c == "$":
            self.i += 1
            return ("eol",)
        if c == "\\":
            e = self.escape(False)
            return e if e[0] == "lit" else ("class", e[1], False, e[2])
        if c in "*+?{":
            self.fail(f"nothing to repeat before {c!r}")
        self.i += 1
        return ("lit", c)

    def cls(self):
        self.i += 1
        negated = False
        if self.peek() == "^":
            negated = True
            self.i += 1
        ranges, first, wide = [], True, False
        while True:
            c = self.peek()
            if c is None:
                self.fail("missing ']'")
            if c == "]" and not first:
                self.i += 1
                break
            first = False
            if c == "\\":
                e = self.escape(True)
                if e[0] == "set":
                    ranges += e[1]
                    wide = wide or e[2]
                    continue
                lo = e[1]
            else:
                self.i += 1
                lo = c
            hi = lo
            if self.peek() == "-" and self.i + 1 < len(self.s) and self.s[self.i + 1] != "]":
                self.i += 1
                hi = self.peek()
                if hi == "\\":
                    e = self.escape(True)
                    if e[0] != "lit":
                        self.fail("a class escape cannot end a range")
                    hi = e[1]
                else:
                    self.i += 1
            if ord(lo) > 127 or ord(hi) > 127:
                self.fail("non-ASCII characters inside a class are not supported")
            if ord(hi) < ord(lo):
                self.fail("reversed range in a class")
            ranges.append((ord(lo), ord(hi)))
        # `wide`: the positive set also holds every non-ASCII character (from \D \W \S)
        return ("class", ranges, negated, wide)


def parse(pattern: str):
    p = _Parser(pattern)
    return p.parse(), p.ngroups


class _Compiler:
    def __init__(self, group: int | None):
        self.inst, self.classes, self.group = [], [], group

    def emit(self, op, a=0, b=0, c=0):
        self.inst.append([op, a, b, c])
        return len(self.inst) - 1

    def node(self, n):
        kind = n[0]
        if kind == "cat":
            for x in n[1]:
                self.node(x)
        elif kind == "lit":
            for byte in n[1].encode("utf-8"):
                self.emit(OP_CHAR, byte)
        elif kind == "any":
            # the VM steps over BYTES: `.` takes the first byte of a character (not a continuation byte, not \n), TAIL the rest of it
            self.emit(OP_ANY)
            self.emit(OP_TAIL)
        elif kind == "class":
            _, ranges, negated, wide = n
            # the VM's class test: ASCII byte in the ranges (XOR negated); a non-ASCII character matches iff (negated XOR wide)
            self.classes.append(sorted(ranges))
            self.emit(OP_CLASS, len(self.classes) - 1, (1 if negated else 0) | (2 if (negated != wide) else 0))
            if negated != wide:   # the class can match a non-ASCII character: its continuation bytes follow
                self.emit(OP_TAIL)
        elif kind == "bol":
            self.emit(OP_BOL)
        elif kind == "eol":
            self.emit(OP_EOL)
        elif kind == "group":
            if n[1] is not None and n[1] == self.group:
                self.emit(OP_SAVE, 0)
                self.node(n[2])
                self.emit(OP_SAVE, 1)
            else:
                self.node(n[2])
        elif kind == "alt":
            s = self.emit(OP_SPLIT)
            self.inst[s][1] = len(self.inst)
            self.node(n[1])
            j = self.emit(OP_JMP)
            self.inst[s][2] = len(self.inst)
            self.node(n[2])
            self.inst[j][1] = len(self.inst)
        elif kind == "rep":
            _, e, lo, hi, greedy = n
            for _ in range(lo):
                self.node(e)
            if hi is None:
                s = self.emit(OP_SPLIT)
                body = len(self.inst)
                self.node(e)
                self.emit(OP_JMP, s)
                after = len(self.inst)
                self.inst[s][1], self.inst[s][2] = (body, after) if greedy else (after, body)
            else:
                splits = []
                for _ in range(hi - lo):
                    s = self.emit(OP_SPLIT)
                    splits.append(s)
                    self.inst[s][1 if greedy else 2] = len(self.inst)
                    self.node(e)
                for s in splits:
                    self.inst[s][2 if greedy else 1] = len(self.inst)
        else:
            raise AssertionError(kind)


def compile_program(pattern: str, group: int | None = None) -> bytes:
    """-> the VM program: [ninst] [ninst x (op, a, b, c)] [classes: (nranges, lo, hi, ...)*]; `group` selects the capture whose span the
    VM reports (0 = the whole match, None = none).  CLASS: a | c << 8 = byte offset of its table, b = flags."""
    ast, ngroups = parse(pattern)
    if group is not None and not 0 <= group <= ngroups:
        raise ValueError(f"regex {pattern!r}: no capture group {group}")
    c = _Compiler(group if group else None)
    if group == 0:
        c.emit(OP_SAVE, 0)
    c.node(ast)
    if group == 0:
        c.emit(OP_SAVE, 1)
    c.emit(OP_MATCH)
    if len(c.inst) > MAX_INST:
        raise ValueError(f"regex {pattern!r}: compiles to {len(c.inst)} instructions, the device program holds {MAX_INST}")
    out = bytearray([len(c.inst)])
    table, offsets = bytearray(), []
    base = 1 + 4 * len(c.inst)
    for ranges in c.classes:
        offsets.append(base + len(table))
        if len(ranges) > 32:
            raise ValueError(f"regex {pattern!r}: a class with more than 32 ranges")
        table.append(len(ranges))
        for lo, hi in ranges:
            table += bytes([lo, hi])
    for op, a, b, cc in c.inst:
        if op == OP_CLASS:   # (op, table offset low byte, flags: 1 = negated, 2 = matches non-ASCII characters, table offset high byte)
            out += bytes([op, offsets[a] & 255, b, offsets[a] >> 8])
        else:
            out += bytes([op, a, b, cc])
    return bytes(out + table)


def is_fully_anchored(pattern: str) -> bool:
    """^ ... $ at the top level: a match, if any, is the whole value"""
    ast, _ = parse(pattern)
    return ast[0] == "cat" and len(ast[1]) >= 2 and ast[1][0] == ("bol",) and ast[1][-1] == ("eol",)


def to_python(pattern: str) -> str:
    """the same pattern for Python's `re` (the oracle): `.` without newline is the default there too, `$` must not match before a trailing
    newline (the regex crate's does not), classes and escapes carry over"""
    ast, _ = parse(pattern)

    def esc(ch):
        import re
        return re.escape(ch)

    def cls(ranges, negated, wide):
        body = "".join(esc(chr(a)) if a == b else f"{esc(chr(a))}-{esc(chr(b))}" for a, b in ranges)
        if wide:
            body += "\\x80-\\U0010ffff"
        return f"[{'^' if negated else ''}{body}]" if body else ("[\\s\\S]" if negated else "(?!)")

    def un(n):
        k = n[0]
        if k == "cat": return "".join(un(x) for x in n[1])
        if k == "lit": return esc(n[1])
        if k == "any": return "."
        if k == "class": return cls(n[1], n[2], n[3])
        if k == "bol": return "\\A"
        if k == "eol": return "\\Z"
        if k == "group": return ("(" if n[1] is not None else "(?:") + un(n[2]) + ")"
        if k == "alt": return "(?:" + un(n[1]) + "|" + un(n[2]) + ")"
        if k == "rep":
            _, e, lo, hi, greedy = n
            q = "{%d,%s}" % (lo, "" if hi is None else hi)
            return "(?:" + un(e) + ")" + q + ("" if greedy else "?")
        raise AssertionError(k)
    return un(ast)
