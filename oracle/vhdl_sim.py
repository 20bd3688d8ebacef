"""vhdl_sim.py - TEST INFRASTRUCTURE ONLY: a cycle-based simulator for the VHDL subset the reference's RTL uses.

No VHDL simulator (ghdl, nvc, ...) can be had in this container, so the RTL entities
- src/cordic_dds.vhd, cordic_dds48.vhd, cordic_dds_scaled.vhd, cordic_atan2.vhd, int_multNxN_dsp48.vhd,
hamming_win.vhd, bh_win_{3,4,5,7}term.vhd, win_selector.vhd - had only been *restated* (oracle/bhw_oracle.c,
oracle/rtl_bitvec.py).  This module instead EXECUTES the reference's own source text: it parses the files where
they lie under /root/reference/src, elaborates an entity for given generics (generate statements, constants,
constant functions, component instances) and clocks it.  tests/golden/make_rtl_golden.py uses it to write
golden vectors that pin oracle/bhw_oracle.c (tests/test_rtl_vhdl_sim.py); nothing in the product imports it.

What is implemented is what those files need, with the semantics of the packages they `use`
(ieee.std_logic_1164, std_logic_arith, std_logic_signed):
  * types std_logic, std_logic_vector(h downto l), integer/natural, string, time, constrained array types
  * std_logic_signed arithmetic: "+" / "-" sign-extend both operands to the longer one (an integer or a
    std_logic operand takes the vector's length), comparisons are signed, SIGNED(a) * SIGNED(b) is a'length +
    b'length bits; "&", "not", "xor", "and", "or"; slices, indexing, 'left 'right 'high 'low 'length
  * aggregates: positional, (others => x), nested; x"..." and "0101" literals
  * concurrent: conditional signal assignment (`... when rising_edge(clk)`), processes, if/for generate,
    `entity work.x` instantiation with generic and port maps (ports alias the actual signals; `open` allowed)
  * sequential: if / elsif / else, case (choices with |, others), for loops, signal and variable assignment,
    null, return; functions without parameters or with positional parameters
  * an assignment whose value length differs from its target's raises - the check a real elaborator does
Every signal is two-valued here: 'U' / 'X' are not modelled (registers start at 0 instead of 'U'), `after`
delays are ignored (they are shorter than a clock), processes sensitive to a clock are evaluated on the rising
edge only (the asynchronous reset of cordic_dds48 / cordic_dds_scaled acts at the next edge).

The TAYLOR path (src/taylor_sincos.vhd, src/tay1_order.vhd, src/mults/mlt35x2{5,7}_dsp48e{1,2}.vhd) is executed the
same way, with two additions the reference does not carry itself:
  * ieee.math_real (MATH_PI, sin, cos, round) and the real -> integer conversion, evaluated in IEEE doubles with the
    C library's sin / cos and round-half-away-from-zero - what the usual simulators do for the ROM constant functions;
  * the Xilinx UNISIM primitives DSP48E1 / DSP48E2 (third party, not in the reference tree, no version pinned there):
    class Dsp48 below is a behavioural model of the configuration these files use - A/B/C/M/P pipeline registers,
    the X/Y/Z/W multiplexers selected by OPMODE, the four ALUMODE sums, synchronous resets, PCIN/PCOUT cascade with
    the 17-bit shift - written from the published port semantics (UG479 / UG579).  Any generic or port value outside
    that configuration raises instead of being guessed.
"""
from __future__ import annotations

import math
import os
import re

# --------------------------------------------------------------------------------------------- values


class BV:
    """std_logic_vector value: `w` bits, unsigned image `v`, declared bounds (left downto right)."""
    __slots__ = ("w", "v", "left", "right")

    def __init__(self, w, v, left=None, right=None):
        self.w = w
        self.v = v & ((1 << w) - 1) if w > 0 else 0
        self.left = w - 1 if left is None else left
        self.right = 0 if right is None else right

    def signed(self):
        return self.v - (1 << self.w) if self.w and (self.v >> (self.w - 1)) & 1 else self.v

    def bit(self, idx):
        pos = idx - self.right
        if not 0 <= pos < self.w:
            raise IndexError(f"index {idx} outside ({self.left} downto {self.right})")
        return SL((self.v >> pos) & 1)

    def slice(self, hi, lo):
        if hi < lo:
            return BV(0, 0, hi, lo)
        if hi > self.left or lo < self.right:
            raise IndexError(f"slice ({hi} downto {lo}) outside ({self.left} downto {self.right})")
        return BV(hi - lo + 1, self.v >> (lo - self.right), hi, lo)

    def __repr__(self):
        return f"BV{self.w}'{self.v:x}"


class UBV(BV):
    """the result of UNSIGNED(x): same bits, read as an unsigned number by conv_integer"""
    __slots__ = ()

    def signed(self):
        return self.v


class SL(int):
    """std_logic value, '0' or '1'."""


class Arr:
    """constrained array of anything; items are stored by ascending index, `down` = declared (hi downto lo)"""
    __slots__ = ("lo", "hi", "items", "down")

    def __init__(self, lo, hi, items, down=False):
        self.lo, self.hi, self.items, self.down = lo, hi, items, down

    def get(self, i):
        if not self.lo <= i <= self.hi:
            raise IndexError(f"array index {i} outside ({self.lo} to {self.hi})")
        return self.items[i - self.lo]

    def seq(self):
        """elements left to right"""
        return list(reversed(self.items)) if self.down else list(self.items)

    def sub(self, a, b):
        """slice (a downto b) / (b to a) with a >= b"""
        if a < b:
            return Arr(b, a, [], self.down)
        if b < self.lo or a > self.hi:
            raise IndexError(f"array slice ({a}, {b}) outside ({self.lo} to {self.hi})")
        return Arr(b, a, self.items[b - self.lo:a - self.lo + 1], self.down)

    def copy(self):
        return Arr(self.lo, self.hi, [x.copy() if isinstance(x, Arr) else x for x in self.items], self.down)


def as_bv(x, w=None):
    if isinstance(x, BV):
        return x
    if isinstance(x, SL):
        return BV(1, int(x))
    if isinstance(x, str):
        if not x or any(c not in "01" for c in x):
            raise TypeError(f"string {x!r} is not a bit string")
        return BV(len(x), int(x, 2))
    if isinstance(x, int) and w is not None:
        return BV(w, x)
    raise TypeError(f"cannot use {x!r} as a vector")


# --------------------------------------------------------------------------------------------- lexer

TOK = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<based>[xXbBoO]"[0-9a-fA-F_]*")
  | (?P<str>"(?:[^"]|"")*")
  | (?P<num>\d[\d_]*(?:\.\d+)?(?:[eE][+-]?\d+)?)
  | (?P<id>[A-Za-z][A-Za-z0-9_]*)
  | (?P<op><=|>=|=>|:=|/=|\*\*|<>|[-+*/&=<>():;,.|'])
""", re.X)

KEYWORDS = {"library", "use", "entity", "is", "generic", "port", "end", "architecture", "of", "begin", "signal",
            "constant", "type", "array", "to", "downto", "function", "return", "variable", "process", "if", "then",
            "elsif", "else", "case", "when", "others", "for", "in", "out", "inout", "loop", "generate", "map", "null",
            "not", "and", "or", "xor", "nand", "nor", "xnor", "mod", "rem", "abs", "after", "attribute", "open",
            "subtype", "component", "work", "all", "range", "buffer"}


def tokenize(text):
    toks, i, n = [], 0, len(text)
    while i < n:
        # character literal '0' (the tick of an attribute, x'left, is never followed by <char>')
        if text[i] == "'" and i + 2 < n and text[i + 2] == "'":
            toks.append(("char", text[i + 1]))
            i += 3
            continue
        m = TOK.match(text, i)
        if not m:
            raise SyntaxError(f"cannot tokenize at {text[i:i + 30]!r}")
        i = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        val = m.group(kind)
        if kind == "id":
            low = val.lower()
            toks.append(("kw", low) if low in KEYWORDS else ("id", low))
        elif kind == "num":
            toks.append(("num", val.replace("_", "")))
        elif kind == "str":
            toks.append(("str", val[1:-1]))
        elif kind == "based":
            base = {"x": 16, "b": 2, "o": 8}[val[0].lower()]
            digits = val[2:-1].replace("_", "")
            width = len(digits) * {16: 4, 2: 1, 8: 3}[base]
            toks.append(("bv", BV(width, int(digits, base) if digits else 0)))
        else:
            toks.append(("op", val))
    toks.append(("eof", None))
    return toks


# --------------------------------------------------------------------------------------------- parser


class Parser:
    def __init__(self, text):
        self.t = tokenize(text)
        self.i = 0

    # -- token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, val, k=0):
        return self.t[self.i + k][1] == val and self.t[self.i + k][0] in ("kw", "op")

    def eat(self, val=None):
        tok = self.t[self.i]
        if val is not None and not (tok[1] == val and tok[0] in ("kw", "op")):
            raise SyntaxError(f"expected {val!r}, got {tok!r} near token {self.i}: {self.t[max(0, self.i - 6):self.i + 4]}")
        self.i += 1
        return tok

    def opt(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def ident(self):
        tok = self.eat()
        if tok[0] != "id":
            raise SyntaxError(f"expected identifier, got {tok!r}: {self.t[max(0, self.i - 6):self.i + 4]}")
        return tok[1]

    # -- design file
    def design_file(self):
        units = {}
        while self.peek()[0] != "eof":
            if self.at("library") or self.at("use"):
                while not self.at(";"):
                    self.eat()
                self.eat(";")
            elif self.at("entity"):
                e = self.entity()
                units.setdefault(e["name"], {})["entity"] = e
            elif self.at("architecture"):
                a = self.architecture()
                units.setdefault(a["of"], {})["arch"] = a
            else:
                raise SyntaxError(f"unexpected {self.peek()!r}")
        return units

    def entity(self):
        self.eat("entity")
        name = self.ident()
        self.eat("is")
        generics, ports = [], []
        if self.opt("generic"):
            generics = self.interface_list()
            self.eat(";")
        if self.opt("port"):
            ports = self.interface_list()
            self.eat(";")
        self.eat("end")
        self.opt("entity")
        if self.peek()[0] == "id":
            self.ident()
        self.eat(";")
        return {"name": name, "generics": generics, "ports": ports}

    def interface_list(self):
        self.eat("(")
        items = []
        while True:
            self.opt("signal")
            names = [self.ident()]
            while self.opt(","):
                names.append(self.ident())
            self.eat(":")
            mode = "in"
            for m in ("in", "out", "inout", "buffer"):
                if self.opt(m):
                    mode = m
            typ = self.subtype()
            default = self.expr() if self.opt(":=") else None
            for nm in names:
                items.append({"name": nm, "mode": mode, "type": typ, "default": default})
            if not self.opt(";"):
                break
        self.eat(")")
        return items

    def subtype(self):
        name = self.ident() if self.peek()[0] == "id" else self.eat()[1]
        if self.at("(") and name in ("std_logic_vector", "signed", "unsigned"):
            self.eat("(")
            a = self.expr()
            down = True
            if self.opt("downto"):
                pass
            else:
                self.eat("to")
                down = False
            b = self.expr()
            self.eat(")")
            return ("slv", a, b, down)
        if self.at("range"):           # integer range a to b: the range is not enforced
            self.eat("range")
            self.expr()
            if not self.opt("to"):
                self.eat("downto")
            self.expr()
        return ("named", name)

    def architecture(self):
        self.eat("architecture")
        name = self.ident()
        self.eat("of")
        of = self.ident()
        self.eat("is")
        decls = self.declarations()
        self.eat("begin")
        stmts = self.concurrent_statements(("end",))
        self.eat("end")
        self.opt("architecture")
        if self.peek()[0] == "id":
            self.ident()
        self.eat(";")
        return {"name": name, "of": of, "decls": decls, "stmts": stmts}

    def declarations(self):
        out = []
        while True:
            if self.at("signal") or self.at("constant") or self.at("variable"):
                kind = self.eat()[1]
                names = [self.ident()]
                while self.opt(","):
                    names.append(self.ident())
                self.eat(":")
                typ = self.subtype()
                init = self.expr() if self.opt(":=") else None
                self.eat(";")
                for nm in names:
                    out.append((kind, nm, typ, init))
            elif self.at("type"):
                self.eat("type")
                nm = self.ident()
                self.eat("is")
                self.eat("array")
                self.eat("(")
                lo = self.expr()
                down = False
                if self.opt("downto"):
                    down = True
                else:
                    self.eat("to")
                hi = self.expr()
                self.eat(")")
                self.eat("of")
                el = self.subtype()
                self.eat(";")
                out.append(("type", nm, ("array", lo, hi, down, el), None))
            elif self.at("attribute"):
                while not self.at(";"):
                    self.eat()
                self.eat(";")
            elif self.at("function"):
                out.append(self.function())
            else:
                return out

    def function(self):
        self.eat("function")
        name = self.ident()
        params = []
        if self.at("("):
            params = self.interface_list()
        self.eat("return")
        rtype = self.subtype()
        self.eat("is")
        decls = self.declarations()
        self.eat("begin")
        body = self.sequential_statements(("end",))
        self.eat("end")
        self.opt("function")
        if self.peek()[0] == "id":
            self.ident()
        self.eat(";")
        return ("function", name, {"params": params, "rtype": rtype, "decls": decls, "body": body}, None)

    # -- concurrent statements
    def concurrent_statements(self, stop):
        out = []
        while not any(self.at(s) for s in stop):
            out.append(self.concurrent_statement())
        return out

    def concurrent_statement(self):
        label = None
        if self.peek()[0] == "id" and self.at(":", 1):
            label = self.ident()
            self.eat(":")
        if self.at("process"):
            self.eat("process")
            if self.opt("("):
                while not self.at(")"):
                    self.eat()
                self.eat(")")
            self.opt("is")
            decls = self.declarations()
            self.eat("begin")
            body = self.sequential_statements(("end",))
            self.eat("end")
            self.eat("process")
            if self.peek()[0] == "id":
                self.ident()
            self.eat(";")
            return ("process", label, decls, body)
        if self.at("if"):
            self.eat("if")
            cond = self.expr()
            self.eat("generate")
            decls = self.generate_decls()
            body = self.concurrent_statements(("end",))
            self.eat("end")
            self.eat("generate")
            if self.peek()[0] == "id":
                self.ident()
            self.eat(";")
            return ("ifgen", label, cond, body, decls)
        if self.at("for"):
            self.eat("for")
            var = self.ident()
            self.eat("in")
            a = self.expr()
            down = not self.opt("to")
            if down:
                self.eat("downto")
            b = self.expr()
            self.eat("generate")
            decls = self.generate_decls()
            body = self.concurrent_statements(("end",))
            self.eat("end")
            self.eat("generate")
            if self.peek()[0] == "id":
                self.ident()
            self.eat(";")
            return ("forgen", label, var, a, b, down, body, decls)
        component = label is not None and self.peek()[0] == "id" and (self.at("generic", 1) or self.at("port", 1))
        if self.at("entity") or component:
            if component:                         # label: DSP48E2 generic map (...) port map (...);
                ent = self.ident()
            else:
                self.eat("entity")
                self.eat("work")
                self.eat(".")
                ent = self.ident()
            gmap, pmap = [], []
            if self.opt("generic"):
                self.eat("map")
                gmap = self.assoc_list()
            if self.opt("port"):
                self.eat("map")
                pmap = self.assoc_list()
            self.eat(";")
            return ("inst", label, ent, gmap, pmap)
        # conditional signal assignment
        target = self.name()
        self.eat("<=")
        waves = []
        while True:
            val = self.expr()
            if self.opt("after"):
                self.expr()
                if self.peek()[0] == "id":
                    self.ident()          # time unit
            cond = None
            if self.opt("when"):
                cond = self.expr()
            waves.append((val, cond))
            if cond is not None and self.opt("else"):
                continue
            break
        self.eat(";")
        return ("cassign", label, target, waves)

    def generate_decls(self):
        """optional declarative part of a generate statement: [declarations] begin"""
        decls = self.declarations()
        if decls or self.at("begin"):
            self.eat("begin")
        return decls

    def assoc_list(self):
        self.eat("(")
        out = []
        while True:
            formal = self.ident()
            self.eat("=>")
            actual = None if self.opt("open") else self.expr()
            out.append((formal, actual))
            if not self.opt(","):
                break
        self.eat(")")
        return out

    # -- sequential statements
    def sequential_statements(self, stop):
        out = []
        while not any(self.at(s) for s in stop):
            out.append(self.sequential_statement())
        return out

    def sequential_statement(self):
        if self.peek()[0] == "id" and self.at(":", 1):
            self.ident()
            self.eat(":")
        if self.at("if"):
            self.eat("if")
            arms = []
            cond = self.expr()
            self.eat("then")
            arms.append((cond, self.sequential_statements(("elsif", "else", "end"))))
            while self.opt("elsif"):
                cond = self.expr()
                self.eat("then")
                arms.append((cond, self.sequential_statements(("elsif", "else", "end"))))
            if self.opt("else"):
                arms.append((None, self.sequential_statements(("end",))))
            self.eat("end")
            self.eat("if")
            self.eat(";")
            return ("if", arms)
        if self.at("case"):
            self.eat("case")
            sel = self.expr()
            self.eat("is")
            arms = []
            while self.opt("when"):
                choices = []
                while True:
                    choices.append(None if self.opt("others") else self.expr())
                    if not self.opt("|"):
                        break
                self.eat("=>")
                arms.append((choices, self.sequential_statements(("when", "end"))))
            self.eat("end")
            self.eat("case")
            self.eat(";")
            return ("case", sel, arms)
        if self.at("for"):
            self.eat("for")
            var = self.ident()
            self.eat("in")
            a = self.expr()
            down = not self.opt("to")
            if down:
                self.eat("downto")
            b = self.expr()
            self.eat("loop")
            body = self.sequential_statements(("end",))
            self.eat("end")
            self.eat("loop")
            if self.peek()[0] == "id":
                self.ident()
            self.eat(";")
            return ("for", var, a, b, down, body)
        if self.opt("null"):
            self.eat(";")
            return ("null",)
        if self.opt("return"):
            e = self.expr()
            self.eat(";")
            return ("return", e)
        target = self.name()
        if self.opt(":="):
            val = self.expr()
            self.eat(";")
            return ("vassign", target, val)
        self.eat("<=")
        val = self.expr()
        if self.opt("after"):
            self.expr()
            if self.peek()[0] == "id":
                self.ident()
        self.eat(";")
        return ("sassign", target, val)

    # -- expressions (VHDL precedence)
    def expr(self):
        left = self.relation()
        while self.peek()[0] == "kw" and self.peek()[1] in ("and", "or", "xor", "nand", "nor", "xnor"):
            op = self.eat()[1]
            left = ("bin", op, left, self.relation())
        return left

    def relation(self):
        left = self.simple()
        if self.peek()[0] == "op" and self.peek()[1] in ("=", "/=", "<", "<=", ">", ">="):
            op = self.eat()[1]
            left = ("bin", op, left, self.simple())
        return left

    def simple(self):
        if self.peek()[0] == "op" and self.peek()[1] in ("+", "-"):
            op = self.eat()[1]
            left = ("un", op, self.term())
        else:
            left = self.term()
        while self.peek()[0] == "op" and self.peek()[1] in ("+", "-", "&"):
            op = self.eat()[1]
            left = ("bin", op, left, self.term())
        return left

    def term(self):
        left = self.factor()
        while (self.peek()[0] == "op" and self.peek()[1] in ("*", "/")) or (self.peek()[0] == "kw" and self.peek()[1] in ("mod", "rem")):
            op = self.eat()[1]
            left = ("bin", op, left, self.factor())
        return left

    def factor(self):
        if self.opt("not"):
            return ("un", "not", self.primary())
        if self.opt("abs"):
            return ("un", "abs", self.primary())
        left = self.primary()
        if self.opt("**"):
            left = ("bin", "**", left, self.primary())
        return left

    def primary(self):
        tok = self.peek()
        if tok[0] == "num":
            self.eat()
            if self.peek()[0] == "id" and self.peek()[1] in ("ns", "ps", "us", "ms", "fs"):
                self.eat()
                return ("lit", 0)
            return ("lit", float(tok[1]) if ("." in tok[1] or "e" in tok[1].lower()) else int(tok[1]))
        if tok[0] == "bv":
            self.eat()
            return ("lit", tok[1])
        if tok[0] == "str":
            self.eat()
            return ("lit", tok[1])
        if tok[0] == "char":
            self.eat()
            return ("lit", SL(1 if tok[1] == "1" else 0))
        if self.at("("):
            self.eat("(")
            if self.at("others"):
                self.eat("others")
                self.eat("=>")
                v = self.expr()
                self.eat(")")
                return ("agg", [], v)
            first = self.expr()
            if self.at("=>"):                 # named association: (i => x, j => y, others => z)
                named, other = [], None
                key = first
                while True:
                    self.eat("=>")
                    val = self.expr()
                    if key is None:
                        other = val
                    else:
                        named.append((key, val))
                    if not self.opt(","):
                        break
                    key = None if self.opt("others") else self.expr()
                self.eat(")")
                return ("nagg", named, other)
            if self.opt(","):
                items = [first, self.expr()]
                while self.opt(","):
                    items.append(self.expr())
                self.eat(")")
                return ("agg", items, None)
            self.eat(")")
            return ("paren", first)
        if tok[0] == "id":
            return self.name()
        raise SyntaxError(f"unexpected token {tok!r} in expression: {self.t[max(0, self.i - 8):self.i + 4]}")

    def name(self):
        node = ("name", self.ident())
        while True:
            if self.at("("):
                self.eat("(")
                a = self.expr()
                if self.opt("downto"):
                    b = self.expr()
                    self.eat(")")
                    node = ("slice", node, a, b)
                elif self.at("to"):
                    self.eat("to")
                    b = self.expr()
                    self.eat(")")
                    node = ("slice", node, b, a)
                else:
                    args = [a]
                    while self.opt(","):
                        args.append(self.expr())
                    self.eat(")")
                    node = ("call", node, args)
            elif self.at("'") and self.peek(1)[0] in ("id", "kw"):
                self.eat("'")
                node = ("attr", node, self.eat()[1])
            else:
                return node


# --------------------------------------------------------------------------------------------- elaboration


class Sig:
    __slots__ = ("name", "typ", "val")

    def __init__(self, name, typ, val):
        self.name, self.typ, self.val = name, typ, val


class Return(Exception):
    def __init__(self, v):
        self.v = v


class Scope:
    """name -> Sig | constant value | ('type', t) | ('function', f); chained to the enclosing scope"""

    def __init__(self, parent=None):
        self.d, self.parent = {}, parent

    def get(self, k):
        s = self
        while s is not None:
            if k in s.d:
                return s.d[k]
            s = s.parent
        raise NameError(f"unknown name {k!r}")

    def has(self, k):
        s = self
        while s is not None:
            if k in s.d:
                return True
            s = s.parent
        return False


class Var:
    __slots__ = ("typ", "val")

    def __init__(self, typ, val):
        self.typ, self.val = typ, val


class Library:
    """The parsed design units of a set of files."""

    def __init__(self, paths):
        self.units = {}
        for p in paths:
            for name, u in Parser(open(p, encoding="latin-1").read()).design_file().items():
                self.units.setdefault(name, {}).update(u)


class Instance:
    """One elaborated entity: its signals, clocked and combinational statements and child instances."""

    def __init__(self, lib: Library, entity: str, generics: dict | None = None, port_sigs: dict | None = None, path=""):
        u = lib.units[entity.lower()]
        self.lib, self.path = lib, path + "/" + entity
        self.ent, arch = u["entity"], u["arch"]
        self.scope = Scope()
        generics = {k.lower(): v for k, v in (generics or {}).items()}
        for g in self.ent["generics"]:
            val = generics[g["name"]] if g["name"] in generics else (self.ev(g["default"], self.scope) if g["default"] is not None else 0)
            self.scope.d[g["name"]] = val
        self.ports = {}
        for p in self.ent["ports"]:
            typ = self.resolve_type(p["type"], self.scope)
            sig = (port_sigs or {}).get(p["name"])
            if sig is None:
                sig = Sig(self.path + "." + p["name"], typ, self.default(typ))
            else:
                self.check_type(sig, typ, p["name"])
            self.scope.d[p["name"]] = sig
            self.ports[p["name"]] = sig
        self.clocked, self.comb, self.children = [], [], []
        self.declare(arch["decls"], self.scope)
        self.elab_stmts(arch["stmts"], self.scope)

    # ---- types
    def resolve_type(self, t, scope):
        if t[0] == "slv":
            a, b = self.ev(t[1], scope), self.ev(t[2], scope)
            if not t[3]:
                raise NotImplementedError("ascending std_logic_vector")
            return ("slv", a, b)
        if t[0] == "array":
            lo, hi = self.ev(t[1], scope), self.ev(t[2], scope)
            if t[3]:
                lo, hi = hi, lo
            return ("array", lo, hi, self.resolve_type(t[4], scope), bool(t[3]))
        name = t[1]
        if name in ("std_logic", "std_ulogic", "bit"):
            return ("sl",)
        if name in ("integer", "natural", "positive"):
            return ("int",)
        if name in ("string", "time", "real", "boolean"):
            return (name,)
        got = scope.get(name)
        if isinstance(got, tuple) and got[0] == "type":
            return got[1]
        raise TypeError(f"unknown type {name}")

    def default(self, typ):
        if typ[0] == "slv":
            return BV(max(0, typ[1] - typ[2] + 1), 0, typ[1], typ[2])
        if typ[0] == "sl":
            return SL(0)
        if typ[0] == "array":
            return Arr(typ[1], typ[2], [self.default(typ[3]) for _ in range(typ[1], typ[2] + 1)], typ[4])
        if typ[0] == "int":
            return 0
        if typ[0] == "real":
            return 0.0
        return None

    def check_type(self, sig, typ, what):
        if typ[0] == "slv" and isinstance(sig.val, BV) and sig.val.w != typ[1] - typ[2] + 1:
            raise TypeError(f"{self.path}: port {what} is {typ[1] - typ[2] + 1} bits, actual {sig.name} is {sig.val.w}")

    def conform(self, val, typ, what="value"):
        """value as an object of type typ (length check for vectors, aggregates expanded)"""
        if typ[0] == "slv":
            w = typ[1] - typ[2] + 1
            if isinstance(val, tuple) and val[0] == "others":
                bit = int(val[1])
                return BV(w, ((1 << w) - 1) if bit else 0, typ[1], typ[2])
            if isinstance(val, tuple) and val[0] == "named":
                out = ((1 << w) - 1) if (val[2] is not None and int(val[2])) else 0
                if val[2] is None and len(val[1]) != w:
                    raise ValueError(f"{self.path}: {what}: aggregate does not cover all {w} bits")
                for idx, bit in val[1]:
                    pos = idx - typ[2]
                    if not 0 <= pos < w:
                        raise IndexError(f"{self.path}: {what}: aggregate index {idx} out of range")
                    out = (out & ~(1 << pos)) | (int(bit) << pos)
                return BV(w, out, typ[1], typ[2])
            bv = as_bv(val)
            if bv.w != w:
                raise ValueError(f"{self.path}: {what}: length {bv.w} assigned to {w} bits")
            return BV(w, bv.v, typ[1], typ[2])
        if typ[0] == "sl":
            if isinstance(val, BV) and val.w == 1:
                return SL(val.v)
            if not isinstance(val, SL):
                raise TypeError(f"{self.path}: {what}: {val!r} is not a std_logic")
            return val
        if typ[0] == "array":
            n = typ[2] - typ[1] + 1
            down = typ[4]
            if isinstance(val, tuple) and val[0] == "others":
                return Arr(typ[1], typ[2], [self.conform(val[1], typ[3], what) for _ in range(n)], down)
            if isinstance(val, Arr):
                val = ("positional", val.seq())                      # arrays are assigned left to right
            if isinstance(val, tuple) and val[0] == "positional":
                if len(val[1]) != n:
                    raise ValueError(f"{self.path}: {what}: {len(val[1])} elements for an array of {n}")
                items = [self.conform(x, typ[3], what) for x in val[1]]
                return Arr(typ[1], typ[2], list(reversed(items)) if down else items, down)
            raise TypeError(f"{self.path}: {what}: {val!r} is not an array")
        if typ[0] == "int":
            if isinstance(val, (BV, SL)) and not isinstance(val, bool):
                raise TypeError(f"{self.path}: {what}: vector assigned to an integer")
            return int(val)
        return val

    # ---- declarations
    def declare(self, decls, scope):
        for kind, name, typ, init in decls:
            if kind == "type":
                scope.d[name] = ("type", self.resolve_type(typ, scope))
            elif kind == "function":
                scope.d[name] = ("function", typ, scope)
            elif kind == "constant":
                t = self.resolve_type(typ, scope)
                scope.d[name] = self.conform(self.ev(init, scope), t, f"constant {name}")
            elif kind == "signal":
                t = self.resolve_type(typ, scope)
                val = self.default(t) if init is None else self.conform(self.ev(init, scope), t, f"signal {name}")
                scope.d[name] = Sig(self.path + "." + name, t, val)
            elif kind == "variable":
                t = self.resolve_type(typ, scope)
                scope.d[name] = Var(t, self.default(t) if init is None else self.conform(self.ev(init, scope), t, f"variable {name}"))

    # ---- concurrent statements
    def elab_stmts(self, stmts, scope):
        for st in stmts:
            k = st[0]
            if k == "process":
                ps = Scope(scope)
                self.declare(st[2], ps)
                (self.clocked if mentions_edge(st[3]) else self.comb).append(("proc", st[3], ps))
            elif k == "cassign":
                clocked = any(c is not None and mentions_edge(c) for _, c in st[3])
                (self.clocked if clocked else self.comb).append(("cassign", st[2], st[3], scope))
            elif k == "ifgen":
                if self.truth(self.ev(st[2], scope)):
                    s2 = Scope(scope)
                    self.declare(st[4], s2)
                    self.elab_stmts(st[3], s2)
            elif k == "forgen":
                a, b = self.ev(st[3], scope), self.ev(st[4], scope)
                rng = range(a, b - 1, -1) if st[5] else range(a, b + 1)
                for i in rng:
                    s2 = Scope(scope)
                    s2.d[st[2]] = i
                    self.declare(st[7], s2)
                    self.elab_stmts(st[6], s2)
            elif k == "inst":
                self.instantiate(st, scope)
            else:
                raise NotImplementedError(k)

    def instantiate(self, st, scope):
        _, label, ent, gmap, pmap = st
        gens = {f: self.ev(a, scope) for f, a in gmap}
        prim = PRIMITIVES.get(ent)
        if prim is not None:
            modes = {n: m for n, (m, _) in prim.PORTS.items()}
        elif ent in self.lib.units:
            modes = {p["name"]: p["mode"] for p in self.lib.units[ent]["entity"]["ports"]}
        else:
            raise NameError(f"{self.path}: no entity or primitive {ent!r}")
        port_sigs, drives, taps = {}, [], []
        for formal, actual in pmap:
            if actual is None:
                continue
            if formal not in modes:
                raise NameError(f"{self.path}: {ent} has no port {formal!r}")
            if actual[0] == "name" and scope.has(actual[1]) and isinstance(scope.get(actual[1]), Sig):
                port_sigs[formal] = scope.get(actual[1])          # the port IS the actual signal
            elif modes[formal] == "in":
                drives.append((formal, actual))                     # an expression drives an input port
            elif actual[0] == "slice":
                taps.append((formal, actual))                       # an output port drives a slice of a signal
            else:
                raise NotImplementedError(f"{self.path}: output port {formal} mapped to an expression")
        path = self.path + "/" + (label or ent)
        child = prim(self, gens, port_sigs, path) if prim is not None else Instance(self.lib, ent, gens, port_sigs, path)
        for formal, actual in drives:
            self.comb.append(("drive", child.ports[formal], actual, scope))
        for formal, actual in taps:
            self.comb.append(("tap", child.ports[formal], actual, scope))
        self.children.append(child)

    # ---- expression evaluation
    def truth(self, v):
        if isinstance(v, bool):
            return v
        raise TypeError(f"{self.path}: condition is {v!r}, not a boolean")

    def ev(self, e, scope, hint=None):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "paren":
            return self.ev(e[1], scope, hint)
        if k == "agg":
            if e[2] is not None:
                return ("others", self.ev(e[2], scope))
            return ("positional", [self.ev(x, scope) for x in e[1]])
        if k == "nagg":
            return ("named", [(self.ev(i, scope), self.ev(v, scope)) for i, v in e[1]], None if e[2] is None else self.ev(e[2], scope))
        if k == "name":
            v = scope.get(e[1]) if scope.has(e[1]) else self.builtin_name(e[1])
            if isinstance(v, (Sig, Var)):
                return v.val
            if isinstance(v, tuple) and v and v[0] == "function":      # a parameterless function is called by its name
                return self.call_function(v, [])
            return v
        if k == "attr":
            base = self.ev(e[1], scope)
            a = e[2]
            if isinstance(base, BV):
                return {"left": base.left, "right": base.right, "high": base.left, "low": base.right, "length": base.w}[a]
            if isinstance(base, Arr):
                return {"left": base.hi if base.down else base.lo, "right": base.lo if base.down else base.hi,
                        "high": base.hi, "low": base.lo, "length": base.hi - base.lo + 1}[a]
            raise TypeError(f"attribute {a} of {base!r}")
        if k == "slice":
            base = self.ev(e[1], scope)
            hi, lo = self.ev(e[2], scope), self.ev(e[3], scope)
            if isinstance(base, Arr):
                return base.sub(hi, lo)
            return as_bv(base).slice(hi, lo)
        if k == "call":
            return self.call(e, scope)
        if k == "un":
            v = self.ev(e[2], scope)
            if e[1] == "not":
                if isinstance(v, SL):
                    return SL(1 - int(v))
                if isinstance(v, bool):
                    return not v
                b = as_bv(v)
                return BV(b.w, ~b.v)
            if e[1] == "-":
                if isinstance(v, (BV,)):
                    return BV(v.w, -v.signed())
                return -v
            if e[1] == "+":
                return v
            if e[1] == "abs":
                return abs(v)
        if k == "bin":
            return self.binop(e[1], self.ev(e[2], scope), self.ev(e[3], scope))
        raise NotImplementedError(e)

    def builtin_name(self, name):
        if name in ("true", "false"):
            return name == "true"
        if name == "math_pi":
            return math.pi
        raise NameError(f"{self.path}: unknown name {name!r}")

    def call(self, e, scope):
        fn, args = e[1], e[2]
        if fn[0] == "name":
            nm = fn[1]
            if scope.has(nm):
                target = scope.get(nm)
                if isinstance(target, tuple) and target[0] == "function":
                    return self.call_function(target, [self.ev(a, scope) for a in args])
                if isinstance(target, tuple) and target[0] == "type":
                    return self.conform(self.ev(args[0], scope), target[1], f"conversion to {nm}")
                base = target.val if isinstance(target, (Sig, Var)) else target
                return self.index(base, self.ev(args[0], scope))
            if nm == "rising_edge":
                return True                      # statements that mention it are evaluated on the rising edge only
            if nm == "unsigned":
                b = as_bv(self.ev(args[0], scope))
                return UBV(b.w, b.v, b.left, b.right)
            if nm in ("signed", "std_logic_vector"):
                b = as_bv(self.ev(args[0], scope))
                return BV(b.w, b.v, b.left, b.right)
            if nm == "conv_signed":
                return BV(self.ev(args[1], scope), int(self.ev(args[0], scope)))
            if nm == "sxt":                                   # std_logic_arith: sign-extend to n bits
                b, n = as_bv(self.ev(args[0], scope)), self.ev(args[1], scope)
                if n < b.w:
                    raise ValueError("SXT to fewer bits")
                return BV(n, b.signed())
            if nm == "real":
                return float(self.as_int(self.ev(args[0], scope)))
            if nm == "integer":                               # real -> integer rounds to nearest, halves away from zero
                v = self.ev(args[0], scope)
                return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1) if isinstance(v, float) else int(v)
            if nm == "round":
                v = self.ev(args[0], scope)
                return float(math.floor(abs(v) + 0.5)) * (1.0 if v >= 0 else -1.0)
            if nm in ("sin", "cos"):
                return getattr(math, nm)(self.ev(args[0], scope))
            if nm == "conv_std_logic_vector":
                return BV(self.ev(args[1], scope), int(self.as_int(self.ev(args[0], scope))))
            if nm in ("conv_integer", "to_integer"):
                return self.as_int(self.ev(args[0], scope))
            raise NameError(f"{self.path}: unknown function {nm!r}")
        base = self.ev(fn, scope)
        return self.index(base, self.ev(args[0], scope))

    def as_int(self, v):
        return v.signed() if isinstance(v, BV) else int(v)

    def index(self, base, i):
        if isinstance(base, Arr):
            return base.get(int(i))
        if isinstance(base, BV):
            return base.bit(int(i))
        raise TypeError(f"{self.path}: cannot index {base!r}")

    def call_function(self, target, argv):
        _, f, defscope = target
        fs = Scope(defscope)
        for p, a in zip(f["params"], argv):
            fs.d[p["name"]] = a
        self.declare(f["decls"], fs)
        try:
            self.exec_seq(f["body"], fs, None)
        except Return as r:
            return self.conform(r.v, self.resolve_type(f["rtype"], fs), "function result") if f["rtype"][0] != "named" or f["rtype"][1] not in ("integer", "natural") else int(r.v)
        raise RuntimeError("function fell off its end")

    def binop(self, op, a, b):
        if op in ("+", "-"):
            if isinstance(a, (BV, SL)) or isinstance(b, (BV, SL)) and not isinstance(a, (int,)) or isinstance(a, SL) or isinstance(b, BV):
                pass
            if isinstance(a, BV) or isinstance(b, BV):
                # std_logic_signed: both operands sign-extended to the longer one; an integer or a std_logic
                # operand adopts the vector's length; the result has that length
                wa = a.w if isinstance(a, BV) else 0
                wb = b.w if isinstance(b, BV) else 0
                w = max(wa, wb)
                va = a.signed() if isinstance(a, BV) else int(a)
                vb = b.signed() if isinstance(b, BV) else int(b)
                return BV(w, va + vb if op == "+" else va - vb)
            if isinstance(a, float) or isinstance(b, float):
                return a + b if op == "+" else a - b
            return int(a) + int(b) if op == "+" else int(a) - int(b)
        if op == "&":
            if isinstance(a, Arr) or isinstance(b, Arr):
                return ("positional", (a.seq() if isinstance(a, Arr) else [a]) + (b.seq() if isinstance(b, Arr) else [b]))
            x, y = as_bv(a), as_bv(b)
            return BV(x.w + y.w, (x.v << y.w) | y.v)
        if op == "*":
            if isinstance(a, BV) and isinstance(b, BV):
                return BV(a.w + b.w, a.signed() * b.signed())     # SIGNED * SIGNED: a'length + b'length bits
            if isinstance(a, BV) or isinstance(b, BV):
                raise NotImplementedError("vector * integer")
            return a * b
        if op == "/":
            return int(a) // int(b) if not isinstance(a, float) and not isinstance(b, float) else a / b
        if op == "mod":
            return int(a) % int(b)
        if op == "rem":
            return int(a) - int(b) * int(int(a) / int(b))
        if op == "**":
            return a ** b
        if op in ("and", "or", "xor", "nand", "nor", "xnor"):
            if isinstance(a, bool) and isinstance(b, bool):
                r = {"and": a and b, "or": a or b, "xor": a != b, "nand": not (a and b), "nor": not (a or b), "xnor": a == b}[op]
                return r
            if isinstance(a, SL) and isinstance(b, SL):
                x, y = int(a), int(b)
                r = {"and": x & y, "or": x | y, "xor": x ^ y, "nand": 1 - (x & y), "nor": 1 - (x | y), "xnor": 1 - (x ^ y)}[op]
                return SL(r)
            x, y = as_bv(a), as_bv(b)
            if x.w != y.w:
                raise ValueError("logical operator on vectors of different length")
            r = {"and": x.v & y.v, "or": x.v | y.v, "xor": x.v ^ y.v}[op.lstrip("n") if op in ("nand", "nor") else op] if op != "xnor" else ~(x.v ^ y.v)
            if op in ("nand", "nor"):
                r = ~r
            return BV(x.w, r)
        if op in ("=", "/=", "<", "<=", ">", ">="):
            if isinstance(a, str) and isinstance(b, str) and not (isinstance(a, BV) or isinstance(b, BV)):
                x, y = a, b
            elif isinstance(a, (BV,)) or isinstance(b, (BV,)):
                x = a.signed() if isinstance(a, BV) else (as_bv(a).signed() if isinstance(a, str) else int(a))
                y = b.signed() if isinstance(b, BV) else (as_bv(b).signed() if isinstance(b, str) else int(b))
                if op in ("=", "/=") and isinstance(a, BV) and isinstance(b, (BV, str)):
                    bb = as_bv(b)
                    if a.w != bb.w:
                        raise ValueError("equality of vectors of different length")
            else:
                x, y = a, b
            return {"=": x == y, "/=": x != y, "<": x < y, "<=": x <= y, ">": x > y, ">=": x >= y}[op]
        raise NotImplementedError(op)

    # ---- sequential execution
    def exec_seq(self, stmts, scope, pending):
        for st in stmts:
            k = st[0]
            if k == "if":
                for cond, body in st[1]:
                    if cond is None or self.truth(self.ev(cond, scope)):
                        self.exec_seq(body, scope, pending)
                        break
            elif k == "case":
                sel = self.ev(st[1], scope)
                done = False
                for choices, body in st[2]:
                    for c in choices:
                        if c is None or self.binop("=", sel, self.ev(c, scope)):
                            self.exec_seq(body, scope, pending)
                            done = True
                            break
                    if done:
                        break
            elif k == "for":
                a, b = self.ev(st[2], scope), self.ev(st[3], scope)
                rng = range(a, b - 1, -1) if st[4] else range(a, b + 1)
                s2 = Scope(scope)
                for i in rng:
                    s2.d[st[1]] = i
                    self.exec_seq(st[5], s2, pending)
            elif k == "sassign":
                self.assign(st[1], self.ev(st[2], scope), scope, pending)
            elif k == "vassign":
                self.assign(st[1], self.ev(st[2], scope), scope, None)
            elif k == "return":
                raise Return(self.ev(st[1], scope))
            elif k == "null":
                pass
            else:
                raise NotImplementedError(k)

    def assign(self, target, val, scope, pending):
        """target: name, name(i), name(h downto l), name(i)(h downto l), name(i)(j).  pending: list of deferred
        signal updates (signal assignment) or None (variable assignment: immediate)."""
        path = []
        node = target
        while node[0] != "name":
            if node[0] == "slice":
                path.append(("slice", self.ev(node[2], scope), self.ev(node[3], scope)))
                node = node[1]
            elif node[0] == "call":
                path.append(("index", self.ev(node[2][0], scope)))
                node = node[1]
            else:
                raise NotImplementedError(f"assignment target {node}")
        path.reverse()
        obj = scope.get(node[1])
        if not isinstance(obj, (Sig, Var)):
            raise TypeError(f"{self.path}: {node[1]} is not assignable")
        if isinstance(obj, Sig) and pending is None:
            raise TypeError(f"{self.path}: := on signal {node[1]}")
        # type of the addressed element
        typ = obj.typ
        for p in path:
            if p[0] == "index":
                typ = typ[3] if typ[0] == "array" else ("sl",)
            else:
                if typ[0] != "slv":
                    raise TypeError("slice of a non-vector")
                typ = ("slv", p[1], p[2])
        v = self.conform(val, typ, f"assignment to {node[1]}")
        if isinstance(obj, Var) or pending is None:
            obj.val = store(obj.val, path, v)
        else:
            pending.append((obj, path, v))

    # ---- simulation
    def all_instances(self):
        out = [self]
        for c in self.children:
            out += c.all_instances()
        return out


def store(cur, path, v):
    """new value of an object after writing v at path"""
    if not path:
        return v
    p = path[0]
    if p[0] == "index":
        if isinstance(cur, Arr):
            new = Arr(cur.lo, cur.hi, list(cur.items), cur.down)
            new.items[p[1] - cur.lo] = store(cur.get(p[1]), path[1:], v)
            return new
        if isinstance(cur, BV):
            pos = p[1] - cur.right
            if not 0 <= pos < cur.w:
                raise IndexError("bit index out of range")
            bit = int(v)
            return BV(cur.w, (cur.v & ~(1 << pos)) | (bit << pos), cur.left, cur.right)
        raise TypeError("index store into scalar")
    hi, lo = p[1], p[2]
    if hi < lo:
        return cur
    if hi > cur.left or lo < cur.right:
        raise IndexError(f"slice ({hi} downto {lo}) outside ({cur.left} downto {cur.right})")
    sub = v if len(path) == 1 else store(cur.slice(hi, lo), path[1:], v)
    w = hi - lo + 1
    sh = lo - cur.right
    mask = ((1 << w) - 1) << sh
    return BV(cur.w, (cur.v & ~mask) | ((sub.v << sh) & mask), cur.left, cur.right)


def mentions_edge(node):
    if isinstance(node, tuple):
        if len(node) >= 2 and node[0] == "name" and node[1] == "rising_edge":
            return True
        return any(mentions_edge(x) for x in node)
    if isinstance(node, list):
        return any(mentions_edge(x) for x in node)
    return False


def same(a, b):
    if isinstance(a, BV) and isinstance(b, BV):
        return a.w == b.w and a.v == b.v
    if isinstance(a, Arr) and isinstance(b, Arr):
        return all(same(x, y) for x, y in zip(a.items, b.items))
    return a == b


class Simulator:
    """Clocks an elaborated design.  `step(**inputs)` = one rising edge with these values on the top-level
    input ports (they keep their values until changed); returns nothing - read outputs with `get`."""

    def __init__(self, lib: Library, entity: str, generics: dict | None = None):
        self.top = Instance(lib, entity, generics)
        self.insts = self.top.all_instances()
        self.settle()

    def set(self, **inputs):
        for k, v in inputs.items():
            sig = self.top.ports[k.lower()]
            sig.val = self.top.conform(BV(sig.val.w, v) if isinstance(sig.val, BV) and isinstance(v, int) and not isinstance(v, SL)
                                       else (SL(v) if isinstance(sig.val, SL) else v), sig.typ, f"input {k}")

    def get(self, name, signed=True):
        v = self.top.ports[name.lower()].val
        if isinstance(v, BV):
            return v.signed() if signed else v.v
        return int(v)

    def run_block(self, inst, blk, pending):
        if blk[0] == "proc":
            inst.exec_seq(blk[1], blk[2], pending)
        elif blk[0] == "cassign":
            _, target, waves, scope = blk
            for val, cond in waves:
                if cond is None or inst.truth(inst.ev(cond, scope)):
                    inst.assign(target, inst.ev(val, scope), scope, pending)
                    break
        elif blk[0] == "drive":
            _, sig, actual, scope = blk
            pending.append((sig, [], inst.conform(inst.ev(actual, scope), sig.typ, "port drive")))
        elif blk[0] == "tap":
            _, sig, target, scope = blk
            inst.assign(target, sig.val, scope, pending)
        elif blk[0] == "prim":
            blk[1].clock(pending)

    def commit(self, pending):
        changed = False
        for sig, path, v in pending:
            new = store(sig.val, path, v)
            if not same(new, sig.val):
                sig.val = new
                changed = True
        return changed

    def settle(self):
        for _ in range(64):
            pending = []
            for inst in self.insts:
                for blk in inst.comb:
                    self.run_block(inst, blk, pending)
            if not self.commit(pending):
                return
        raise RuntimeError("combinational logic does not settle")

    def step(self, **inputs):
        if inputs:
            self.set(**inputs)
        self.settle()
        pending = []
        for inst in self.insts:
            for blk in inst.clocked:
                self.run_block(inst, blk, pending)
        self.commit(pending)
        self.settle()


# --------------------------------------------------------------------------------------------- primitives


class Dsp48:
    """Behavioural DSP48E1 / DSP48E2 (Xilinx UNISIM; absent from the reference tree) for the configuration
    src/tay1_order.vhd and src/mults/*.vhd instantiate: USE_MULT = MULTIPLY, direct A / B inputs, no pre-adder,
    AREG = BREG in {1, 2}, CREG = MREG = PREG = 1, registered OPMODE / ALUMODE, all clock enables '1', synchronous
    resets.  Per rising edge (all right-hand sides are the values before the edge):
        A2 <= A (AREG = 1) | A1 (AREG = 2);  A1 <= A;  likewise B;  C' <= C
        M  <= signed(A2[AW-1:0]) * signed(B2[17:0])            AW = 25 (E1) / 27 (E2)
        P  <= alu(X, Y, Z, W)                                    48 bits, wrapping
    X = OPMODE[1:0]: 00 -> 0, 01 -> M (with Y = 01), 10 -> P;   Y = OPMODE[3:2]: 00 -> 0, 01 -> M, 11 -> C
    Z = OPMODE[6:4]: 000 -> 0, 001 -> PCIN, 010 -> P, 011 -> C, 101 -> PCIN >> 17 (arithmetic), 110 -> P >> 17
    W = OPMODE[8:7] (E2 only): 00 -> 0, 01 -> P, 11 -> C
    ALUMODE 0000: Z + W + X + Y;  0011: Z - (W + X + Y);  0001: -Z + (W + X + Y) - 1;  0010: -(Z + W + X + Y) - 1
    PCOUT = P.  Anything else raises."""
    AW = 25
    OPW = 7
    DW = 25
    _IN = dict(a=30, b=18, c=48, pcin=48, acin=30, bcin=18, alumode=4, carryinsel=3, inmode=5,
               carryin=None, carrycascin=None, multsignin=None, clk=None)
    _CE = ("cea1", "cea2", "cead", "cealumode", "ceb1", "ceb2", "cec", "cecarryin", "cectrl", "ced", "ceinmode", "cem", "cep")
    _RST = ("rsta", "rstallcarryin", "rstalumode", "rstb", "rstc", "rstctrl", "rstd", "rstinmode", "rstm", "rstp")
    _OUT = dict(p=48, pcout=48, acout=30, bcout=18)
    _GEN_OK = dict(a_input="DIRECT", b_input="DIRECT", use_dport=False, use_mult="MULTIPLY", use_simd="ONE48",
                   amultsel="A", bmultsel="B", preaddinsel="A", creg=1, mreg=1, preg=1, opmodereg=1, alumodereg=1,
                   carryinreg=1, carryinselreg=1, inmodereg=1, acascreg=None, bcascreg=None, adreg=None, dreg=None)

    @classmethod
    def ports_table(cls):
        t = {n: ("in", w) for n, w in cls._IN.items()}
        t["d"] = ("in", cls.DW)
        t["opmode"] = ("in", cls.OPW)
        t.update({n: ("in", None) for n in cls._CE + cls._RST})
        t.update({n: ("out", w) for n, w in cls._OUT.items()})
        return t

    def __init__(self, parent, generics, port_sigs, path):
        self.path = path
        g = {k.lower(): v for k, v in generics.items()}
        for k, v in g.items():
            if k in ("areg", "breg"):
                if v not in (1, 2):
                    raise NotImplementedError(f"{path}: {k} = {v}")
            elif k not in self._GEN_OK:
                raise NotImplementedError(f"{path}: generic {k} is not modelled")
            elif self._GEN_OK[k] is not None and v != self._GEN_OK[k]:
                raise NotImplementedError(f"{path}: {k} = {v!r} is not modelled")
        self.areg, self.breg = g.get("areg", 1), g.get("breg", 1)
        self.ports = {}
        for name, (mode, w) in self.PORTS.items():
            typ = ("sl",) if w is None else ("slv", w - 1, 0)
            sig = port_sigs.get(name)
            if sig is None:
                sig = Sig(path + "." + name, typ, SL(1 if name in self._CE else 0) if w is None else BV(w, 0))
            else:
                have = sig.val.w if isinstance(sig.val, BV) else None
                if have != w:
                    raise TypeError(f"{path}: port {name} is {w} bits, actual {sig.name} is {have}")
            sig_typ = typ
            self.ports[name] = sig
            if sig.typ is None:
                sig.typ = sig_typ
        self.comb, self.clocked, self.children = [], [("prim", self)], []
        self.a1 = self.a2 = self.b1 = self.b2 = self.c = self.m = 0
        self.opmode = self.alumode = 0

    def all_instances(self):
        return [self]

    # the Simulator calls these two through run_block / conform on port drives
    def conform(self, val, typ, what="value"):
        raise NotImplementedError

    def _in(self, name, signed=True):
        v = self.ports[name].val
        if isinstance(v, BV):
            return v.signed() if signed else v.v
        return int(v)

    @staticmethod
    def _sx(v, w):
        v &= (1 << w) - 1
        return v - (1 << w) if v >> (w - 1) else v

    def clock(self, pending):
        for n in self._CE:
            if self._in(n) != 1:
                raise NotImplementedError(f"{self.path}: clock enable {n} is not '1'")
        for n in ("carryin", "carrycascin", "multsignin", "carryinsel", "inmode"):
            if self._in(n, False) != 0:
                raise NotImplementedError(f"{self.path}: {n} /= 0")
        rst = {n: self._in(n) for n in self._RST}
        p_old = self.ports["p"].val.signed()
        op, alu = self.opmode, self.alumode
        x, y, z, w = op & 3, (op >> 2) & 3, (op >> 4) & 7, (op >> 7) & 3
        if (x == 1) != (y == 1):
            raise NotImplementedError(f"{self.path}: OPMODE {op:b}: X and Y must select M together")
        pcin = self._in("pcin")
        xv = {0: 0, 1: self.m, 2: p_old}.get(x)
        yv = {0: 0, 1: 0, 3: self.c}.get(y)
        zv = {0: 0, 1: pcin, 2: p_old, 3: self.c, 5: pcin >> 17, 6: p_old >> 17}.get(z)
        wv = {0: 0, 1: p_old, 3: self.c}.get(w)
        if None in (xv, yv, zv, wv):
            raise NotImplementedError(f"{self.path}: OPMODE {op:b} selects an input that is not modelled")
        s = wv + xv + yv
        if alu == 0:
            p_new = zv + s
        elif alu == 3:
            p_new = zv - s
        elif alu == 1:
            p_new = -zv + s - 1
        elif alu == 2:
            p_new = -(zv + s) - 1
        else:
            raise NotImplementedError(f"{self.path}: ALUMODE {alu:b}")
        a_in, b_in = self._in("a", False), self._in("b", False)
        m_new = self._sx(self.a2, self.AW) * self._sx(self.b2, 18)
        a2_new = a_in if self.areg == 1 else self.a1
        b2_new = b_in if self.breg == 1 else self.b1
        self.a1, self.b1 = (0 if rst["rsta"] else a_in), (0 if rst["rstb"] else b_in)
        self.a2, self.b2 = (0 if rst["rsta"] else a2_new), (0 if rst["rstb"] else b2_new)
        self.c = 0 if rst["rstc"] else self._in("c")
        self.m = 0 if rst["rstm"] else m_new
        self.opmode = 0 if rst["rstctrl"] else self._in("opmode", False)
        self.alumode = 0 if rst["rstalumode"] else self._in("alumode", False)
        p_bv = BV(48, 0 if rst["rstp"] else p_new)
        pending.append((self.ports["p"], [], p_bv))
        pending.append((self.ports["pcout"], [], BV(48, p_bv.v)))


class Dsp48E1(Dsp48):
    AW, OPW, DW = 25, 7, 25


class Dsp48E2(Dsp48):
    AW, OPW, DW = 27, 9, 27


Dsp48E1.PORTS = Dsp48E1.ports_table()
Dsp48E2.PORTS = Dsp48E2.ports_table()
PRIMITIVES = {"dsp48e1": Dsp48E1, "dsp48e2": Dsp48E2}


# --------------------------------------------------------------------------------------------- drivers

REF_SRC = "/root/reference/src"
RTL_FILES = ("cordic_dds.vhd", "cordic_dds48.vhd", "cordic_dds_scaled.vhd", "cordic_atan2.vhd", "int_multNxN_dsp48.vhd",
             "hamming_win.vhd", "bh_win_3term.vhd", "bh_win_4term.vhd", "bh_win_5term.vhd", "bh_win_7term.vhd", "win_selector.vhd",
             "taylor_sincos.vhd", "tay1_order.vhd", "mults/mlt35x25_dsp48e1.vhd", "mults/mlt35x27_dsp48e2.vhd")


def reference_library(src=REF_SRC, files=RTL_FILES):
    """The reference's RTL, parsed where it lies."""
    return Library([os.path.join(src, f) for f in files])


def run_dds(lib, entity, phase_width, data_width, phases, precision=None):
    """DT_SIN / DT_COS of a DDS entity for a list of PH_IN values, with PH_EN held high.  The latency is found
    from DT_VAL (the entity's own valid flag), not assumed.  -> list of (sin, cos)"""
    gen = {"PHASE_WIDTH": phase_width, "DATA_WIDTH": data_width}
    if precision is not None:
        gen["PRECISION"] = precision
    sim = Simulator(lib, entity, gen)
    for _ in range(4):
        sim.step(reset=1, ph_en=0, ph_in=0)
    out, first_valid, t = [], None, 0
    total = len(phases)
    # feed the phases, then keep clocking with PH_EN low until all results have come out
    while len(out) < total:
        if t < total:
            sim.step(reset=0, ph_en=1, ph_in=phases[t])
        else:
            sim.step(reset=0, ph_en=0, ph_in=0)
        t += 1
        if sim.get("dt_val") == 1:
            if first_valid is None:
                first_valid = t
            out.append((sim.get("dt_sin"), sim.get("dt_cos")))
        if t > total + 4 * data_width + 64:
            raise RuntimeError("DT_VAL never covered all phases")
    return out, first_valid


def run_window(lib, entity, generics, aa, clocks):
    """Clock a window entity (or win_selector) with ENABLE high: -> list of (DT_WIN, DT_VLD) per clock after reset."""
    sim = Simulator(lib, entity, generics)
    ports = {f"aa{k}": v for k, v in enumerate(aa) if f"aa{k}" in sim.top.ports}
    for _ in range(4):
        sim.step(reset=1, enable=0, **ports)
    out = []
    for _ in range(clocks):
        sim.step(reset=0, enable=1, **ports)
        out.append((sim.get("dt_win"), sim.get("dt_vld")))
    return out


def run_atan2(lib, input_width, angle_width, precision, pairs):
    """Clock cordic_atan2 with one (VEC_DX, VEC_DY) pair per clock, then idle clocks: -> per-clock list of
    (PHI_DT, PHI_VL).  (PHI_VL is not aligned with PHI_DT in the entity - its shift register is two stages
    shorter than the data path - so the caller finds the data latency from the data.)"""
    sim = Simulator(lib, "cordic_atan2", {"INPUT_WIDTH": input_width, "ANGLE_WIDTH": angle_width, "PRECISION": precision})
    for _ in range(4):
        sim.step(reset=1, vec_en=0, vec_dx=0, vec_dy=0)
    out = []
    for t in range(len(pairs) + angle_width + 8):
        if t < len(pairs):
            sim.step(reset=0, vec_en=1, vec_dx=pairs[t][0], vec_dy=pairs[t][1])
        else:
            sim.step(reset=0, vec_en=0, vec_dx=0, vec_dy=0)
        out.append((sim.get("phi_dt"), sim.get("phi_vl")))
    return out


def run_taylor(lib, phase_width, data_width, lut_size, xseries, clocks, start=0):
    """Clock taylor_sincos with PHI_ENA high (its phase counter is internal): -> per-clock list of (OUT_SIN, OUT_COS).
    start: value deposited into the phase counter `cnt` after reset (a simulator `force -deposit`), so that a long
    period can be entered anywhere without clocking up to it."""
    sim = Simulator(lib, "taylor_sincos", {"PHASE_WIDTH": phase_width, "DATA_WIDTH": data_width, "LUT_SIZE": lut_size, "XSERIES": xseries})
    for _ in range(4):
        sim.step(rst=1, phi_ena=0)
    if start:
        cnt = sim.top.scope.get("cnt")
        cnt.val = BV(phase_width, start, cnt.val.left, cnt.val.right)
        sim.settle()
    out = []
    for _ in range(clocks):
        sim.step(rst=0, phi_ena=1)
        out.append((sim.get("out_sin"), sim.get("out_cos")))
    return out


def run_mult(lib, dtw, pairs):
    """Clock int_multNxN_dsp48 with one (DAT_A, DAT_B) per clock: -> per-clock DAT_Q (signed)."""
    sim = Simulator(lib, "int_multNxN_dsp48", {"DTW": dtw})
    for _ in range(2):
        sim.step(rst=1, dat_a=0, dat_b=0)
    out = []
    for t in range(len(pairs) + 4):
        a, b = pairs[t] if t < len(pairs) else (0, 0)
        sim.step(rst=0, dat_a=a, dat_b=b)
        out.append(sim.get("dat_q"))
    return out
