#!/usr/bin/env python3
"""cs2cpp -- MECHANICAL transliteration of the reference's C# battle code into C++ (test infrastructure).

Why: the reference's engine (Unity C#) cannot run offline and ships no tests, so the hand-written CPU oracle
(oracle/footsies_oracle.c) had nothing reference-held to be pinned against.  This tool turns the reference's OWN source
text into a second, independent checker: it reads /root/reference/Assets/Script/*.cs where they lie, rewrites SYNTAX only
(token by token, statement order and line structure preserved so that the output can be read side by side with the .cs)
and writes oracle/_ref/footsies_ref_gen.cpp, which is compiled against oracle/ref_shim/ (UnityEngine stand-ins + a harness
with the oracle's C API).  The output is generated, git-ignored and never committed: no reference source enters the repo.

Rewrite rules (everything else is copied verbatim):
  R1  reference types are held as Ref<T>; `new T(a)` -> New<T>(a); `new T[n]` -> NewArray<T>(n); `null` -> nullptr
  R2  every member access `.` -> `->` (value types in the shim define operator-> returning this);
      `Type.member` -> `Type::member` when the left side is a type name (statics, enums, nested types)
  R3  computed properties `T p { get { body } }` -> method `T p()`, uses get `()`; auto-properties -> plain fields
  R4  `var` -> auto; `foreach (T x in e)` -> `for (T x : e)`; `(x) => e` -> `[&](auto x) { return e; }`
  R5  fields without initializer get `{}` (C# zero-initialises); float literals `4f` -> `4.0f`; uint -> uint32_t ...
  R6  `x is T` -> Is<T>(x); statement `a?.M(..)` -> `if ((a) != nullptr) (a)->M(..)`
  R7  object initialisers `new T { a = 1, b = 2 }` -> `([&]{ auto _o = New<T>(); _o->a = 1; _o->b = 2; return _o; }())`
  R8  `case X:` bodies are wrapped in braces (C++ forbids jumping over initialisations)
  R9  statements starting with `Debug.` are dropped (string concatenation with ints has no C++ spelling; logging only)
  R10 methods are declared in the class and defined after all classes (C# has no declaration order)
Per-file exceptions are listed in FILES below (methods skipped or replaced by hand, each with the reason).
"""
import argparse
import os
import re
import sys

REF_SCRIPTS = "/root/reference/Assets/Script"

# file -> options.  skip: members not emitted (reason given); replace: member body replaced by hand-written C++;
# drop_lines: regexes of source lines removed before tokenising.
FILES = [
    ("InputData.cs", dict(replace={
        # MemberwiseClone is a runtime service (shallow field copy of the dynamic type); the copy constructor is its C++ spelling
        "InputData.ShallowCopy": "{ return New<InputData>(*this); }"})),
    ("AttackData.cs", {}),
    ("ActionData.cs", {}),
    ("ActionDataContainer.cs", {}),
    ("AttackDataContainer.cs", {}),
    ("MotionDataContainer.cs", {}),
    ("FighterData.cs", {}),
    ("FighterState.cs", {}),
    ("Fighter.cs", dict(skip={"Fighter.GetCurrentMotionSprite": "sprites are presentation"})),
    ("BattleState.cs", {}),
    ("EnvironmentState.cs", {}),
    ("TrainingActor.cs", {}),
    ("TrainingBattleAIActor.cs", {}),
    ("TrainingManager.cs", dict(skip={"TrainingManager.Setup": "async socket accept (boundary ii, deleted)"})),
    ("BattleAI.cs", {}),
    ("BattleCore.cs", dict(drop_lines=[
        # waits for the TCP peers to connect (boundary ii, deleted); everything else in Start() is kept
        r"Task\.WhenAll\(new Task\[\] \{trainingManager\.Setup\(\), trainingRemoteControl\.Setup\(\)\}\)\.Wait\(\);"])),
]

PRIMS = {"int": "int", "uint": "uint32_t", "float": "float", "double": "double", "bool": "bool", "string": "std::string",
         "void": "void", "byte": "uint8_t", "long": "long", "object": "Ref<Object>"}
SHIM_VALUE = {"Vector2", "Vector2Int", "Rect", "Task"}
SHIM_REF = {"List", "Queue", "Dictionary", "AudioClip", "Sprite", "GameObject", "Animator", "ScriptableObject",
            "MonoBehaviour", "GameManager", "SoundManager", "InputManager", "TrainingRemoteControl", "Object"}
SHIM_STATIC = {"Mathf", "Random", "Time", "Debug", "Array", "Application"}
SHIM_NESTED = {"TrainingRemoteControl": {"Command": "enum"}}
SHIM_PROPS = {"Count", "Length", "xMin", "xMax", "yMin", "yMax"}
MODIFIERS = {"public", "private", "protected", "internal", "static", "readonly", "const", "abstract", "virtual",
             "override", "async", "sealed", "new"}

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<comment>//[^\n]*|/\*.*?\*/)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<chr>'(?:\\.|[^'\\])')
  | (?P<num>\d+\.\d+[fFdD]?|\d+[fFdDuUlL]*)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>=>|\?\.|\?\?|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|\|=|&=|\^=|[{}()\[\];,.<>=+\-*/%!&|^~?:])
""", re.S | re.X)


class Tok:
    __slots__ = ("kind", "text", "line")

    def __init__(self, kind, text, line):
        self.kind, self.text, self.line = kind, text, line

    def __repr__(self):
        return f"{self.kind}:{self.text!r}@{self.line}"


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenise at line {line}: {src[pos:pos + 30]!r}")
        toks.append(Tok(m.lastgroup, m.group(), line))
        line += m.group().count("\n")
        pos = m.end()
    return toks


class CsType:
    def __init__(self, name, args=(), array=False):
        self.name, self.args, self.array = name, tuple(args), array

    def elem(self):
        if self.array:
            return CsType(self.name, self.args, False)
        if self.name in ("List", "Queue") and self.args:
            return self.args[0]
        if self.name == "Dictionary" and len(self.args) == 2:
            return self.args[1]
        return None

    def __repr__(self):
        return self.name + ("<" + ",".join(map(repr, self.args)) + ">" if self.args else "") + ("[]" if self.array else "")


class ClassInfo:
    def __init__(self, name, kind, outer=None):
        self.name, self.kind, self.outer = name, kind, outer      # kind: class | interface | enum
        self.bases, self.members, self.fields, self.props, self.nested = [], [], {}, set(), {}
        self.file = self.line = None

    def qual(self):
        return (self.outer.qual() + "::" if self.outer else "") + self.name


class Translator:
    def __init__(self):
        self.classes = {}            # simple name -> ClassInfo (top level and nested)
        self.order = []              # top-level classes in input order
        self.all_props = set(SHIM_PROPS)

    # ------------------------------------------------------------------ type knowledge
    def kind_of(self, name, ctx=None):
        if name in PRIMS or name == "var":
            return "prim"
        if name in SHIM_VALUE:
            return "value"
        if name in SHIM_REF:
            return "ref"
        if name in SHIM_STATIC:
            return "static"
        if name in self.classes:
            return "enum" if self.classes[name].kind == "enum" else "ref"
        for outer, inner in SHIM_NESTED.items():
            if name in inner:
                return inner[name]
        if name == "System":
            return "ns"
        return None

    def cpp_name(self, name, cls):
        """C++ spelling of a simple type name as seen from class `cls` (nested types need qualification elsewhere)."""
        info = self.classes.get(name)
        if info is not None and info.outer is not None:
            c = cls
            while c is not None:
                if c is info.outer or c is info:
                    return name
                c = c.outer
            return info.qual()
        return name

    def cpp_type(self, t, cls, wrap=True):
        if t.name == "System.Action":
            return "std::function<void(" + ", ".join(self.cpp_type(a, cls) for a in t.args) + ")>"
        if t.array:
            return "Ref<CsArray<" + self.cpp_type(t.elem(), cls) + ">>"
        if t.name == "var":
            return "auto"
        if t.name in PRIMS:
            return PRIMS[t.name]
        parts = t.name.split(".")
        base = "::".join([self.cpp_name(parts[0], cls)] + parts[1:])
        if t.args:
            base += "<" + ", ".join(self.cpp_type(a, cls) for a in t.args) + ">"
        k = self.kind_of(parts[-1])
        return "Ref<" + base + ">" if (k == "ref" and wrap) else base

    # ------------------------------------------------------------------ token helpers
    @staticmethod
    def sig(toks, i):
        """index of the next significant token at or after i"""
        while i < len(toks) and toks[i].kind in ("ws", "comment"):
            i += 1
        return i

    @staticmethod
    def match_close(toks, i):
        """toks[i] is an opening bracket; index of its partner"""
        pairs = {"(": ")", "{": "}", "[": "]"}
        o, c, depth = toks[i].text, pairs[toks[i].text], 0
        while i < len(toks):
            if toks[i].kind == "op":
                if toks[i].text == o:
                    depth += 1
                elif toks[i].text == c:
                    depth -= 1
                    if depth == 0:
                        return i
            i += 1
        raise SyntaxError("unbalanced bracket")

    def try_type(self, toks, i):
        """Parse a C# type starting at significant index i -> (CsType, next index) or None."""
        i = self.sig(toks, i)
        if i >= len(toks) or toks[i].kind != "id":
            return None
        name = toks[i].text
        if self.kind_of(name) is None or self.kind_of(name) == "static":
            return None
        j = i + 1
        # dotted: System.Action / Outer.Nested
        while True:
            k = self.sig(toks, j)
            if k + 1 < len(toks) and toks[k].text == "." and toks[self.sig(toks, k + 1)].kind == "id":
                nxt = toks[self.sig(toks, k + 1)].text
                if name == "System" or self.kind_of(nxt) in ("ref", "enum") and (
                        nxt in SHIM_NESTED.get(name.split(".")[-1], {}) or
                        (nxt in self.classes and self.classes[nxt].outer is not None
                         and self.classes[nxt].outer.name == name.split(".")[-1])):
                    name += "." + nxt
                    j = self.sig(toks, k + 1) + 1
                    continue
            break
        if name == "System":
            return None
        args = []
        k = self.sig(toks, j)
        if k < len(toks) and toks[k].text == "<" and (name in ("List", "Queue", "Dictionary", "System.Action", "Task")):
            k += 1
            while True:
                r = self.try_type(toks, k)
                if r is None:
                    return None
                args.append(r[0])
                k = self.sig(toks, r[1])
                if toks[k].text == ",":
                    k += 1
                    continue
                if toks[k].text == ">":
                    j = k + 1
                    break
                return None
        array = False
        k = self.sig(toks, j)
        if k + 1 < len(toks) and toks[k].text == "[" and toks[self.sig(toks, k + 1)].text == "]":
            array = True
            j = self.sig(toks, k + 1) + 1
        return CsType(name, args, array), j

    # ------------------------------------------------------------------ pass 1: structure
    def parse_file(self, fname, opts):
        path = os.path.join(REF_SCRIPTS, fname)
        src = open(path, encoding="utf-8-sig").read()
        for pat in opts.get("drop_lines", []):
            src, n = re.subn(r"[^\n]*" + pat + r"[^\n]*", "/* dropped by cs2cpp (FILES.drop_lines) */", src)
            if n != 1:
                raise SystemExit(f"{fname}: drop_lines pattern matched {n} times: {pat}")
        toks = tokenize(src)
        i = 0
        while i < len(toks):
            t = toks[i]
            if t.kind == "id" and t.text == "using":
                while toks[i].text != ";":
                    i += 1
            elif t.kind == "id" and t.text == "namespace":
                i = self.sig(toks, i + 1) + 1                     # name
                i = self.sig(toks, i)                             # {
                end = self.match_close(toks, i)
                self.parse_types(toks, i + 1, end, None, fname, opts)
                i = end
            i += 1

    def skip_attributes(self, toks, i):
        i = self.sig(toks, i)
        while i < len(toks) and toks[i].text == "[":
            i = self.sig(toks, self.match_close(toks, i) + 1)
        return i

    def parse_types(self, toks, lo, hi, outer, fname, opts):
        i = lo
        while True:
            i = self.skip_attributes(toks, i)
            if i >= hi:
                return
            while toks[i].kind == "id" and toks[i].text in MODIFIERS:
                i = self.sig(toks, i + 1)
            kw = toks[i].text
            if kw not in ("class", "enum", "interface"):
                raise SyntaxError(f"{fname}:{toks[i].line}: expected a type declaration, got {kw}")
            i = self.parse_type_decl(toks, i, outer, fname, opts) + 1

    def parse_type_decl(self, toks, i, outer, fname, opts):
        kw = toks[i].text
        i = self.sig(toks, i + 1)
        info = ClassInfo(toks[i].text, kw, outer)
        info.file, info.line, info.opts = fname, toks[i].line, opts
        self.classes[info.name] = info
        if outer is None:
            self.order.append(info)
        else:
            outer.nested[info.name] = info
            outer.members.append(("nested", info))
        i = self.sig(toks, i + 1)
        if toks[i].text == ":":
            i = self.sig(toks, i + 1)
            while toks[i].text != "{":
                if toks[i].kind == "id":
                    info.bases.append(toks[i].text)
                i = self.sig(toks, i + 1)
        end = self.match_close(toks, i)
        if kw == "enum":
            info.enum_body = (toks, i + 1, end)
        else:
            info.pending = (toks, i + 1, end)
        return end

    def parse_members(self, info):
        toks, i, hi = info.pending
        fname = info.file
        while True:
            i = self.skip_attributes(toks, i)
            if i >= hi:
                return
            first_line = toks[i].line
            mods = []
            while toks[i].kind == "id" and toks[i].text in MODIFIERS:
                # `new()` never starts a member, so `new` here is the hiding modifier
                mods.append(toks[i].text)
                i = self.sig(toks, i + 1)
            if toks[i].text in ("class", "enum", "interface"):
                i = self.sig(toks, self.parse_type_decl(toks, i, info, fname, info.opts) + 1)
                nested = info.members[-1][1]
                if nested.kind != "enum":
                    self.parse_members(nested)
                continue
            # constructor?
            if toks[i].kind == "id" and toks[i].text == info.name and toks[self.sig(toks, i + 1)].text == "(":
                p0 = self.sig(toks, i + 1)
                p1 = self.match_close(toks, p0)
                b0 = self.sig(toks, p1 + 1)
                b1 = self.match_close(toks, b0)
                info.members.append(("ctor", dict(name=info.name, params=(p0, p1), body=(b0, b1), line=first_line, mods=mods)))
                i = b1 + 1
                continue
            r = self.try_type(toks, i)
            if r is None:
                raise SyntaxError(f"{fname}:{toks[i].line}: cannot parse member type at {toks[i].text!r}")
            typ, j = r
            j = self.sig(toks, j)
            name = toks[j].text
            k = self.sig(toks, j + 1)
            if toks[k].text == "(":                               # method
                p1 = self.match_close(toks, k)
                b0 = self.sig(toks, p1 + 1)
                if toks[b0].text == ";":                          # interface method
                    info.members.append(("method", dict(name=name, type=typ, params=(k, p1), body=None, line=first_line, mods=mods)))
                    i = b0 + 1
                else:
                    b1 = self.match_close(toks, b0)
                    info.members.append(("method", dict(name=name, type=typ, params=(k, p1), body=(b0, b1), line=first_line, mods=mods)))
                    i = b1 + 1
            elif toks[k].text == "{":                             # property
                e = self.match_close(toks, k)
                inner = [t for t in toks[k + 1:e] if t.kind not in ("ws", "comment")]
                g = self.sig(toks, k + 1)
                if toks[g].text == "get" and toks[self.sig(toks, g + 1)].text == "{":
                    b0 = self.sig(toks, g + 1)
                    b1 = self.match_close(toks, b0)
                    if self.sig(toks, b1 + 1) != e:
                        raise SyntaxError(f"{fname}:{toks[k].line}: property {name} has more than a getter body")
                    info.members.append(("prop", dict(name=name, type=typ, body=(b0, b1), line=first_line, mods=mods)))
                    info.props.add(name)
                    self.all_props.add(name)
                    i = e + 1
                else:
                    if any(t.text == "{" for t in inner):
                        raise SyntaxError(f"{fname}:{toks[k].line}: unsupported property form for {name}")
                    init = None
                    i = self.sig(toks, e + 1)
                    if toks[i].text == "=":                       # auto-property initialiser
                        s = i + 1
                        while toks[i].text != ";":
                            i += 1
                        init = (s, i)
                        i += 1
                    info.members.append(("field", dict(name=name, type=typ, init=init, line=first_line, mods=mods)))
                    info.fields[name] = typ
            else:                                                 # field(s)
                init = None
                if toks[k].text == "=":
                    s = k + 1
                    depth = 0
                    while not (toks[k].text == ";" and depth == 0):
                        if toks[k].text in "({[":
                            depth += 1
                        elif toks[k].text in ")}]":
                            depth -= 1
                        k += 1
                    init = (s, k)
                if toks[k].text != ";":
                    raise SyntaxError(f"{fname}:{toks[k].line}: unsupported field declaration {name}")
                info.members.append(("field", dict(name=name, type=typ, init=init, line=first_line, mods=mods)))
                info.fields[name] = typ
                i = k + 1

    # ------------------------------------------------------------------ pass 2: bodies
    def params_cpp(self, toks, p0, p1, cls, with_defaults, locals_):
        out, i = [], self.sig(toks, p0 + 1)
        while i < p1:
            typ, j = self.try_type(toks, i)
            j = self.sig(toks, j)
            name = toks[j].text
            locals_[name] = typ
            j = self.sig(toks, j + 1)
            default = ""
            if toks[j].text == "=":
                s = j + 1
                while j < p1 and toks[j].text != ",":
                    j += 1
                default = " = " + self.body(toks, s, j, cls, {}, None).strip()
            out.append(self.cpp_type(typ, cls) + " " + name + (default if with_defaults else ""))
            if toks[j].text == ",":
                j += 1
            i = self.sig(toks, j)
        return ", ".join(out)

    def body(self, toks, lo, hi, cls, locals_, ret_type, expected=None):
        """Translate toks[lo:hi] (statements or an expression)."""
        out = []                      # (text, paren_depth) ; depth bookkeeping serves the look-back rules R6
        depth = 0
        lambda_stack = []             # paren depth at which an expression lambda must be closed
        last_decl = None              # (name, CsType) of the declaration being initialised
        i = lo

        def emit(s):
            out.append((s, depth))

        def prev_sig_out():
            for s, _ in reversed(out):
                if s.strip() and not s.lstrip().startswith("//") and not s.lstrip().startswith("/*"):
                    return s
            return ""

        def at_statement_start():
            p = prev_sig_out()
            return p in ("", "{", "}", ";") or p.endswith(":") and not p.endswith("::")

        def lookup(name):
            if name in locals_:
                return locals_[name]
            c = cls
            while c is not None:
                if name in c.fields:
                    return c.fields[name]
                c = c.outer
            return None

        def target_type_before(idx):
            """type expected by a target-typed `new` whose `new` token sits at idx (R1)"""
            if expected is not None and self.sig(toks, lo) == idx:
                return expected
            k = idx - 1
            while toks[k].kind in ("ws", "comment"):
                k -= 1
            if toks[k].text == "return":
                return ret_type
            if toks[k].text != "=":
                return expected
            k -= 1
            while toks[k].kind in ("ws", "comment"):
                k -= 1
            if toks[k].text == "]":
                d = 0
                while True:
                    if toks[k].text == "]":
                        d += 1
                    elif toks[k].text == "[":
                        d -= 1
                        if d == 0:
                            break
                    k -= 1
                k -= 1
                while toks[k].kind in ("ws", "comment"):
                    k -= 1
                t = lookup(toks[k].text)
                return t.elem() if t else None
            if toks[k].kind == "id":
                if last_decl and last_decl[0] == toks[k].text:
                    return last_decl[1]
                return lookup(toks[k].text)
            return None

        def emit_initializer(ctor_cpp, typ, b0):
            """R7: toks[b0] == '{' of an object initialiser"""
            b1 = self.match_close(toks, b0)
            info = self.classes.get(typ.name.split(".")[-1]) if typ else None
            parts, s, d, k = [], b0 + 1, 0, b0 + 1
            while k < b1:
                tx = toks[k].text
                if tx in "({[":
                    d += 1
                elif tx in ")}]":
                    d -= 1
                elif tx == "," and d == 0:
                    parts.append((s, k))
                    s = k + 1
                k += 1
            if any(t.kind not in ("ws", "comment") for t in toks[s:b1]):
                parts.append((s, b1))
            emit("([&]{ auto _o = " + ctor_cpp + ";")
            for (s, e) in parts:
                n = self.sig(toks, s)
                eq = self.sig(toks, n + 1)
                assert toks[eq].text == "=", f"{cls.file}:{toks[n].line}: initialiser entry"
                ftype = info.fields.get(toks[n].text) if info else None
                lead = "".join(t.text for t in toks[s:n])
                emit(lead + "_o->" + toks[n].text + " =" + self.body(toks, eq + 1, e, cls, locals_, ret_type, expected=ftype) + ";")
            emit(" return _o; }())")
            return b1 + 1

        while i < hi:
            t = toks[i]
            if t.kind in ("ws", "comment", "str", "chr"):
                emit(t.text)
                i += 1
                continue
            if t.kind == "num":
                s = t.text
                if re.fullmatch(r"\d+[fF]", s):
                    s = s[:-1] + ".0f"
                elif re.fullmatch(r"\d+[uU]", s):
                    s = s
                emit(s)
                i += 1
                continue
            if t.kind == "op":
                tx = t.text
                if tx in ")" and lambda_stack and lambda_stack[-1] == depth:
                    lambda_stack.pop()
                    emit("; }")
                if tx == "," and lambda_stack and lambda_stack[-1] == depth:
                    lambda_stack.pop()
                    emit("; }")
                if tx in "([":
                    # lambda `(x) => ...`  (R4)
                    if tx == "(":
                        a = self.sig(toks, i + 1)
                        b = self.sig(toks, a + 1)
                        c = self.sig(toks, b + 1)
                        if toks[a].kind == "id" and toks[b].text == ")" and toks[c].text == "=>":
                            locals_[toks[a].text] = None
                            n = self.sig(toks, c + 1)
                            if toks[n].text == "{":
                                emit("[&](auto " + toks[a].text + ")")
                                i = c + 1
                            else:
                                emit("[&](auto " + toks[a].text + ") { return")
                                lambda_stack.append(depth)
                                i = c + 1
                            continue
                        # cast `(Type)expr`
                        r = self.try_type(toks, i + 1)
                        if r is not None and toks[self.sig(toks, r[1])].text == ")" and r[0].name != "var":
                            after = toks[self.sig(toks, self.sig(toks, r[1]) + 1)]
                            if after.kind in ("id", "num") or after.text == "(":
                                emit("(" + self.cpp_type(r[0], cls) + ")")
                                i = self.sig(toks, r[1]) + 1
                                continue
                    emit(tx)
                    depth += 1
                    i += 1
                    continue
                if tx in ")]":
                    depth -= 1
                    emit(tx)
                    i += 1
                    continue
                if tx == ".":
                    emit("->")
                    i += 1
                    continue
                if tx == "?.":                                     # R6, statement level only
                    k = len(out)
                    while k > 0 and not (out[k - 1][0] in ("{", "}", ";") and out[k - 1][1] == depth):
                        k -= 1
                    while out[k][0].strip() == "" or out[k][0].lstrip().startswith("//"):
                        k += 1                                      # leading whitespace / comments stay where they are
                    recv = "".join(s for s, _ in out[k:]).strip()
                    del out[k:]
                    emit("if ((" + recv + ") != nullptr) (" + recv + ")->")
                    i += 1
                    continue
                if tx in "{}":
                    if tx == "{":
                        depth += 1
                        emit(tx)                                    # recorded at the depth of the statements it opens
                    else:
                        depth -= 1
                        emit(tx)
                    i += 1
                    continue
                if tx == ";":
                    last_decl = None
                emit(tx)
                i += 1
                continue
            # ---- identifiers
            name = t.text
            prev = prev_sig_out()
            after_member = prev in ("->", "::")
            if not after_member:
                if name == "null":
                    emit("nullptr"); i += 1; continue
                if name == "this":
                    emit("this"); i += 1; continue
                if name == "Debug" and at_statement_start():       # R9
                    k = i
                    while toks[k].text != "(":
                        k += 1
                    k = self.match_close(toks, k)
                    k = self.sig(toks, k + 1)
                    assert toks[k].text == ";"
                    emit("/* Debug.Log dropped */;")
                    i = k + 1
                    continue
                if name == "foreach":                              # R4
                    p0 = self.sig(toks, i + 1)
                    p1 = self.match_close(toks, p0)
                    typ, j = self.try_type(toks, p0 + 1)
                    j = self.sig(toks, j)
                    var = toks[j].text
                    j = self.sig(toks, j + 1)
                    assert toks[j].text == "in"
                    locals_[var] = typ if typ.name != "var" else None
                    emit("for (" + self.cpp_type(typ, cls) + " " + var + " :" + self.body(toks, j + 1, p1, cls, locals_, ret_type) + ")")
                    i = p1 + 1
                    continue
                if name == "switch":                               # R8
                    p0 = self.sig(toks, i + 1)
                    p1 = self.match_close(toks, p0)
                    b0 = self.sig(toks, p1 + 1)
                    b1 = self.match_close(toks, b0)
                    emit("switch (" + self.body(toks, p0 + 1, p1, cls, locals_, ret_type) + ")")
                    emit("".join(x.text for x in toks[p1 + 1:b0]) + "{")
                    # split the body at top-level labels
                    k, d, labels = b0 + 1, 0, []
                    while k < b1:
                        tx = toks[k].text
                        if toks[k].kind == "op" and tx in "({[":
                            d += 1
                        elif toks[k].kind == "op" and tx in ")}]":
                            d -= 1
                        elif d == 0 and toks[k].kind == "id" and tx in ("case", "default"):
                            c = k
                            while toks[c].text != ":":
                                c += 1
                            labels.append((k, c))
                            k = c
                        k += 1
                    pos = b0 + 1
                    for n, (l0, l1) in enumerate(labels):
                        emit("".join(x.text for x in toks[pos:l0]))
                        emit(self.body(toks, l0, l1, cls, locals_, ret_type).replace("->", "::") + ":")
                        nxt = labels[n + 1][0] if n + 1 < len(labels) else b1
                        if all(x.kind in ("ws", "comment") for x in toks[l1 + 1:nxt]):
                            emit("".join(x.text for x in toks[l1 + 1:nxt]))
                        else:
                            # keep the trailing whitespace (indentation of the next label) outside the brace
                            e = nxt
                            while toks[e - 1].kind == "ws":
                                e -= 1
                            emit(" {" + self.body(toks, l1 + 1, e, cls, locals_, ret_type) + " }")
                            emit("".join(x.text for x in toks[e:nxt]))
                        pos = nxt
                    emit("}")
                    i = b1 + 1
                    continue
                if name == "new":                                  # R1 / R7
                    n = self.sig(toks, i + 1)
                    if toks[n].text == "(":                        # target-typed
                        typ = target_type_before(i)
                        if typ is None:
                            raise SyntaxError(f"{cls.file}:{t.line}: cannot infer the type of target-typed new")
                        p1 = self.match_close(toks, n)
                        ctor = "New<" + self.cpp_type(typ, cls, wrap=False) + ">(" + self.body(toks, n + 1, p1, cls, locals_, ret_type) + ")"
                        nb = self.sig(toks, p1 + 1)
                        if toks[nb].text == "{":
                            emit("".join(x.text for x in toks[p1 + 1:nb]))
                            i = emit_initializer(ctor, typ, nb)
                        else:
                            emit(ctor)
                            i = p1 + 1
                        continue
                    typ, j = self.try_type(toks, n)
                    j = self.sig(toks, j)
                    if toks[j].text == "[" and not typ.array:      # new T[n] (optionally { a, b })
                        e = self.match_close(toks, j)
                        size = self.body(toks, j + 1, e, cls, locals_, ret_type)
                        nb = self.sig(toks, e + 1)
                        if toks[nb].text == "{":
                            e2 = self.match_close(toks, nb)
                            emit("NewArray<" + self.cpp_type(typ, cls) + ">(" + size + ", {" + self.body(toks, nb + 1, e2, cls, locals_, ret_type) + "})")
                            i = e2 + 1
                        else:
                            emit("NewArray<" + self.cpp_type(typ, cls) + ">(" + size + ")")
                            i = e + 1
                        continue
                    is_value = self.kind_of(typ.name) == "value"
                    if toks[j].text == "(":
                        p1 = self.match_close(toks, j)
                        args = self.body(toks, j + 1, p1, cls, locals_, ret_type)
                        ctor = (self.cpp_type(typ, cls) + "(" + args + ")") if is_value else \
                            ("New<" + self.cpp_type(typ, cls, wrap=False) + ">(" + args + ")")
                        nb = self.sig(toks, p1 + 1)
                        if toks[nb].text == "{":
                            i = emit_initializer(ctor, typ, nb)
                        else:
                            emit(ctor)
                            i = p1 + 1
                        continue
                    if toks[j].text == "{":
                        ctor = (self.cpp_type(typ, cls) + "()") if is_value else ("New<" + self.cpp_type(typ, cls, wrap=False) + ">()")
                        emit("".join(x.text for x in toks[i + 1:n]).replace(" ", "", 1))
                        i = emit_initializer(ctor, typ, j)
                        continue
                    raise SyntaxError(f"{cls.file}:{t.line}: unsupported new-expression")
                if name == "is":                                   # R6
                    r = self.try_type(toks, i + 1)
                    k = len(out)
                    stop = {"(", "&&", "||", ",", "=", "return", "!"}
                    while k > 0 and not (out[k - 1][0] in stop and out[k - 1][1] <= depth):
                        k -= 1
                    operand = "".join(s for s, _ in out[k:]).strip()
                    del out[k:]
                    emit("Is<" + self.cpp_type(r[0], cls, wrap=False) + ">(" + operand + ")")
                    i = r[1]
                    continue
                r = self.try_type(toks, i)
                if r is not None and name not in locals_:
                    typ, j = r
                    n = self.sig(toks, j)
                    if toks[n].text == "." and not typ.args and not typ.array:      # R2 static / enum access
                        emit(self.cpp_type(typ, cls, wrap=False) + "::")
                        i = n + 1
                        continue
                    if toks[n].kind == "id" and toks[n].text not in ("in", "is"):   # local declaration
                        locals_[toks[n].text] = None if typ.name == "var" else typ
                        last_decl = (toks[n].text, None if typ.name == "var" else typ)
                        emit(self.cpp_type(typ, cls))
                        i = j
                        continue
                    emit(self.cpp_type(typ, cls))                   # generic argument, etc.
                    i = j
                    continue
                if self.kind_of(name) == "static" and toks[self.sig(toks, i + 1)].text == ".":
                    emit(name + "::")
                    i = self.sig(toks, i + 1) + 1
                    continue
            # plain identifier / member name: computed properties get `()`  (R3)
            nxt = toks[self.sig(toks, i + 1)].text if self.sig(toks, i + 1) < len(toks) else ""
            is_prop = False
            if nxt != "(":
                if prev == "->" and name in self.all_props:
                    is_prop = True
                    if name in self.ambiguous:                      # a field of one class, a property of another
                        recv = [x for x, _ in out if x.strip()][-2]
                        rt = lookup(recv)
                        rc = self.classes.get(rt.name) if rt is not None else None
                        if rc is None:
                            raise SyntaxError(f"{cls.file}:{t.line}: cannot resolve {recv}.{name} (field or property?)")
                        is_prop = name in rc.props
                elif prev == "::" and name in self.all_props and False:
                    is_prop = True
                elif not after_member and name not in locals_:
                    c = cls
                    while c is not None and not is_prop:
                        is_prop = name in c.props or any(
                            name in self.classes[b].props for b in c.bases if b in self.classes)
                        c = c.outer
            if prev == "::" and self.kind_of(name) in ("ref", "enum") and nxt == ".":
                emit(name + "::")                                   # Outer.Nested.Member
                i = self.sig(toks, i + 1) + 1
                continue
            emit(name + ("()" if is_prop else ""))
            i += 1
        while lambda_stack:
            lambda_stack.pop()
            emit("; }")
        return "".join(s for s, _ in out)

    # ------------------------------------------------------------------ output
    def sorted_classes(self):
        done, out = set(), []

        def visit(c):
            if c.name in done:
                return
            done.add(c.name)
            for b in c.bases:
                if b in self.classes and self.classes[b].outer is None:
                    visit(self.classes[b])
            out.append(c)
        for c in self.order:
            visit(c)
        return out

    def emit_enum(self, info, indent):
        toks, lo, hi = info.enum_body
        return f"{indent}enum class {info.name} : int {{" + "".join(t.text for t in toks[lo:hi]) + "};\n"

    def emit_class(self, info, indent, defs):
        opts = info.opts
        bases = [b for b in info.bases]
        if not bases:
            bases = ["Object"]
        lines = [f"{indent}struct {info.name} : " + ", ".join(bases) + f" {{   // {info.file}:{info.line}\n"]
        ind = indent + "    "
        toks = info.pending[0]
        for kind, m in info.members:
            if kind == "nested":
                if m.kind == "enum":
                    lines.append(self.emit_enum(m, ind))
                else:
                    lines.append(self.emit_class(m, ind, defs))
                continue
            full = f"{info.name}.{m['name']}"
            if full in opts.get("skip", {}):
                lines.append(f"{ind}// {m['name']} not transliterated: {opts['skip'][full]}\n")
                continue
            static = "static " if "static" in m["mods"] else ""
            if kind == "field":
                ctype = self.cpp_type(m["type"], info)
                init = "{}"
                if m["init"] is not None:
                    init = " =" + self.body(toks, m["init"][0], m["init"][1], info, {}, None, expected=m["type"])
                if static:
                    const = "const " if ("readonly" in m["mods"] or "const" in m["mods"]) else ""
                    lines.append(f"{ind}static inline {const}{ctype} {m['name']}{init};\n")
                else:
                    lines.append(f"{ind}{ctype} {m['name']}{init};\n")
                continue
            locals_ = {}
            qual = info.qual()
            if kind == "prop":
                rtype = self.cpp_type(m["type"], info)
                lines.append(f"{ind}{static}{rtype} {m['name']}();   // property\n")
                body = self.body(toks, m["body"][0], m["body"][1] + 1, info, locals_, m["type"])
                defs.append(f"// {info.file}:{m['line']}\nauto {qual}::{m['name']}() -> {rtype} {body}\n")
                continue
            if kind == "ctor":
                decl = self.params_cpp(toks, *m["params"], info, True, {})
                pdef = self.params_cpp(toks, *m["params"], info, False, locals_)
                lines.append(f"{ind}{m['name']}({decl});\n")
                body = self.body(toks, m["body"][0], m["body"][1] + 1, info, locals_, None)
                defs.append(f"// {info.file}:{m['line']}\n{qual}::{m['name']}({pdef}) {body}\n")
                continue
            rtype = self.cpp_type(m["type"], info)
            decl = self.params_cpp(toks, *m["params"], info, True, {})
            if m["body"] is None:                                   # interface method
                lines.append(f"{ind}virtual {rtype} {m['name']}({decl}) = 0;\n")
                continue
            virt = "virtual " if (not static and any(
                b in self.classes and self.classes[b].kind == "interface" for b in info.bases)) else ""
            lines.append(f"{ind}{static}{virt}{rtype} {m['name']}({decl});\n")
            pdef = self.params_cpp(toks, *m["params"], info, False, locals_)
            if full in opts.get("replace", {}):
                body = opts["replace"][full] + "   // hand-written replacement, see tools/cs2cpp.py FILES"
            else:
                body = self.body(toks, m["body"][0], m["body"][1] + 1, info, locals_, m["type"])
            defs.append(f"// {info.file}:{m['line']}\nauto {qual}::{m['name']}({pdef}) -> {rtype} {body}\n")
        lines.append(f"{indent}}};\n")
        return "".join(lines)

    def run(self):
        for fname, opts in FILES:
            self.parse_file(fname, opts)
        for c in list(self.order):
            if c.kind != "enum":
                self.parse_members(c)
        fields = set()
        for c in self.classes.values():
            fields |= set(c.fields)
        self.ambiguous = self.all_props & fields
        out = ["// GENERATED by tools/cs2cpp.py from /root/reference/Assets/Script/*.cs -- test infrastructure, do not edit,\n"
               "// do not commit (oracle/_ref/ is git-ignored).  Rewrite rules R1-R10 are documented in tools/cs2cpp.py.\n"
               "#include \"unity_shim.h\"\n"
               "namespace Footsies {\n"]
        for c in self.order:
            if c.kind != "enum":
                out.append(f"struct {c.name};\n")
        out.append("}  // namespace Footsies\n#include \"game_standins.h\"\nnamespace Footsies {\n")
        for c in self.order:
            if c.kind == "enum":
                out.append(self.emit_enum(c, ""))
        defs = []
        for c in self.sorted_classes():
            if c.kind != "enum":
                out.append(self.emit_class(c, "", defs))
        out.append("\n// ---------------------------------------------------------------- member definitions (R10)\n")
        out.extend(defs)
        out.append("}  // namespace Footsies\n")
        return "".join(out)


def emit_assets():
    """The ScriptableObject assets (Unity YAML) the scene wires into BattleCore.fighterDataList, as C++ that builds the
    same object graph: F00.asset -> F00_ActionDataContainer.asset (guid list, resolved through the .meta files) ->
    Actions/*.asset, and F00_AttackDataContainer.asset.  Data only; the YAML is read where it lies."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import gen_frame_data as g
    consts, attacks, actions = g.load_all()
    f00 = g.F00
    guid_of = {}
    adir = os.path.join(f00, "Actions")
    for fn in os.listdir(adir):
        if fn.endswith(".asset.meta"):
            meta = open(os.path.join(adir, fn)).read()
            guid_of[re.search(r"guid: (\w+)", meta).group(1)] = fn[:-len(".asset.meta")]
    cont = g.load_unity_yaml(os.path.join(f00, "F00_ActionDataContainer.asset"))["MonoBehaviour"]["actions"]
    order = [guid_of[e["guid"]] for e in cont]
    by_name = {a["actionName"]: a for a in actions}
    assert sorted(order) == sorted(by_name), "container and Actions/ directory disagree"
    att_order = [a["attackID"] for a in g.load_unity_yaml(
        os.path.join(f00, "F00_AttackDataContainer.asset"))["MonoBehaviour"]["attackDataList"]]
    by_att = {a["attackID"]: a for a in attacks}
    cf = g.cf
    o = ["// GENERATED by tools/cs2cpp.py from /root/reference/Assets/Fighter/F00/*.asset -- data only, do not commit.\n"
         "namespace Footsies {\nstatic Ref<FighterData> LoadF00() {\n    auto fd = New<FighterData>();\n"]
    w = o.append
    for k in ("startGuardHealth", "dashAllowFrame", "specialAttackHoldFrame"):
        w(f"    fd->{k} = {consts[k]};\n")
    w(f"    fd->forwardMoveSpeed = {cf(consts['forwardMoveSpeed'])};\n    fd->backwardMoveSpeed = {cf(consts['backwardMoveSpeed'])};\n")
    w(f"    fd->canCancelOnWhiff = {'true' if consts['canCancelOnWhiff'] else 'false'};\n")
    for k in ("baseHurtBoxRect", "basePushBoxRect"):
        w(f"    fd->{k}.Set({', '.join(cf(v) for v in consts[k])});\n")
    w(f"    fd->actionDataContainer = New<ActionDataContainer>();\n"
      f"    fd->actionDataContainer->actions = NewArray<Ref<ActionData>>({len(order)});\n")

    def se(h):
        return f"e->startEndFrame.x = {h['se'][0]}; e->startEndFrame.y = {h['se'][1]};"
    for idx, name in enumerate(order):
        a = by_name[name]
        w(f"    {{ auto a = New<ActionData>();   // Actions/{name}.asset\n")
        w(f"      a->actionID = {a['actionID']}; a->actionName = \"{name}\"; a->Type = (ActionType){a['type']}; "
          f"a->frameCount = {a['frameCount']}; a->isLoop = {str(bool(a['isLoop'])).lower()}; "
          f"a->loopFromFrame = {a['loopFromFrame']}; a->alwaysCancelable = {str(bool(a['alwaysCancelable'])).lower()};\n")
        w("      a->motions = NewArray<Ref<MotionFrameData>>(0); a->status = NewArray<Ref<StatusData>>(0);   // presentation / never read\n")
        w(f"      a->hitboxes = NewArray<Ref<HitboxData>>({len(a['hitboxes'])});\n")
        for i, h in enumerate(a["hitboxes"]):
            w(f"      {{ auto e = New<HitboxData>(); {se(h)} e->rect.Set({', '.join(cf(v) for v in h['rect'])}); "
              f"e->attackID = {h['attackID']}; e->proximity = {str(bool(h['proximity'])).lower()}; a->hitboxes[{i}] = e; }}\n")
        for key, cls in (("hurtboxes", "HurtboxData"), ("pushboxes", "PushboxData")):
            w(f"      a->{key} = NewArray<Ref<{cls}>>({len(a[key])});\n")
            for i, h in enumerate(a[key]):
                w(f"      {{ auto e = New<{cls}>(); {se(h)} e->rect.Set({', '.join(cf(v) for v in h['rect'])}); "
                  f"e->useBaseRect = {str(bool(h['useBaseRect'])).lower()}; a->{key}[{i}] = e; }}\n")
        w(f"      a->movements = NewArray<Ref<MovementData>>({len(a['movements'])});\n")
        for i, h in enumerate(a["movements"]):
            w(f"      {{ auto e = New<MovementData>(); {se(h)} e->velocity_x = {cf(h['velocity_x'])}; a->movements[{i}] = e; }}\n")
        w(f"      a->cancels = NewArray<Ref<CancelData>>({len(a['cancels'])});\n")
        for i, h in enumerate(a["cancels"]):
            adds = " ".join(f"e->actionID->Add({v});" for v in h["actionID"])
            w(f"      {{ auto e = New<CancelData>(); {se(h)} e->buffer = {str(bool(h['buffer'])).lower()}; "
              f"e->execute = {str(bool(h['execute'])).lower()}; {adds} a->cancels[{i}] = e; }}\n")
        w(f"      fd->actionDataContainer->actions[{idx}] = a; }}\n")
    w(f"    fd->attackDataContainer = New<AttackDataContainer>();\n"
      f"    fd->attackDataContainer->attackDataList = NewArray<Ref<AttackData>>({len(att_order)});\n")
    for idx, aid in enumerate(att_order):
        t = by_att[aid]
        fields = " ".join(f"t->{k} = {t[k]};" for k in ("attackID", "damageActionID", "guardActionID", "numberOfHit",
                                                         "vitalHealthDamage", "guardHealthDamage", "hitStunFrame",
                                                         "guardStunFrame", "guardBreakStunFrame"))
        w(f"    {{ auto t = New<AttackData>(); t->attackName = \"{t['attackName']}\"; {fields} "
          f"fd->attackDataContainer->attackDataList[{idx}] = t; }}\n")
    w("    fd->motionDataContainer = New<MotionDataContainer>();\n"
      "    fd->motionDataContainer->motionDataList = NewArray<Ref<MotionData>>(0);   // sprites: presentation\n")
    w("    return fd;\n}\n")
    w(f"static const float kSceneBattleAreaWidth = {cf(consts['battleAreaWidth'])};   // BattleScene.unity:273\n")
    w(f"static const float kFixedDeltaTime = {cf(consts['fixedDeltaTime'])};         // ProjectSettings/TimeManager.asset:6\n")
    w("}  // namespace Footsies\n")
    return "".join(o)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--output", required=True, help="generated C++ (classes + member definitions)")
    ap.add_argument("--assets", help="generated C++ that builds the F00 asset graph")
    args = ap.parse_args()
    if not os.path.isdir(REF_SCRIPTS):
        raise SystemExit(f"{REF_SCRIPTS} not found: the transliteration can only be generated where the reference lies")
    text = Translator().run()
    os.makedirs(os.path.dirname(os.path.abspath(args.output)), exist_ok=True)
    with open(args.output, "w") as f:
        f.write(text)
    print(f"wrote {args.output}: {text.count(chr(10))} lines", file=sys.stderr)
    if args.assets:
        with open(args.assets, "w") as f:
            f.write(emit_assets())
        print(f"wrote {args.assets}", file=sys.stderr)


if __name__ == "__main__":
    main()
