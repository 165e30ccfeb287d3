"""The transliteration itself (tools/cs2cpp.py): it must rewrite SYNTAX only.  For every member of every transliterated
class the "logic skeleton" -- the sequence of numeric literals, of the operators that carry logic (== != <= >= && || ! ~ ^ %
/ * + - & | ++ -- += -= *= /= |= &= ^= ?) and of the identifiers -- is extracted from the C# body and from the generated C++
body and compared, after undoing exactly the documented rewrite rules (R1-R10).  Anything the tool dropped, duplicated,
reordered or invented inside a body fails here.  Runs where /root/reference lies (the authoring container)."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/Assets/Script"), reason="/root/reference not present")

LOGIC_OPS = {"==", "!=", "<=", ">=", "&&", "||", "!", "~", "^", "%", "/", "*", "+", "-", "&", "|", "++", "--", "+=", "-=",
             "*=", "/=", "|=", "&=", "^=", "?"}
# identifiers the rewrite rules introduce or rename (everything else must survive verbatim)
CS_TO_CPP = {"var": "auto", "null": "nullptr", "uint": "uint32_t", "string": "std", "foreach": "for", "in": None, "new": None,
             "is": None, "this": "this"}
CPP_NOISE = {"Ref", "New", "NewArray", "CsArray", "Is", "_o", "auto", "std", "function", "void", "return", "nullptr", "for",
             "uint32_t", "string", "if", "Footsies"}
CS_NOISE = {"var", "null", "uint", "string", "foreach", "in", "new", "is", "return", "if", "void", "int", "float", "bool", "for"}


def skeleton(tokens, cpp):
    import cs2cpp
    out = []
    for t in tokens:
        if t.kind == "num":
            s = t.text.lower().rstrip("f").rstrip("u")
            out.append(("num", float(s) if "." in s else float(int(s))))
        elif t.kind == "op" and t.text in LOGIC_OPS:
            out.append(("op", t.text))
        elif t.kind == "id":
            name = t.text
            if (cpp and name in CPP_NOISE) or (not cpp and name in CS_NOISE):
                continue
            if name in ("int", "float", "bool", "double"):
                continue
            out.append(("id", name))
    return out


def test_bodies_keep_their_logic_skeleton():
    import cs2cpp
    tr = cs2cpp.Translator()
    text = tr.run()
    defs = text[text.index("member definitions (R10)"):]
    checked = 0
    for info in tr.classes.values():
        if info.kind == "enum":
            continue
        toks = info.pending[0]
        for kind, m in info.members:
            if kind not in ("method", "prop", "ctor") or m.get("body") is None:
                continue
            full = f"{info.name}.{m['name']}"
            if full in info.opts.get("skip", {}) or full in info.opts.get("replace", {}):
                continue
            b0, b1 = m["body"]
            src = [t for t in toks[b0:b1 + 1] if t.kind not in ("ws", "comment")]
            # R9: Debug.* statements are dropped
            cs_text = "".join(t.text if t.kind != "ws" else " " for t in toks[b0:b1 + 1] if t.kind != "comment")
            cs_text = re.sub(r"Debug\s*\.\s*\w+\s*\((?:[^()]|\([^()]*\))*\)\s*;", ";", cs_text)
            cs_text = cs_text.replace("?.", ".")
            # locate the generated definition by its marker comment and signature
            qual = info.qual()
            head = f"// {info.file}:{m['line']}\n"
            i = defs.index(head + ("auto " if kind != "ctor" else "") + f"{qual}::{m['name']}(")
            j = defs.index("{", i)
            depth, k = 0, j
            while True:
                depth += defs[k] == "{"
                depth -= defs[k] == "}"
                if depth == 0:
                    break
                k += 1
            cpp_text = defs[j:k + 1]
            cpp_text = cpp_text.replace("->", ".").replace("::", ".").replace("[&]", "")
            cpp_text = re.sub(r"/\* Debug\.Log dropped \*/", "", cpp_text)
            cpp_text = re.sub(r"/\* dropped by cs2cpp[^*]*\*/", "", cpp_text)
            cpp_text = re.sub(r"if \(\((\w+(?:\.\w+)*)\) != nullptr\) \(\1\)", r"\1", cpp_text)       # R6 `a?.b` spelled out
            cpp_text = re.sub(r"\(\)", "( )", cpp_text)
            a = skeleton(cs2cpp.tokenize(cs_text), cpp=False)
            b = skeleton(cs2cpp.tokenize(cpp_text), cpp=True)
            # drop type names (they appear in both but Ref<>-wrapped / cast syntax may differ in count): compare the rest
            types = set(tr.classes) | set(cs2cpp.SHIM_REF) | set(cs2cpp.SHIM_VALUE) | {"List", "Queue", "Dictionary", "Task", "System", "Action"}
            a = [x for x in a if not (x[0] == "id" and x[1] in types)]
            b = [x for x in b if not (x[0] == "id" and x[1] in types)]
            assert a == b, f"{full} ({info.file}:{m['line']}): logic skeleton differs\n C#: {a[:60]}\nC++: {b[:60]}"
            checked += 1
    assert checked > 120


def test_only_the_declared_exceptions_are_not_transliterated():
    import cs2cpp
    skipped = {k for _, o in cs2cpp.FILES for k in list(o.get("skip", {})) + list(o.get("replace", {}))}
    assert skipped == {"Fighter.GetCurrentMotionSprite", "TrainingManager.Setup", "InputData.ShallowCopy"}
    dropped = [p for _, o in cs2cpp.FILES for p in o.get("drop_lines", [])]
    assert len(dropped) == 1 and "Task" in dropped[0]
    # the whole engine is in: every hot-path file of SURVEY.md section 8(c)
    files = {f for f, _ in cs2cpp.FILES}
    assert {"Fighter.cs", "BattleAI.cs", "BattleCore.cs", "ActionData.cs", "FighterData.cs", "AttackData.cs", "InputData.cs",
            "TrainingManager.cs", "TrainingBattleAIActor.cs", "EnvironmentState.cs", "FighterState.cs", "BattleState.cs"} <= files
