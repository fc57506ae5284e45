"""Generate tests/golden/loaders_refbytecode.json: what the REFERENCE'S OWN LOADER BYTECODE makes of tricky inputs.

The two readers in front of the hot path have no source in the reference — they ship as bytecode only:

    coevolution.jar  org/gersteinlab/coevolution/core/io/NewickTreeTokenizer + NewickTreeReader.readTree()
                     (called by build_tree, HC.java:1144-1160, on a StringReader of the Newick text)
    compiled/Fasta_Manager.class next() + compiled/Fasta_Record.class getHeader() / getSequence()
                     (called by build_seg_sites, HC.java:1252-1266)

Both are executed here by the class-file interpreter (refjvm.RefJVM, no JVM in the authoring container) with the few
JDK classes they touch emulated from the JDK's documented behaviour: StringReader / PushbackReader (read, unread),
StringBuffer, Character.isWhitespace, Float.parseFloat; FileReader / LineNumberReader (readLine with \\n, \\r, \\r\\n
terminators, mark / reset, getLineNumber), String.substring / trim.  try/catch tables are not interpreted: a Java
exception ends the call and is recorded with its class and message, System.exit is recorded as "exit".

Per Newick case the fixture holds the parsed tree (parent index, id or null, distToParent as float bits, in node creation
order = NewickTreeNode construction order; the NewickTree leaf list) or the exception; per FASTA case the (header, sequence)
pairs next() returns until null, or the exit.  tests/test_loaders_refbytecode.py checks both hosts' loaders (Python
hostio.parse_newick, C++ parse_newick via poutine_b200_host, the mmap'ed FASTA index + packer in csrc/fasta.cpp)
against it.  This pins SURVEY section 8f rows 1 and 3 to the reference itself.

run (in the authoring container, needs /root/reference):  python oracle/tools/gen_loaders_refbytecode_golden.py
"""
import json
import os
import struct
import sys
import tempfile
import unicodedata

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from minijvm import JavaThrow, JObject, _f32  # noqa: E402
from refjvm import JStringBuilder, RefJVM, _impl  # noqa: E402

PKG = "org/gersteinlab/coevolution/core/"
NODE, TREE = PKG + "data/NewickTreeNode", PKG + "data/NewickTree"
TOKENIZER, READER = PKG + "io/NewickTreeTokenizer", PKG + "io/NewickTreeReader"


class SystemExit_(Exception):
    pass


class JPushback:
    """java.io.PushbackReader(new StringReader(text)) with the default one-char pushback buffer"""
    def __init__(self, text):
        self.s, self.i, self.pushed = text, 0, None


class JLineReader:
    """java.io.LineNumberReader(new FileReader(path)): readLine() ends a line at \\n, \\r or \\r\\n; mark / reset"""
    def __init__(self, text):
        self.lines, i, n = [], 0, len(text)
        while i < n:
            j = i
            while j < n and text[j] not in "\n\r":
                j += 1
            self.lines.append(text[i:j])
            if j < n and text[j] == "\r" and j + 1 < n and text[j + 1] == "\n":
                j += 1
            i = j + 1
        self.k, self.marked = 0, None


def java_is_whitespace(c):
    ch = chr(c)
    if ch in "\t\n\x0b\x0c\r\x1c\x1d\x1e\x1f":
        return True
    return unicodedata.category(ch) in ("Zs", "Zl", "Zp") and ch not in "\u00a0\u2007\u202f"


def java_parse_float(s):
    """Float.parseFloat: trims whitespace, optional sign, decimal (or NaN / Infinity), optional f|F|d|D suffix"""
    t = s.strip(" \t\n\r\x0b\x0c") if s is not None else None
    if t is None:
        raise JavaThrow("java/lang/NullPointerException")
    body = t[:-1] if t[-1:] in ("f", "F", "d", "D") and len(t) > 1 else t
    ok = body not in ("", "+", "-") and all(ch in "0123456789.eE+-" for ch in body)
    if body.lstrip("+-") in ("NaN", "Infinity"):
        ok = True
    if ok:
        try:
            return _f32(float(body.replace("Infinity", "inf").replace("NaN", "nan")))
        except ValueError:
            pass
    e = JObject("java/lang/NumberFormatException")
    e.fields["$message"] = 'For input string: "%s"' % s
    raise JavaThrow(e)


class LoaderJVM(RefJVM):
    def invoke(self, kind, cname, name, desc, args):
        if name == "<init>" and cname in ("java/lang/Exception", "java/lang/Throwable", "java/lang/RuntimeException"):
            args[0].fields["$message"] = args[1] if desc.startswith("(Ljava/lang/String;") else None
            return None
        return super().invoke(kind, cname, name, desc, args)

    def jdk(self, kind, cname, name, desc, args):
        if name == "<init>":
            o = args[0]
            if cname == "java/io/StringReader":
                o.fields["$impl"] = ("text", args[1])
                return None
            if cname == "java/io/PushbackReader" and desc == "(Ljava/io/Reader;)V":
                o.fields["$impl"] = JPushback(_impl(args[1])[1])
                return None
            if cname == "java/lang/StringBuffer" and desc == "()V":
                o.fields["$impl"] = JStringBuilder()
                return None
            if cname == "java/io/FileReader" and desc == "(Ljava/lang/String;)V":
                with open(args[1], encoding="latin-1", newline="") as f:
                    o.fields["$impl"] = ("text", f.read())
                return None
            if cname == "java/io/LineNumberReader" and desc == "(Ljava/io/Reader;)V":
                o.fields["$impl"] = JLineReader(_impl(args[1])[1])
                return None
        if kind == "static":
            if cname == "java/lang/Character" and name == "isWhitespace":
                return int(java_is_whitespace(args[0]))
            if cname == "java/lang/Float" and name == "parseFloat":
                return java_parse_float(args[0])
            if cname == "java/lang/System" and name == "exit":
                raise SystemExit_(args[0])
        recv = _impl(args[0]) if args and kind != "static" else None
        a = args[1:]
        if isinstance(recv, JPushback):
            if name == "read" and desc == "()I":
                if recv.pushed is not None:
                    c, recv.pushed = recv.pushed, None
                    return c
                if recv.i >= len(recv.s):
                    return -1
                recv.i += 1
                return ord(recv.s[recv.i - 1])
            if name == "unread":
                assert recv.pushed is None, "pushback buffer overflow"
                recv.pushed = a[0]
                return None
            if name == "close":
                return None
        if isinstance(recv, JStringBuilder) and cname == "java/lang/StringBuffer":
            if name == "append" and desc.startswith("(C)"):
                recv.parts.append(chr(a[0]))
                return args[0]
            if name == "length":
                return sum(len(p) for p in recv.parts)
            if name == "toString":
                return "".join(recv.parts)
        if isinstance(recv, JStringBuilder) and name == "append" and desc.startswith("(I)"):
            recv.parts.append(str(a[0]))
            return args[0]
        if isinstance(recv, JLineReader):
            if name == "readLine":
                if recv.k >= len(recv.lines):
                    return None
                recv.k += 1
                return recv.lines[recv.k - 1]
            if name == "mark":
                recv.marked = recv.k          # the lines read between mark and reset are header lines, far below the 1000-char limit
                return None
            if name == "reset":
                recv.k = recv.marked
                return None
            if name == "getLineNumber":
                return recv.k
            if name == "close":
                return None
        if isinstance(recv, str):
            if name == "substring" and desc == "(I)Ljava/lang/String;":
                if a[0] > len(recv):
                    raise JavaThrow("java/lang/StringIndexOutOfBoundsException")
                return recv[a[0]:]
            if name == "trim":
                i, j = 0, len(recv)
                while i < j and recv[i] <= " ":
                    i += 1
                while j > i and recv[j - 1] <= " ":
                    j -= 1
                return recv[i:j]
            if name == "charAt" and not 0 <= a[0] < len(recv):
                raise JavaThrow("java/lang/StringIndexOutOfBoundsException")
        return super().jdk(kind, cname, name, desc, args)


def thrown(e):
    o = e.obj
    if isinstance(o, str):
        return {"throws": o.split("/")[-1], "message": None}
    return {"throws": o.cls.split("/")[-1], "message": o.fields.get("$message")}


def run_newick(text):
    vm = LoaderJVM()
    try:
        sr = vm.new("java/io/StringReader", "(Ljava/lang/String;)V", text)
        tok = vm.new(TOKENIZER, "(Ljava/io/Reader;)V", sr)
        rd = vm.new(READER, "(L" + TOKENIZER + ";)V", tok)
        tree = vm.call(rd, "readTree", "()L" + TREE + ";")
    except JavaThrow as e:
        return thrown(e)
    # node numbering = construction order = preorder of the nested brackets (what both hosts use)
    nodes, index = [], {}

    def walk(n, par):
        index[id(n)] = len(nodes)
        me = len(nodes)
        dist = n.fields.get("distToParent", 0.0)
        nodes.append({"parent": par, "id": n.fields.get("id"), "dist_bits": "%08x" % struct.unpack("<I", struct.pack("<f", float(dist)))[0]})
        kids = n.fields.get("children")
        for c in (_impl(kids).d if kids is not None else []):
            assert c.fields.get("parent") is n
            walk(c, me)

    walk(tree.fields["root"], -1)
    leaves = [index[id(n)] for n in _impl(vm.call(tree, "getLeafNodes", "()Ljava/util/List;")).d]
    return {"nodes": nodes, "leaves": leaves}


def run_fasta(raw: bytes):
    vm = LoaderJVM()
    fd, path = tempfile.mkstemp(suffix=".fasta")
    os.write(fd, raw)
    os.close(fd)
    out = []
    try:
        fm = vm.new("Fasta_Manager", "(Ljava/lang/String;)V", path)
        while True:
            rec = vm.call(fm, "next", "()LFasta_Record;")
            if rec is None:
                break
            out.append([vm.call(rec, "getHeader", "()Ljava/lang/String;"), vm.call(rec, "getSequence", "()Ljava/lang/String;")])
        return {"records": out}
    except SystemExit_:
        return {"records": out, "exit": True, "console": "".join(vm.console.text).strip()}
    except JavaThrow as e:
        return dict(thrown(e), records=out)
    finally:
        os.unlink(path)


NEWICK_CASES = [
    "(A,B)R;", "(A:0.1,B:0.25)R:0.5;", " ( A:0.1 ,\n B :2e-3)R ; trailing garbage", "((A,B)I1,(C,D)I2,E)ROOT;", "(A,,(B,C));", "(,);", "(A,B);",
    "((A:1,B:2)N1:3,(C:4,(D:5,E:6)N3:7)N2:8,F:9)N0;", "(A:1e-3,B:1.5E2,C:.5,D:5.,E:-1)R;", "(A:1f,B:2D,C:1e2F,D:+3.5,E:-0.0)R;",
    "A;", "A:1;", ";", ":1;", "(A B,C)R;", "(A\tB,C)R;", "(A,C\u00a0D\x0bE\x1cF)R;", "('A b',C)R;", "(A[x],B)R;", "((A,B)),C);", "(A,B", "(A,B)", "(A,B));", "(A:x,B);",
    "(A:,B);", "(A:1:2,B);", "((A,B);", ");", "(A,B)R S;", "(A,B)R:1 X;", "(A,B):1;", "(A,(B,C)I:0.5,)R;", "(A,B)R;(C,D)Q;", "", "   ", "(A;B)R;",
    "(A,B)R:1.0.0;", "(:1,:2):3;", "((((A))));", "(A,B,C,D,E,F,G,H)R;", "(A:1e400,B:1e-400)R;",
]

FASTA_CASES = [
    b">A\nACGT\n>B\nAGGT\n", b">A\nAC\nGT\n>B\nAG\nGT", b">A\r\nACGT\r\n>B\r\nAGGT\r\n", b">A\rACGT\r>B\rAGGT\r", b"> A b  \nACGT\n>\tB\t\nAG GT\n",
    b">A\nAC\n\nGT\n>B\n\n\nAGGT\n", b">A\n>B\nACGT\n", b">A\nacgt\n>B\nNN-?\n", b">A\nACGT \n>B\n AGGT\n", b">A\nACGT\n>B\nAGG\n", b">A", b">A\n", b"",
    b"ACGT\n>B\nAGGT\n", b"\n>A\nACGT\n", b">A\nACGT\n\n", b">A\nAC>GT\n>B\nAGGT\n", b">A\nACGT\n>A\nTTTT\n", b">\nACGT\n>B\nAGGT\n",
    b">A\n" + b"ACGT" * 400 + b"\n>B\n" + b"AGGT" * 400 + b"\n", b">A\n" + b"\n".join(b"ACGTA"[: 1 + i % 5] for i in range(40)) + b"\n>B\nT\n",
    b">A\nACGT\n >B\nAGGT\n",
]


def fuzz_cases():
    """seeded token soups: what a hand-written list does not think of"""
    import random
    rnd = random.Random(20261)
    nw, fa = [], []
    pieces = ["(", ")", ",", ":", ";", " ", "\t", "\n", "A", "B", "C1", "x_y", "1", "0.5", "2e-3", "-1", ".", "e", "'", "[", "]", "\u00a0", "+", "1f", "NaN"]
    for _ in range(260):
        k = rnd.randint(1, 14)
        nw.append("".join(rnd.choice(pieces) for _ in range(k)))
    for _ in range(60):                                   # mostly well-formed trees with one random edit
        base = rnd.choice(["((A:1,B:2)N1:3,(C:4,D:5)N2:6)R;", "(A,B,(C,D)I)R;", "((A,(B,C)X)Y,D);", "(A:0.1,B:0.2,C:0.3);"])
        i = rnd.randrange(len(base))
        edit = rnd.choice(["del", "ins", "sub"])
        ch = rnd.choice(pieces)
        nw.append(base[:i] + ("" if edit == "del" else ch) + (base[i:] if edit == "ins" else base[i + 1:]))
    fpieces = [b">", b"A", b"C", b"G", b"T", b"N", b"-", b"a", b" ", b"\t", b"\n", b"\r", b"\r\n", b"ACGT", b">S1", b">n 2 ", b"\n>", b"\r>"]
    for _ in range(200):
        k = rnd.randint(0, 12)
        raw = b"".join(rnd.choice(fpieces) for _ in range(k))
        if rnd.random() < 0.7:
            raw = b">" + raw                              # most start with a header, as a FASTA file must
        fa.append(raw)
    return nw, fa


def main():
    fz_nw, fz_fa = fuzz_cases()
    NEWICK_CASES.extend(fz_nw)
    FASTA_CASES.extend(fz_fa)
    out = {"comment": "generated by oracle/tools/gen_loaders_refbytecode_golden.py from the reference's bytecode (coevolution.jar "
                      "NewickTreeTokenizer/NewickTreeReader, compiled/Fasta_Manager.class, Fasta_Record.class); do not edit",
           "newick": [], "fasta": []}
    for text in NEWICK_CASES:
        r = run_newick(text)
        out["newick"].append({"text": text, **r})
        print("newick %-48r -> %s" % (text[:46], r.get("throws") and (r["throws"], r["message"]) or "%d nodes" % len(r["nodes"])))
    for raw in FASTA_CASES:
        r = run_fasta(raw)
        out["fasta"].append({"raw_latin1": raw.decode("latin-1"), **r})
        print("fasta  %-48r -> %s" % (raw[:46], (r.get("exit") and "EXIT " + r["console"][:50]) or r.get("throws") or
                                      [(h, len(s)) for h, s in r["records"]][:4]))
    path = os.path.join(ROOT, "tests", "golden", "loaders_refbytecode.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
