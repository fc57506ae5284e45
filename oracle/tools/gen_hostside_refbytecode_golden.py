"""Generate tests/golden/hostside_refbytecode.json: the host-side methods of Homoplasy_Counter.call() around the hot path,
executed from the REFERENCE'S OWN BYTECODE (compiled/Homoplasy_Counter.class) by the class-file interpreter:

    nexus_to_newick()                         HC.java:866-914    treetime's annotated_tree.nexus -> Newick text
    read_phenos()                             HC.java:1972-2019  pheno file -> HashMap + the case / control / missing totals it logs
    get_physical_positions()                  HC.java:1207-1235  map file -> int[]
    output_significance_assessments_binom()   HC.java:2137-2210  (+ Homoplasy_Events.toString 5574, Binomial_Test_Stat.toString
                                              5918, the Row comparators): the three result files

The JDK classes they touch are emulated from the JDK's documented behaviour (BufferedReader.readLine, String.split /
equalsIgnoreCase, StringBuilder.indexOf / delete, Integer.parseInt, HashMap, the stream calls, Arrays.parallelSort on objects
= a stable sort by the comparator, Arrays.toString(int[]), Double.compare / toString, BufferedWriter).  try/catch tables are
not interpreted: a Java exception ends the call; every catch block on these paths prints a message and calls System.exit(-1),
so an exception is recorded as {"exits": <exception class>}.

The writer runs on Homoplasy_Events / Binomial_Test_Stat objects filled from tests/golden/{events,assoc}_refbytecode.json —
themselves outputs of the reference's bytecode — so the fixture's three file texts are what the reference writes for those runs.
tests/test_hostside_refbytecode.py holds the Python host (poutine_b200/hostio.py) and the native host
(poutine_b200/host/hostio.cpp) to all of it, byte for byte.

run (in the authoring container, needs /root/reference):  python oracle/tools/gen_hostside_refbytecode_golden.py
"""
import functools
import json
import os
import struct
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refjvm  # noqa: E402
from gen_loaders_refbytecode_golden import JLineReader, LoaderJVM, SystemExit_, thrown  # noqa: E402
from minijvm import JArray, JavaThrow, JObject  # noqa: E402
from refjvm import JArrayList, JStream, JStringBuilder, JView, JWriter, _impl  # noqa: E402

HC = "Homoplasy_Counter"


def java_double_to_string(x):
    """java.lang.Double.toString (JDK >= 19): shortest decimal that round-trips (at least two digits), plain notation for
    1e-3 <= |x| < 1e7, otherwise computerised scientific notation."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if struct.pack(">d", x)[0] & 0x80 else "0.0"
    sign, ax = ("-" if x < 0 else ""), abs(x)
    from decimal import Decimal
    _, dg, ex = Decimal(repr(ax)).as_tuple()               # repr: the shortest digits that round-trip
    digits = "".join(map(str, dg)).lstrip("0")
    e10 = len(digits) - 1 + ex                             # decimal exponent of the first digit
    digits = digits.rstrip("0") or "0"
    if len(digits) == 1 and float("%.1e" % ax) == ax:      # two digits at least, the closest such decimal (4.9E-324)
        two = "%.1e" % ax
        digits, e10 = (two[0] + two[2]).rstrip("0") or "0", int(two.split("e")[1])
    if 1e-3 <= ax < 1e7:
        if e10 >= 0:
            return sign + digits[:e10 + 1].ljust(e10 + 1, "0") + "." + (digits[e10 + 1:] or "0")
        return sign + "0." + "0" * (-e10 - 1) + digits
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


_orig_jstr = refjvm.jstr


def _jstr(v, desc_char=None):
    if desc_char == "D" or isinstance(v, float):
        return java_double_to_string(float(v))
    if isinstance(v, JObject) and "$impl" not in v.fields:      # String.valueOf(Object) = obj.toString(), the reference's own
        return HostJVM.current.call(v, "toString", "()Ljava/lang/String;")
    return _orig_jstr(v, desc_char)


refjvm.jstr = _jstr


def java_parse_int(s):
    body = s[1:] if s[:1] in ("+", "-") else s
    if s is None or not body or any(c not in "0123456789" for c in body) or not -(1 << 31) <= int(s) < (1 << 31):
        e = JObject("java/lang/NumberFormatException")
        e.fields["$message"] = 'For input string: "%s"' % s
        raise JavaThrow(e)
    return int(s)


def java_split(s, sep):
    """String.split(regex) for a one-char literal regex: trailing empty strings removed; no match -> [s]"""
    parts = s.split(sep)
    if len(parts) == 1:
        return parts
    while parts and parts[-1] == "":
        parts.pop()
    return parts


class JIntStream:
    def __init__(self, items):
        self.items = items


class HostJVM(LoaderJVM):
    current = None

    def __init__(self):
        super().__init__()
        HostJVM.current = self

    def jdk(self, kind, cname, name, desc, args):
        if name == "<init>":
            o = args[0]
            if cname == "java/io/FileReader" and desc == "(Ljava/lang/String;)V":
                # File.separator is an opaque JDK static here (null): "<dir>" + null + "annotated_tree.nexus"
                path = args[1].replace("nullannotated_tree.nexus", "/annotated_tree.nexus")
                with open(path, encoding="latin-1", newline="") as f:
                    o.fields["$impl"] = ("text", f.read())
                return None
            if cname == "java/io/BufferedReader" and desc == "(Ljava/io/Reader;)V":
                o.fields["$impl"] = JLineReader(_impl(args[1])[1])
                return None
            if cname == "java/lang/StringBuilder" and desc == "(Ljava/lang/String;)V":
                sb = JStringBuilder()
                if args[1] is None:
                    raise JavaThrow("java/lang/NullPointerException")
                sb.parts.append(args[1])
                o.fields["$impl"] = sb
                return None
        if kind == "static":
            if cname == "java/lang/Integer" and name == "parseInt":
                return java_parse_int(args[0])
            if cname == "java/lang/Double" and name == "compare":
                a, b = args
                if a < b:
                    return -1
                if a > b:
                    return 1
                ba, bb = struct.unpack(">q", struct.pack(">d", a))[0], struct.unpack(">q", struct.pack(">d", b))[0]
                if a != a or b != b:                        # doubleToLongBits collapses NaNs; NaN is above everything
                    ba = 0x7ff8000000000000 if a != a else ba
                    bb = 0x7ff8000000000000 if b != b else bb
                return 0 if ba == bb else (-1 if ba < bb else 1)
            if cname == "java/util/Arrays" and name == "toString" and desc == "([I)Ljava/lang/String;":
                return "null" if args[0] is None else "[" + ", ".join(str(int(x)) for x in args[0].data) + "]"
            if cname == "java/util/Arrays" and name == "parallelSort" and desc == "([Ljava/lang/Object;Ljava/util/Comparator;)V":
                # objects: a stable merge sort (TimSort below the parallel granularity) by the comparator
                arr, cmp = args
                arr.data.sort(key=functools.cmp_to_key(lambda x, y: self.call_lambda(cmp, [x, y])))
                return None
            if cname == "java/util/stream/Collectors" and name == "toList":
                return "toList"
        recv = _impl(args[0]) if args and kind != "static" else None
        a = args[1:]
        if recv is None and kind != "static" and cname in ("java/lang/String", "java/lang/StringBuilder"):
            raise JavaThrow("java/lang/NullPointerException")       # e.g. readLine() returned null at end of file
        if isinstance(recv, JLineReader) and name in ("readLine", "close"):
            return super().jdk(kind, "java/io/LineNumberReader", name, desc, args)
        if isinstance(recv, JStringBuilder):
            text = "".join(recv.parts)
            if name == "indexOf":
                return text.find(a[0])
            if name == "delete":
                start, end = a
                end = min(end, len(text))
                if start < 0 or start > len(text) or start > end:
                    raise JavaThrow("java/lang/StringIndexOutOfBoundsException")
                recv.parts[:] = [text[:start] + text[end:]]
                return args[0]
        if isinstance(recv, str):
            if name == "equalsIgnoreCase":
                return int(a[0] is not None and recv.lower() == a[0].lower())
            if name == "split" and desc == "(Ljava/lang/String;)[Ljava/lang/String;":
                assert a[0] == "\t"
                return JArray(java_split(recv, "\t"), "Ljava/lang/String;")
        if isinstance(recv, int) and name == "intValue":
            return recv
        if isinstance(recv, (JView, JArrayList)) and name == "stream":
            return JStream(self._iterate(recv))
        if isinstance(recv, JStream):
            if name == "map":
                return JStream([self.call_lambda(a[0], [x]) for x in recv.items])
            if name == "mapToInt":
                return JIntStream([self.call_lambda(a[0], [x]) for x in recv.items])
            if name == "collect" and a[0] == "toList":
                return JArrayList(list(recv.items))
        if isinstance(recv, JIntStream):
            if name == "sum":
                return sum(recv.items)
            if name == "toArray":
                return JArray(list(recv.items), "I")
        if isinstance(recv, JWriter) and name == "newLine":
            recv.text.append("\n")
            return None
        return super().jdk(kind, cname, name, desc, args)


def new_counter(vm, **files):
    hc = JObject(HC)
    oo = JObject(HC + "$OutputOptions")
    for k in ("log", "debug", "out", "out_sorted_by_a1_maxT", "out_sorted_by_a2_maxT"):
        oo.fields[k] = JWriter()
    oo.fields["anc_recon_dir"] = files.get("anc_recon_dir")
    other = JObject(HC + "$InputFiles$OtherInputFiles")
    other.fields["pheno_filename"] = files.get("pheno")
    other.fields["map_filename"] = files.get("map")
    inp = JObject(HC + "$InputFiles")
    inp.fields["otherInputFiles"] = other
    hc.fields["outputOptions"], hc.fields["inputFiles"] = oo, inp
    hc.fields["DEBUG_MODE"] = 0
    vm.load(HC)
    return hc, oo


def guarded(fn):
    try:
        return fn()
    except SystemExit_:
        return {"exits": "System.exit"}
    except JavaThrow as e:
        return {"exits": thrown(e)["throws"]}


def run_nexus(raw: bytes):
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "annotated_tree.nexus"), "wb") as f:
        f.write(raw)
    vm = HostJVM()
    hc, _ = new_counter(vm, anc_recon_dir=d)
    r = guarded(lambda: {"newick": vm.call(hc, "nexus_to_newick", "()Ljava/lang/String;")})
    os.unlink(os.path.join(d, "annotated_tree.nexus"))
    os.rmdir(d)
    return r


def run_phenos(raw: bytes):
    fd, p = tempfile.mkstemp()
    os.write(fd, raw)
    os.close(fd)
    vm = HostJVM()
    hc, oo = new_counter(vm, pheno=p)

    def go():
        m = vm.call(hc, "read_phenos", "()Ljava/util/HashMap;")
        return {"entries": {c[0]: c[1] for c in _impl(m).ordered()}, "hashmap_order": [c[0] for c in _impl(m).ordered()],
                "cases": hc.fields["tot_num_sample_cases"], "controls": hc.fields["tot_num_sample_controls"], "log": "".join(oo.fields["log"].text)}
    r = guarded(go)
    os.unlink(p)
    return r


def run_map(raw: bytes):
    fd, p = tempfile.mkstemp()
    os.write(fd, raw)
    os.close(fd)
    vm = HostJVM()
    hc, _ = new_counter(vm, map=p)
    r = guarded(lambda: {"positions": [int(x) for x in vm.call(hc, "get_physical_positions", "()[I").data]})
    os.unlink(p)
    return r


def dbl(bits_hex):
    return struct.unpack(">d", bytes.fromhex(bits_hex))[0]


def run_writer(ev_case, as_case):
    vm = HostJVM()
    hc, oo = new_counter(vm)
    events = {e["segsite_id"]: e for e in ev_case["events"]}
    sites, stats, rows = JArrayList(), [], []
    nan = float("nan")
    for s in as_case["sites"]:
        e = events[s["segsite_id"]]
        he = JObject(HC + "$Homoplasy_Events")
        he.fields.update(segsite_ID=e["segsite_id"], physical_pos=e["physical_pos"], allele1=ord(e["allele1"]), allele2=ord(e["allele2"]),
                         a1_count=e["a1_count"], a2_count=e["a2_count"], a1_count_extant_only=e["a1_extant"],
                         a2_count_extant_only=e["a2_extant"], obs_counts=JArray(list(s["obs_counts"]), "I"))
        st = JObject(HC + "$Binomial_Test_Stat")
        vals = {k: (nan if s[k].startswith("7ff8") or s[k].startswith("fff8") else dbl(s[k]))
                for k in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2", "familywise_a1", "familywise_a2")}
        st.fields.update(r_a1=s["r_a1"], r_a2=s["r_a2"], r_maxT_a1=s["r_maxT_a1"], r_maxT_a2=s["r_maxT_a2"],
                         pointwise_pvalue_a1=vals["pointwise_a1"], pointwise_pvalue_a2=vals["pointwise_a2"],
                         obs_binom_pvalue_a1=vals["obs_p_a1"], obs_binom_pvalue_a2=vals["obs_p_a2"],
                         familywise_pvalue_a1=vals["familywise_a1"], familywise_pvalue_a2=vals["familywise_a2"])
        sites.d.append(he)
        stats.append(st)
        rows.append({"segsite_id": e["segsite_id"], "physical_pos": e["physical_pos"], "allele1": e["allele1"], "allele2": e["allele2"],
                     "a1_count": e["a1_count"], "a2_count": e["a2_count"], "a1_extant": e["a1_extant"], "a2_extant": e["a2_extant"],
                     **{k: s[k] for k in ("obs_counts", "r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2", "obs_p_a1", "obs_p_a2", "pointwise_a1",
                                          "pointwise_a2", "familywise_a1", "familywise_a2")}})
    lst = JObject("java/util/ArrayList")
    lst.fields["$impl"] = sites
    vm.call(hc, "output_significance_assessments_binom", "(Ljava/util/ArrayList;[L" + HC + "$Binomial_Test_Stat;)V", lst,
            JArray(stats, "L" + HC + "$Binomial_Test_Stat;"))
    return {"name": as_case["name"], "rows": rows, "out": "".join(oo.fields["out"].text),
            "sorted_by_a2_maxT": "".join(oo.fields["out_sorted_by_a2_maxT"].text),
            "sorted_by_a1_maxT": "".join(oo.fields["out_sorted_by_a1_maxT"].text)}


NEXUS = [
    b"#NEXUS\nBegin Taxa;\n Dimensions NTax=3;\nEnd;\nBegin Trees;\n Tree tree1=[&U](A:1[&mutations=\"A1C\"],(B:2,C:3)N1[&x]:0.5)N0:0;\nEnd;\n",
    b"#NEXUS\r\nbegin TREES;\r\n Tree t=[&U](A[&x=1]:1,B:2)R[&y];\r\nEnd;\r\n",
    b"#NEXUS\nBegin Trees;\n(A,B)R;\nEnd;\n", b"#NEXUS\nBegin Trees;\nTree t = [&R] ((A,B)[&a,b],C);", b"#NEXUS\n Begin Trees;\n(A,B)R;\nBegin Trees;\n(C,D)Q;\n",
    b"#NEXUS\nBegin Taxa;\nEnd;\n", b"#NEXUS\nBegin Trees;\n", b"#NEXUS\nBegin Trees;\nTree t = A;\n", b"#NEXUS\nBegin Trees;\n(A]:1,B[x:2);\n",
    b"#NEXUS\nBegin Trees;\n(A[&x:1,B:2);\n", b"#NEXUS\nBegin Trees;\n(A[[x]]:1,B:2);\n", b"#NEXUS\nBEGIN TREES;\n\n(A,B)R;\n",
]
PHENOS = [
    b"s1\t1\ns2\t0\ns3\tNA\n", b"s1\t1\r\ns2\t0\textra\ns3\tNA\ns1\t0\ns4\t1\t\t\n", b"s1\t1\ns2\n", b"s1\t1\ns2\t\n", b"s1\t1\n\ns3\t0\n", b"",
    b"a\t1\nb\t1\nc\t1\n", b"a\t01\nb\t 1\nc\t1 \nd\t0\n", b"s9\t0\ns10\t1\nAa\t1\nBB\t0\n", b"s1 1\ns2 0\n", b"\tx\t1\n",
    b"".join(b"S%06d\t%d\n" % (i, i % 3 == 0) for i in range(40)),
]
MAPS = [
    b"1\tsnp1\t0\t17\n1\tsnp2\t0\t23\n", b"1\tsnp1\t0\t17\n1\tsnp2\t0\t+23\n9\n-4\t\t\n", b"17\r\n23\r\n", b"1\tsnp\t0\t12x\n", b"1\tsnp\t0\t 12\n",
    b"1\t2147483647\n", b"1\t2147483648\n", b"\n", b"", b"1 2 3\n", b"1\tsnp\t0\t007\n", b"5\n\n6\n", b"1\t\t\n",
]


def fuzz_cases():
    """seeded byte soups for the three text readers"""
    import random
    rnd = random.Random(20262)
    ph, mp, nx = [], [], []
    pp = [b"s1", b"s2", b"s3", b"\t", b"\t", b"\n", b"\r\n", b"\r", b"0", b"1", b"NA", b" ", b"", b"x"]
    for _ in range(40):
        ph.append(b"".join(rnd.choice(pp) for _ in range(rnd.randint(0, 14))))
    for _ in range(80):                                     # mostly well-formed rows, now and then a broken one
        rows = []
        for i in range(rnd.randint(1, 8)):
            sid = rnd.choice([b"s%d" % rnd.randint(1, 6), b"S 0%d" % i, b"id_7", b"a.b-c"])
            lab = rnd.choice([b"0", b"1", b"1", b"NA", b"", b"2", b"01", b" 1"])
            row = sid + b"\t" + lab + rnd.choice([b"", b"", b"\textra", b"\t", b"\t\t"])
            if rnd.random() < 0.04:
                row = rnd.choice([sid, b"", sid + b" " + lab])
            rows.append(row + rnd.choice([b"\n", b"\n", b"\r\n", b"\r"]))
        if rnd.random() < 0.3:
            rows[-1] = rows[-1].rstrip(b"\r\n")              # no final line terminator
        ph.append(b"".join(rows))
    mm = [b"1", b"23", b"-4", b"+5", b"\t", b"\t", b"\n", b"\r\n", b"snp", b" ", b"007", b"2147483647", b"2147483648", b"x", b"1_0", b"\r"]
    for _ in range(40):
        mp.append(b"".join(rnd.choice(mm) for _ in range(rnd.randint(0, 12))))
    for _ in range(80):
        rows = []
        for i in range(rnd.randint(1, 8)):
            pos = rnd.choice([b"%d" % rnd.randint(0, 5000000), b"+%d" % i, b"-%d" % i, b"00%d" % i, b"2147483647", b"-2147483648"])
            row = rnd.choice([b"", b"1\t", b"1\tsnp%d\t0\t" % i, b"chr 1\t"]) + pos + rnd.choice([b"", b"", b"\t", b"\t\t"])
            if rnd.random() < 0.04:
                row = rnd.choice([b"", b"1\tx", pos + b" ", b"1\t2147483648", b"1\t1.0"])
            rows.append(row + rnd.choice([b"\n", b"\n", b"\r\n", b"\r"]))
        if rnd.random() < 0.3:
            rows[-1] = rows[-1].rstrip(b"\r\n")
        mp.append(b"".join(rows))
    nn = [b"#NEXUS\n", b"Begin Trees;\n", b"begin trees;\r\n", b"BEGIN TREES;", b" Begin Trees;\n", b"End;\n", b"\n", b"Tree t=", b"[&U]", b"(A:1,B:2)R;",
          b"(A[&x]:1,", b"B)", b"[", b"]", b"(", b";", b"\r", b" "]
    for _ in range(30):
        nx.append(b"#NEXUS\n" + b"".join(rnd.choice(nn) for _ in range(rnd.randint(0, 10))))
    for _ in range(50):                                     # treetime-shaped files with varied spelling, annotations and line ends
        eol = rnd.choice([b"\n", b"\n", b"\r\n", b"\r"])
        begin = rnd.choice([b"Begin Trees;", b"begin trees;", b"BEGIN TREES;", b"Begin Trees;", b"Begin trees;", b"Begin Trees; ", b"Begin  Trees;"])
        ann = lambda: rnd.choice([b"", b"", b'[&mutations="A1C,G7T"]', b"[&x]", b"[]", b"[&a=[1]"])
        tree = b"(A" + ann() + b":1,(B:2" + ann() + b",C" + ann() + b":3)N1" + ann() + b":0.5)N0" + ann() + b":0;"
        head = rnd.choice([b" Tree tree1=[&U]", b"Tree t = [&R] ", b"", b"\tTREE x=", b"Tree tree1="])
        body = b"#NEXUS" + eol + b"Begin Taxa;" + eol + b" Dimensions NTax=3;" + eol + b"End;" + eol + begin + eol + head + tree + eol + b"End;" + eol
        if rnd.random() < 0.1:
            body = body.replace(tree, tree.replace(b"(", b"", 1) if rnd.random() < 0.5 else b"")
        nx.append(body)
    return nx, ph, mp


def main():
    fx, fp, fm = fuzz_cases()
    NEXUS.extend(fx)
    PHENOS.extend(fp)
    MAPS.extend(fm)
    out = {"comment": "generated by oracle/tools/gen_hostside_refbytecode_golden.py from compiled/Homoplasy_Counter.class; do not edit",
           "nexus": [], "phenos": [], "map": [], "writer": []}
    for raw in NEXUS:
        r = run_nexus(raw)
        out["nexus"].append({"raw_latin1": raw.decode("latin-1"), **r})
        print("nexus  %-50r -> %r" % (raw[7:55], r))
    for raw in PHENOS:
        r = run_phenos(raw)
        out["phenos"].append({"raw_latin1": raw.decode("latin-1"), **r})
        print("phenos %-50r -> %r" % (raw[:48], {k: v for k, v in r.items() if k != "log"} if len(raw) < 100 else "..."))
    for raw in MAPS:
        r = run_map(raw)
        out["map"].append({"raw_latin1": raw.decode("latin-1"), **r})
        print("map    %-50r -> %r" % (raw[:48], r))
    ev = json.load(open(os.path.join(ROOT, "tests", "golden", "events_refbytecode.json")))["cases"]
    asc = json.load(open(os.path.join(ROOT, "tests", "golden", "assoc_refbytecode.json")))["cases"]
    for e, a in zip(ev, asc):
        assert e["name"] == a["name"]
        w = run_writer(e, a)
        out["writer"].append(w)
        print("writer %-24s %3d rows -> %d / %d / %d lines" % (w["name"], len(w["rows"]), w["out"].count("\n"),
                                                              w["sorted_by_a2_maxT"].count("\n"), w["sorted_by_a1_maxT"].count("\n")))
    path = os.path.join(ROOT, "tests", "golden", "hostside_refbytecode.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
