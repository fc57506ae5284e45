"""Runs the reference's OWN compiled event-identification code — TEST INFRASTRUCTURE ONLY.

`/root/reference/compiled/Homoplasy_Counter.class` (+ inner classes) and `coevolution.jar` (NewickTree, NewickTreeNode)
are the reference as it ships; there is no JVM in the authoring container, so `minijvm.MiniJVM` interprets their
bytecode and this module supplies what that bytecode expects from the JDK:

  * java.util.HashMap / HashSet / ArrayList / Iterator / Map.Entry with the JDK's observable behaviour, including
    HashMap's ITERATION ORDER (bucket index = (h ^ h>>>16) & (capacity-1), insertion order inside a bucket, capacity 16
    doubling at load factor 0.75) — `identify_major_minor_alleles` (HC.java:1712-1743) depends on it;
  * java.lang.Character boxing, String.hashCode/equals;
  * the two invokedynamic bootstraps javac emits: LambdaMetafactory (the lambda BODIES are the reference's bytecode,
    e.g. lambda$get_alleles$2) and StringConcatFactory;
  * the java.util.stream calls of get_alleles (HC.java:1593-1603): ArrayList.stream().filter(p).collect(groupingBy(f));
  * sinks for the console / log writers (what the reference writes to its log is captured: the class counters of
    HC.java:1477-1480 are parsed from there).

Nothing here restates the reference's algorithm: the control flow that decides alleles, MRCAs and counts is executed
from the .class files.  `gen_events_refbytecode_golden.py` uses it to write tests/golden/events_refbytecode.json.
Needs /root/reference; never imported by the product or on the GPU box.
"""
from __future__ import annotations

import os
import struct
import zipfile

from minijvm import NATIVES, ClassFile, JArray, JObject, JLong, JFloat, MiniJVM, parse_desc, _i32, _i64

REF = "/root/reference/compiled"
NODE = "org/gersteinlab/coevolution/core/data/NewickTreeNode"
TREE = "org/gersteinlab/coevolution/core/data/NewickTree"


# ----------------------------------------------------------------------------- JDK value emulation
class JChar:
    """java.lang.Character"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v & 0xFFFF


def jhash(k):
    if k is None:
        return 0
    if isinstance(k, str):                      # String.hashCode
        h = 0
        for ch in k:
            h = (31 * h + ord(ch)) & 0xFFFFFFFF
        return _i32(h)
    if isinstance(k, JChar):
        return k.v
    if isinstance(k, int):                      # Integer
        return _i32(k)
    return id(k) & 0x7FFFFFFF                   # identity hash (never iterated in an order-sensitive way here)


def jequals(a, b):
    if a is b:
        return True
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if isinstance(a, JChar) and isinstance(b, JChar):
        return a.v == b.v
    return False


def _pykey(k):
    if isinstance(k, str):
        return ("s", k)
    if isinstance(k, JChar):
        return ("c", k.v)
    if k is None:
        return ("n",)
    return ("o", id(k))


def _table_size_for(c):
    n = 1
    while n < c:
        n <<= 1
    return max(n, 1)


class JHashMap:
    """java.util.HashMap: same contents AND same iteration order as the JDK's (no tree bins: fine below 8 collisions)."""

    def __init__(self, initial_capacity=None):
        self.cap = 0                                    # table allocated by the first put
        self.init_cap = 16 if initial_capacity is None else _table_size_for(initial_capacity)
        self.e = {}                                     # pykey -> [key, value, spread hash, insertion seq]
        self.seq = 0

    @staticmethod
    def spread(k):
        h = jhash(k) & 0xFFFFFFFF
        return h ^ (h >> 16)

    def put(self, k, v):
        pk = _pykey(k)
        cur = self.e.get(pk)
        if cur is not None:
            old = cur[1]
            cur[1] = v
            return old
        if self.cap == 0:
            self.cap = self.init_cap
        self.seq += 1
        self.e[pk] = [k, v, self.spread(k), self.seq]
        if len(self.e) > 0.75 * self.cap:
            self.cap *= 2                               # resize keeps the relative order inside the split buckets
        return None

    def get(self, k):
        cur = self.e.get(_pykey(k))
        return None if cur is None else cur[1]

    def contains(self, k):
        return _pykey(k) in self.e

    def remove(self, k):
        cur = self.e.pop(_pykey(k), None)
        return None if cur is None else cur[1]

    def clear(self):
        self.e.clear()                                  # the table (capacity) stays

    def size(self):
        return len(self.e)

    def ordered(self):
        cap = self.cap or self.init_cap
        return sorted(self.e.values(), key=lambda t: (t[2] & (cap - 1), t[3]))


class JHashSet:
    def __init__(self, initial_capacity=None):
        self.m = JHashMap(initial_capacity)


class JArrayList:
    def __init__(self, data=None):
        self.d = list(data or [])


class JIterator:
    def __init__(self, items):
        self.items = list(items)
        self.i = 0


class JEntry:
    def __init__(self, cell):
        self.cell = cell


class JView:
    """entrySet() / keySet() / values() of a JHashMap"""
    def __init__(self, m, kind):
        self.m, self.kind = m, kind

    def items(self):
        o = self.m.ordered()
        if self.kind == "entry":
            return [JEntry(c) for c in o]
        return [c[0] if self.kind == "key" else c[1] for c in o]


class JLambda:
    def __init__(self, kind, cname, name, desc, captured):
        self.kind, self.cname, self.name, self.desc, self.captured = kind, cname, name, desc, captured


class JStream:
    def __init__(self, items):
        self.items = items


class JCollectorGroupingBy:
    def __init__(self, fn):
        self.fn = fn


class JDouble:
    """java.lang.Double (boxed: one stack slot, unlike the primitive, which the interpreter keeps as a python float)"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = float(v)


class JBoxedLong:
    """java.lang.Long (boxed)"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v)


class JStringBuilder:
    def __init__(self):
        self.parts = []


class JDoubleStream:
    def __init__(self, items):
        self.items = items


class JOptionalDouble:
    def __init__(self, v):
        self.v = v


class JLatch:
    def __init__(self, n):
        self.n = n


class JOpaque:
    """a JDK object the path only passes around (ZonedDateTime, Duration, ExecutorService, Runtime)"""
    def __init__(self, what):
        self.what = what


class JavaRandom:
    """java.util.Random (the documented LCG) — stands in for the unseeded static Random inside Collections.shuffle(List)"""
    def __init__(self, seed):
        self.seed = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next(self, bits):
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        return _i32(self.seed >> (48 - bits))

    def next_int(self, bound):
        r = self.next(31)
        m = bound - 1
        if (bound & m) == 0:
            return _i32((bound * r) >> 31)
        u = r
        while True:
            r = u % bound
            if _i32(u - r + m) >= 0:
                return r
            u = self.next(31)


class JWriter:
    """BufferedWriter / PrintStream sink: keeps what was written"""
    def __init__(self):
        self.text = []


def _impl(x):
    if isinstance(x, JObject):
        return x.fields.get("$impl", x)
    return x


def jstr(v, desc_char=None):
    """String.valueOf as string concatenation applies it"""
    if v is None:
        return "null"
    if isinstance(v, str):
        return v
    if isinstance(v, JChar):
        return chr(v.v)
    if isinstance(v, JBoxedLong):
        return str(v.v)
    if desc_char == "C":
        return chr(v)
    if desc_char == "Z":
        return "true" if v else "false"
    if isinstance(v, int):
        return str(int(v))
    raise NotImplementedError(("jstr", type(v)))


# ----------------------------------------------------------------------------- invokedynamic metadata
def parse_indy(cf: ClassFile):
    """-> (indy: cp index -> (bootstrap index, name, desc), bootstraps: [(mh_kind, (cls, name, desc), [static args])])"""
    d = cf.data
    pos = 8
    (n,) = struct.unpack_from(">H", d, pos)
    pos += 2
    mh, mt, indy = {}, {}, {}
    i = 1
    while i < n:
        t = d[pos]
        pos += 1
        if t == 1:
            (ln,) = struct.unpack_from(">H", d, pos); pos += 2 + ln
        elif t in (3, 4):
            pos += 4
        elif t in (5, 6):
            pos += 8; i += 1
        elif t in (7, 8, 16):
            if t == 16:
                mt[i] = struct.unpack_from(">H", d, pos)[0]
            pos += 2
        elif t in (9, 10, 11, 12):
            pos += 4
        elif t == 15:
            mh[i] = (d[pos], struct.unpack_from(">H", d, pos + 1)[0]); pos += 3
        elif t == 18:
            indy[i] = struct.unpack_from(">HH", d, pos); pos += 4
        else:
            raise ValueError(t)
        i += 1
    pos += 6
    (n_if,) = struct.unpack_from(">H", d, pos); pos += 2 + 2 * n_if
    for _ in range(2):                                   # fields, then methods
        (cnt,) = struct.unpack_from(">H", d, pos); pos += 2
        for _ in range(cnt):
            na = struct.unpack_from(">HHHH", d, pos)[3]; pos += 8
            for _ in range(na):
                aln = struct.unpack_from(">HI", d, pos)[1]; pos += 6 + aln
    (na,) = struct.unpack_from(">H", d, pos); pos += 2
    boots = []
    for _ in range(na):
        ani, aln = struct.unpack_from(">HI", d, pos); pos += 6
        if cf.utf(ani) == "BootstrapMethods":
            p = pos
            (nb,) = struct.unpack_from(">H", d, p); p += 2
            for _ in range(nb):
                ref, nargs = struct.unpack_from(">HH", d, p); p += 4
                args = list(struct.unpack_from(">" + "H" * nargs, d, p)); p += 2 * nargs
                boots.append((ref, args))
        pos += aln
    out = {}
    for idx, (bi, nti) in indy.items():
        _, ni, di = cf.cp[nti]
        out[idx] = (bi, cf.utf(ni), cf.utf(di))
    return out, boots, mh, mt


# ----------------------------------------------------------------------------- the VM
class RefJVM(MiniJVM):
    def __init__(self):
        self.jars = [zipfile.ZipFile(os.path.join(REF, j)) for j in ("coevolution.jar", "commons-math3-3.6.1.jar")]
        self.jar_of = {}
        for z in self.jars:
            for nm in z.namelist():
                if nm.endswith(".class"):
                    self.jar_of.setdefault(nm, z)
        self.shuffle_rng = None          # JavaRandom: the source of Collections.shuffle(List); set by the driver
        self.shuffles = []               # every shuffle's result as (before, after) python lists
        self.sorted_arrays = []          # (copy before, the JArray) of every Arrays.parallelSort(double[])
        self.classes = {}
        self.initialized = set()
        self.n_bytecodes = 0
        self._indy = {}
        self.console = JWriter()

    # -- class loading: loose .class files of the reference + coevolution.jar
    def has_class(self, name):
        return os.path.exists(os.path.join(REF, name + ".class")) or (name + ".class") in self.jar_of

    def load(self, name):
        c = self.classes.get(name)
        if c is None:
            p = os.path.join(REF, name + ".class")
            data = open(p, "rb").read() if os.path.exists(p) else self.jar_of[name + ".class"].read(name + ".class")
            c = ClassFile(data)
            self.classes[name] = c
        if name not in self.initialized:
            self.initialized.add(name)
            if c.super_name and self.has_class(c.super_name):
                self.load(c.super_name)
            if name == "Homoplasy_Counter":
                # its <clinit> only sets $assertionsDisabled = !Homoplasy_Counter.class.desiredAssertionStatus(): a JVM started
                # without -ea (poutine.sh does not pass it) runs with assertions disabled
                c.static_fields["$assertionsDisabled"] = 1
            else:
                m = c.methods.get(("<clinit>", "()V"))
                if m is not None:
                    self.run(m, [])
        return c

    # -- invokedynamic
    def invokedynamic(self, cf, cp_index, st):
        key = id(cf)
        if key not in self._indy:
            self._indy[key] = parse_indy(cf)
        sites, boots, mh, mt = self._indy[key]
        bi, name, desc = sites[cp_index]
        ref, bargs = boots[bi]
        kind, mref = mh[ref]
        bcls, bname, _ = cf.memberref(mref)
        nargs = len(parse_desc(desc))
        args = st[len(st) - nargs:] if nargs else []
        if nargs:
            del st[len(st) - nargs:]
        if bcls == "java/lang/invoke/LambdaMetafactory" and bname == "metafactory":
            ikind, iref = mh[bargs[1]]                      # implMethod
            icls, iname, idesc = cf.memberref(iref)
            st.append(JLambda(ikind, icls, iname, idesc, list(args)))
        elif bcls == "java/lang/invoke/StringConcatFactory" and bname == "makeConcatWithConstants":
            recipe = cf.const(bargs[0])
            consts = [cf.const(b) for b in bargs[1:]]
            types = _arg_types(desc)
            out, ai, ci = [], 0, 0
            for ch in recipe:
                if ch == "\x01":
                    out.append(jstr(args[ai], types[ai])); ai += 1
                elif ch == "\x02":
                    out.append(jstr(consts[ci])); ci += 1
                else:
                    out.append(ch)
            st.append("".join(out))
        else:
            raise NotImplementedError(("bootstrap", bcls, bname))

    def call_lambda(self, lam: JLambda, args):
        a = list(lam.captured) + list(args)
        if lam.kind == 6:                                    # REF_invokeStatic
            return self.invoke("static", lam.cname, lam.name, lam.desc, a)
        if lam.kind in (5, 7, 9):                            # invokeVirtual / invokeSpecial / invokeInterface
            return self.invoke("virtual", lam.cname, lam.name, lam.desc, a)
        raise NotImplementedError(("lambda kind", lam.kind))

    # -- method dispatch: reference bytecode first, JDK emulation for everything else
    def invoke(self, kind, cname, name, desc, args):
        if name == "get_thread_pool_status" and cname == "Homoplasy_Counter":
            return "[thread pool emulated: replicates run serially in submission order]"   # console decoration only (HC.java:3188-3227)
        if args and isinstance(args[0], JArray) and name == "clone":
            return JArray(list(args[0].data), args[0].kind)
        if name == "ordinal" and desc == "()I" and args and isinstance(args[0], JObject) and "$ordinal" in args[0].fields:
            return args[0].fields["$ordinal"]
        if kind in ("virtual", "interface", "special") and args and isinstance(args[0], JObject) and self.has_class(args[0].cls) \
                and "$impl" not in args[0].fields:
            start = args[0].cls if kind != "special" else cname
            m = self.find_method(start, name, desc)
            if m is not None and m.code is not None:
                return self.run(m, args)
            if name == "equals":
                return int(args[0] is args[1])
            if name == "hashCode":
                return id(args[0]) & 0x7FFFFFFF
            if name == "<init>" and cname == "java/lang/Object":
                return None
        if kind == "static" and self.has_class(cname):
            m = self.find_method(cname, name, desc)
            if m is not None and m.code is not None:
                return self.run(m, args)
        return self.jdk(kind, cname, name, desc, args)

    def jdk(self, kind, cname, name, desc, args):
        nat = NATIVES.get((cname, name, desc))           # java.lang.Double / Math statics of the numeric chain
        if nat is not None:
            return nat(args)
        # ---- constructors: the `new` opcode made an empty JObject shell; give it its emulation object
        if name == "<init>":
            o = args[0]
            if cname == "java/lang/Object":
                return None
            if cname == "java/lang/Enum":                # Enum(String name, int ordinal)
                o.fields["$name"], o.fields["$ordinal"] = args[1], args[2]
                return None
            if cname == "java/util/concurrent/ConcurrentHashMap" and desc == "()V":
                o.fields["$impl"] = JHashMap()           # no order-sensitive use: get / containsKey / put with String keys
                return None
            if cname == "java/util/concurrent/CountDownLatch":
                o.fields["$impl"] = JLatch(args[1])
                return None
            if cname == "java/lang/StringBuilder" and desc == "()V":
                o.fields["$impl"] = JStringBuilder()
                return None
            if cname in ("java/util/HashMap", "java/util/LinkedHashMap"):
                if cname != "java/util/HashMap":
                    raise NotImplementedError(cname)
                if desc == "()V":
                    o.fields["$impl"] = JHashMap()
                elif desc == "(I)V":
                    o.fields["$impl"] = JHashMap(args[1])
                else:
                    raise NotImplementedError((cname, desc))
                return None
            if cname == "java/util/HashSet":
                if desc == "()V":
                    o.fields["$impl"] = JHashSet()
                else:
                    raise NotImplementedError((cname, desc))
                return None
            if cname == "java/util/ArrayList":
                if desc == "()V":
                    o.fields["$impl"] = JArrayList()
                elif desc == "(Ljava/util/Collection;)V":
                    o.fields["$impl"] = JArrayList(self._iterate(args[1]))
                elif desc == "(I)V":
                    o.fields["$impl"] = JArrayList()
                else:
                    raise NotImplementedError((cname, desc))
                return None
            raise NotImplementedError(("new", cname, desc))
        # ---- statics
        if kind == "static":
            if cname == "java/lang/Character" and name == "valueOf":
                return JChar(args[0])
            if cname == "java/lang/Integer" and name == "valueOf":
                return int(args[0])
            if cname == "java/lang/Double" and name == "valueOf" and desc == "(D)Ljava/lang/Double;":
                return JDouble(args[0])
            if cname == "java/lang/Long" and name == "valueOf" and desc == "(J)Ljava/lang/Long;":
                return JBoxedLong(args[0])
            if cname == "java/lang/Runtime" and name == "getRuntime":
                return JOpaque("runtime")
            if cname == "java/util/concurrent/Executors" and name == "newFixedThreadPool":
                return JOpaque("executor")
            if cname == "java/time/ZonedDateTime" and name == "now":
                return JOpaque("zoned-date-time")
            if cname == "java/time/Duration" and name == "between":
                return JOpaque("duration")
            if cname == "java/lang/Math" and name == "toIntExact":
                if not -(1 << 31) <= int(args[0]) < (1 << 31):
                    raise RuntimeError("Math.toIntExact overflow")
                return int(args[0])
            if cname == "java/util/Arrays" and name == "parallelSort" and desc == "([D)V":
                self.sorted_arrays.append((list(args[0].data), args[0]))
                args[0].data.sort()                      # p-values in [0, 1]: no NaN / -0.0 ordering subtleties
                return None
            if cname == "java/util/Arrays" and name == "stream" and desc == "([D)Ljava/util/stream/DoubleStream;":
                return JDoubleStream(list(args[0].data))
            if cname == "java/util/Collections" and name == "shuffle" and desc == "(Ljava/util/List;)V":
                lst = _impl(args[0])
                before = list(lst.d)
                rnd = self.shuffle_rng
                for i in range(len(lst.d), 1, -1):       # JDK: for (int i = size; i > 1; i--) swap(list, i - 1, rnd.nextInt(i));
                    j = rnd.next_int(i)
                    lst.d[i - 1], lst.d[j] = lst.d[j], lst.d[i - 1]
                self.shuffles.append((before, list(lst.d)))
                return None
            if cname == "java/util/stream/Collectors" and name == "groupingBy" and len(args) == 1:
                return JCollectorGroupingBy(args[0])
            if cname == "java/lang/String" and name == "format":
                return _java_format(args[0], _impl(args[1]).data if args[1] is not None else [])
            if cname == "java/lang/System" and name == "exit":
                raise RuntimeError("the reference called System.exit(%d)" % args[0])
            raise NotImplementedError(("static", cname, name, desc))
        # ---- instance methods, by the receiver's emulation type
        recv = _impl(args[0])
        a = args[1:]
        if recv is None:
            # console: System.out and picocli's Ansi.AUTO are opaque statics (null here)
            if cname == "picocli/CommandLine$Help$Ansi" and name == "string":
                return a[0]
            if cname == "java/io/PrintStream" and name in ("printf", "println", "print"):
                self.console.text.append(a[0] if a else "")
                return None
            raise NotImplementedError(("null receiver", cname, name, desc))
        if isinstance(recv, JOpaque):
            if recv.what == "runtime" and name == "availableProcessors":
                return 1
            if recv.what == "executor":
                if name == "submit":                     # the Runnable runs here and now: serial, in submission order
                    self.invoke("interface", "java/lang/Runnable", "run", "()V", [a[0]])
                    return None
                if name in ("shutdown",):
                    return None
                if name == "toString":
                    return "[emulated executor]"
            if recv.what == "duration":
                if name == "toHours":
                    return JLong(0)
                if name in ("toMinutesPart", "toSecondsPart"):
                    return 0
            if name == "toString":
                return "[" + recv.what + "]"
        if isinstance(recv, JLatch):
            if name == "countDown":
                recv.n = max(0, recv.n - 1)
                return None
            if name == "await":
                if recv.n != 0:
                    raise RuntimeError("CountDownLatch.await with %d replicates outstanding" % recv.n)
                return None
            if name == "getCount":
                return JLong(recv.n)
            if name == "toString":
                return "[latch %d]" % recv.n
        if isinstance(recv, JStringBuilder):
            if name == "append":
                t = _arg_types(desc)[0]
                recv.parts.append(jstr(a[0], t))
                return args[0]
            if name == "toString":
                return "".join(recv.parts)
        if isinstance(recv, JDouble):
            if name == "doubleValue":
                return recv.v
        if isinstance(recv, JDoubleStream):
            if name == "filter":
                return JDoubleStream([x for x in recv.items if self.call_lambda(a[0], [x])])
            if name == "count":
                return JLong(len(recv.items))
            if name == "min":
                return JOptionalDouble(min(recv.items) if recv.items else None)
        if isinstance(recv, JOptionalDouble):
            if name == "getAsDouble":
                if recv.v is None:
                    raise RuntimeError("NoSuchElementException: a replicate with no allele in use")
                return recv.v
        if isinstance(recv, JChar):
            if name == "charValue":
                return recv.v
            if name == "equals":
                return int(jequals(recv, a[0]))
            if name == "hashCode":
                return recv.v
        if isinstance(recv, str):
            if name == "equals":
                return int(jequals(recv, a[0]))
            if name == "hashCode":
                return jhash(recv)
            if name == "length":
                return len(recv)
            if name == "charAt":
                return ord(recv[a[0]])
        if isinstance(recv, JHashMap):
            if name == "put":
                return recv.put(a[0], a[1])
            if name == "get":
                return recv.get(a[0])
            if name == "containsKey":
                return int(recv.contains(a[0]))
            if name == "size":
                return recv.size()
            if name == "isEmpty":
                return int(recv.size() == 0)
            if name == "clear":
                return recv.clear()
            if name == "remove":
                return recv.remove(a[0])
            if name == "entrySet":
                return JView(recv, "entry")
            if name == "keySet":
                return JView(recv, "key")
            if name == "values":
                return JView(recv, "value")
        if isinstance(recv, JHashSet):
            if name == "add":
                had = recv.m.contains(a[0])
                recv.m.put(a[0], True)
                return int(not had)
            if name == "contains":
                return int(recv.m.contains(a[0]))
            if name == "size":
                return recv.m.size()
            if name == "iterator":
                return JIterator(c[0] for c in recv.m.ordered())
        if isinstance(recv, JView):
            if name == "iterator":
                return JIterator(recv.items())
            if name == "size":
                return recv.m.size()
        if isinstance(recv, JArrayList):
            if name == "add" and len(a) == 1:
                recv.d.append(a[0])
                return 1
            if name == "get":
                return recv.d[a[0]]
            if name == "size":
                return len(recv.d)
            if name == "isEmpty":
                return int(not recv.d)
            if name == "iterator":
                return JIterator(recv.d)
            if name == "stream":
                return JStream(list(recv.d))
            if name == "addAll":
                add = self._iterate(a[0])
                recv.d.extend(add)
                return int(bool(add))
            if name == "forEach":
                for x in list(recv.d):
                    self.call_lambda(a[0], [x])
                return None
        if isinstance(recv, JIterator):
            if name == "hasNext":
                return int(recv.i < len(recv.items))
            if name == "next":
                v = recv.items[recv.i]
                recv.i += 1
                return v
        if isinstance(recv, JEntry):
            if name == "getKey":
                return recv.cell[0]
            if name == "getValue":
                return recv.cell[1]
        if isinstance(recv, JStream):
            if name == "filter":
                return JStream([x for x in recv.items if self.call_lambda(a[0], [x])])
            if name == "mapToDouble":
                return JDoubleStream([self.call_lambda(a[0], [x]) for x in recv.items])
            if name == "collect" and isinstance(a[0], JCollectorGroupingBy):
                # Collectors.groupingBy(f) = groupingBy(f, HashMap::new, toList()): a java.util.HashMap of ArrayLists,
                # filled in encounter order
                out = JHashMap()
                for x in recv.items:
                    k = self.call_lambda(a[0].fn, [x])
                    lst = out.get(k)
                    if lst is None:
                        lst = JArrayList()
                        out.put(k, lst)
                    lst.d.append(x)
                return out
        if isinstance(recv, JWriter):
            if name == "write":
                recv.text.append(a[0])
                return None
            if name in ("flush", "close", "newLine"):
                return None
        raise NotImplementedError(("jdk", type(recv).__name__, cname, name, desc))

    def _iterate(self, coll):
        c = _impl(coll)
        if isinstance(c, JArrayList):
            return list(c.d)
        if isinstance(c, JHashSet):
            return [x[0] for x in c.m.ordered()]
        if isinstance(c, JView):
            return c.items()
        raise NotImplementedError(type(c))

    # -- helpers for the driver script
    def new(self, cname, desc="()V", *args):
        self.load(cname) if self.has_class(cname) else None
        o = JObject(cname)
        self.invoke("special", cname, "<init>", desc, [o] + list(args))
        return o

    def call(self, obj, name, desc, *args):
        return self.invoke("virtual", obj.cls, name, desc, [obj] + list(args))


def _arg_types(desc):
    out, i = [], 1
    while desc[i] != ")":
        c = desc[i]
        if c == "L":
            j = desc.index(";", i)
            out.append("L"); i = j + 1
        elif c == "[":
            while desc[i] == "[":
                i += 1
            i = desc.index(";", i) + 1 if desc[i] == "L" else i + 1
            out.append("[")
        else:
            out.append(c); i += 1
    return out


def _java_format(fmt, args):
    """String.format for the few patterns the path uses (%n, %d, %s)"""
    out, ai, i = [], 0, 0
    while i < len(fmt):
        ch = fmt[i]
        if ch == "%" and i + 1 < len(fmt):
            c = fmt[i + 1]
            if c == "n":
                out.append("\n")
            elif c == "%":
                out.append("%")
            elif c in "ds":
                out.append(jstr(args[ai])); ai += 1
            elif fmt[i + 1:i + 4] == "02d":
                out.append("%02d" % int(args[ai].v if isinstance(args[ai], JBoxedLong) else args[ai])); ai += 1
                i += 2
            else:
                raise NotImplementedError("format %" + c)
            i += 2
        else:
            out.append(ch); i += 1
    return "".join(out)


# ----------------------------------------------------------------------------- the reference entry point
class ReferenceEvents:
    """count_all_homoplasy_events(tree, seg_sites, physical_poss) — HC.java:1402 — run from the reference's bytecode."""

    def __init__(self):
        self.vm = RefJVM()

    def build_tree(self, parent, names, children_order=None):
        """NewickTreeNode objects wired through the class's own setters, children in the given order, then
        `new NewickTree(root)` (its buildCaches makes the leaf list in DFS order)."""
        vm = self.vm
        n = len(parent)
        nodes = [vm.new(NODE) for _ in range(n)]
        for v in range(n):
            vm.call(nodes[v], "setId", "(Ljava/lang/String;)V", names[v])
        kids = {v: [] for v in range(n)}
        root = None
        for v in range(n):
            if parent[v] < 0:
                root = v
            else:
                kids[parent[v]].append(v)
        if children_order is not None:
            kids = children_order
        for u in range(n):
            for c in kids[u]:
                vm.call(nodes[c], "setParent", "(L" + NODE + ";)V", nodes[u])
                vm.call(nodes[u], "addChild", "(L" + NODE + ";)V", nodes[c])
        tree = vm.new(TREE, "(L" + NODE + ";)V", nodes[root])
        return tree

    def run(self, parent, names, chars, physical_poss=None):
        """chars: N x S array of character codes (row v = node v).  -> (list of per-biallelic-site dicts, log text)"""
        vm = self.vm
        n, S = len(parent), len(chars[0])
        tree = self.build_tree(parent, names)
        seg = vm.new("java/util/HashMap")
        for v in range(n):
            vm.call(seg, "put", "(Ljava/lang/Object;Ljava/lang/Object;)Ljava/lang/Object;", names[v], JArray([int(c) for c in chars[v]], "C"))
        pos = JArray(list(physical_poss) if physical_poss is not None else list(range(101, 101 + S)), "I")
        hc = JObject("Homoplasy_Counter")
        oo = JObject("Homoplasy_Counter$OutputOptions")
        log = JWriter()
        oo.fields["log"] = log
        oo.fields["debug"] = JWriter()
        hc.fields["outputOptions"] = oo
        hc.fields["DEBUG_MODE"] = 0
        vm.load("Homoplasy_Counter")
        m = vm.find_method("Homoplasy_Counter", "count_all_homoplasy_events",
                           "(L" + TREE + ";Ljava/util/HashMap;[I)Ljava/util/ArrayList;")
        res = vm.run(m, [hc, tree, seg, pos])
        self.hc, self.all_events, self.log = hc, res, log
        out = []
        for he in _impl(res).d:
            f = he.fields
            a1 = sorted(c[0] for c in _impl(f["allele1_MRCAs"]).ordered())
            a2 = sorted(c[0] for c in _impl(f["allele2_MRCAs"]).ordered())
            out.append({"segsite_id": f["segsite_ID"], "physical_pos": f["physical_pos"], "allele1": chr(f["allele1"]), "allele2": chr(f["allele2"]),
                        "a1_count": f["a1_count"], "a2_count": f["a2_count"],
                        "a1_extant": f.get("a1_count_extant_only", 0), "a2_extant": f.get("a2_count_extant_only", 0),
                        "a1_mrcas": a1, "a2_mrcas": a2})
        return out, "".join(log.text), "".join(vm.console.text)


    # ---- entry point 2: assoc(all_events, phenos) -> HC.java:2029 -> 2096 -> subset_by_min_hcount + the live resampling driver
    def assoc(self, sample_names, labels, m, min_hcount, shuffle_seed):
        """Runs, from the reference's bytecode and on the Homoplasy_Events objects the run() above left in the VM:
        subset_by_min_hcount (HC.java:3783) and resample_all_mutations_binom_test_combined_nulldists_memoization_concurrent
        (HC.java:3013: observed counts and binomial p-values — commons-math 3.6.1's own bytecode —, m Replicate.run() calls,
        point-wise and family-wise estimates).  Two things are supplied from outside, because the JDK leaves them open:
        the thread pool runs the replicates one after the other in submission order (the reference's `r++` races make
        the serial result the only defined one), and Collections.shuffle(List) draws from a java.util.Random seeded with
        `shuffle_seed` instead of an unseeded one.  Every shuffle is recorded, so the same label arrangements can be
        replayed through the oracle and the GPU.
        labels: the pheno-file strings ("1", "0", anything else = missing), one per sample, in file order.
        -> dict(p_success, sites=[...], maxT_unsorted, maxT_sorted, perms (m x P indices into `labels`), log)"""
        vm, hc = self.vm, self.hc
        phenos = vm.new("java/util/HashMap")
        for nm, lab in zip(sample_names, labels):
            vm.call(phenos, "put", "(Ljava/lang/Object;Ljava/lang/Object;)Ljava/lang/Object;", nm, lab)
        # what read_phenos (HC.java:1999-2007) leaves behind
        hc.fields["tot_num_sample_cases"] = sum(1 for x in labels if x == "1")
        hc.fields["tot_num_sample_controls"] = sum(1 for x in labels if x == "0")
        # instance initialisers of Homoplasy_Counter (HC.java:167-181): the live configuration
        hc.fields["EXTANT_NODES_ONLY"] = 1
        hc.fields["m"] = int(m)
        alt = "org/apache/commons/math3/stat/inference/AlternativeHypothesis"
        vm.load(alt)
        hc.fields["TAIL_TYPE"] = vm.find_static_holder(alt, "GREATER_THAN").static_fields["GREATER_THAN"]
        vm.load("Homoplasy_Counter$AlgoParams").static_fields["min_hcount"] = int(min_hcount)
        vm.load("Homoplasy_Counter$RuntimeSettings").static_fields["num_threads"] = 1
        vm.shuffle_rng = JavaRandom(shuffle_seed)
        vm.shuffles, vm.sorted_arrays = [], []
        sub = vm.call(hc, "subset_by_min_hcount", "(Ljava/util/ArrayList;I)Ljava/util/ArrayList;", self.all_events, int(min_hcount))
        stats = vm.call(hc, "resample_all_mutations_binom_test_combined_nulldists_memoization_concurrent",
                        "(Ljava/util/ArrayList;Ljava/util/HashMap;)[LHomoplasy_Counter$Binomial_Test_Stat;", sub, phenos)
        p_success = vm.call(hc, "calc_background_p_success", "()D")
        sites = []
        for he, st in zip(_impl(sub).d, stats.data):
            f, g = he.fields, st.fields
            row = {"segsite_id": f["segsite_ID"], "a1_in_use": int(f.get("a1_in_use", 0)), "a2_in_use": int(f.get("a2_in_use", 0)),
                   "obs_counts": [int(x) for x in f["obs_counts"].data]}
            for k in ("r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
                row[k] = int(g.get(k, 0))
            for k in ("obs_binom_pvalue_a1", "obs_binom_pvalue_a2", "pointwise_pvalue_a1", "pointwise_pvalue_a2", "familywise_pvalue_a1",
                      "familywise_pvalue_a2"):
                row[k] = float(g[k])
            sites.append(row)
        assert len(vm.sorted_arrays) == 1 and len(vm.shuffles) == m
        unsorted, arr = vm.sorted_arrays[0]
        # the label arrangement of every replicate as a permutation of the pheno-file rows.  permute_phenos (HC.java:4890)
        # lists the samples in phenos.entrySet() order, shuffles the label list and pairs them up again by position.
        order = [c[0] for c in _impl(phenos).ordered()]          # sample names in HashMap iteration order
        row_of = {nm: j for j, nm in enumerate(sample_names)}
        perms = []
        for before, after in vm.shuffles:
            assert before == [labels[row_of[nm]] for nm in order]
            pools = {}
            for j, lab in enumerate(labels):                     # pheno-file rows carrying each label string, in file order
                pools.setdefault(lab, []).append(j)
            perm = [0] * len(labels)
            for nm, lab in zip(order, after):                    # sample nm ends up with label string `lab`
                perm[row_of[nm]] = pools[lab].pop(0)
            perms.append(perm)
        return {"p_success": p_success, "sites": sites, "maxT_unsorted": unsorted, "maxT_sorted": list(arr.data), "perms": perms,
                "log": "".join(self.log.text)}
