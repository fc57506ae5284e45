"""Minimal JVM bytecode interpreter — TEST INFRASTRUCTURE ONLY.

Purpose: pin the oracle's binomial p-value chain against the *actual bytecode* of the
reference's un-sourced arithmetic dependency (Apache commons-math3 3.6.1, shipped only as
`/root/reference/compiled/commons-math3-3.6.1.jar`).  There is no JVM in the authoring
container, so this interpreter executes the handful of pure-numeric classes the hot path
reaches (`BinomialTest.binomialTest` -> `BinomialDistribution.cumulativeProbability` ->
`Beta.regularizedBeta` -> `ContinuedFraction.evaluate`, `Gamma.*`, `FastMath.exp/log/log1p`)
directly from the jar and emits golden vectors (see `gen_binom_golden.py`).

It is NOT a general JVM: no threads, no exceptions tables, no invokedynamic (refjvm.py adds the
two bootstrap kinds javac emits), no JDK class library beyond a few `java.lang.Double/Math` natives.  Objects of classes outside the jar
are opaque.  Python `float` is IEEE-754 binary64, Java's `double` arithmetic (strict, no
FMA contraction) maps 1:1 onto it.

Nothing in the product path or on the GPU box uses this file; it needs /root/reference.
"""
from __future__ import annotations

import math
import struct
import zipfile


# ----------------------------------------------------------------------------- values
class JLong(int):
    """Category-2 integer (Java long)."""
    __slots__ = ()


class JFloat(float):
    """Category-1 float (Java float, value already rounded to binary32)."""
    __slots__ = ()


class JObject:
    __slots__ = ("cls", "fields")

    def __init__(self, cls):
        self.cls = cls
        self.fields = {}


class JArray:
    __slots__ = ("data", "kind")

    def __init__(self, data, kind):
        self.data = data
        self.kind = kind


class JavaThrow(Exception):
    def __init__(self, obj):
        super().__init__(getattr(obj, "cls", obj))
        self.obj = obj


def _i32(v):
    v &= 0xFFFFFFFF
    return v - 0x100000000 if v & 0x80000000 else v


def _i64(v):
    v &= 0xFFFFFFFFFFFFFFFF
    return v - 0x10000000000000000 if v & 0x8000000000000000 else v


def _f32(v):
    try:
        return JFloat(struct.unpack("<f", struct.pack("<f", v))[0])
    except OverflowError:
        return JFloat(math.copysign(math.inf, v))


def _is_cat2(v):
    return isinstance(v, JLong) or (isinstance(v, float) and not isinstance(v, JFloat))


def _d2i(v, bits):
    if v != v:
        return 0
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    if v >= hi:
        return hi
    if v <= lo:
        return lo
    return int(v)


def _ddiv(a, b):
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf
    try:
        return a / b
    except OverflowError:
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def _dmul(a, b):
    try:
        return a * b
    except OverflowError:
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


# ----------------------------------------------------------------------------- class file
class Method:
    __slots__ = ("name", "desc", "flags", "code", "max_locals", "nargs", "cls", "argcats")


class ClassFile:
    def __init__(self, data: bytes):
        self.data = data
        pos = 8
        (n,) = struct.unpack_from(">H", data, pos)
        pos += 2
        cp = [None] * n
        i = 1
        while i < n:
            t = data[pos]
            pos += 1
            if t == 1:
                (ln,) = struct.unpack_from(">H", data, pos)
                cp[i] = ("utf8", data[pos + 2:pos + 2 + ln].decode("utf-8", "replace"))
                pos += 2 + ln
            elif t == 3:
                cp[i] = ("int", struct.unpack_from(">i", data, pos)[0]); pos += 4
            elif t == 4:
                cp[i] = ("float", struct.unpack_from(">f", data, pos)[0]); pos += 4
            elif t == 5:
                cp[i] = ("long", struct.unpack_from(">q", data, pos)[0]); pos += 8; i += 1
            elif t == 6:
                cp[i] = ("double", struct.unpack_from(">d", data, pos)[0]); pos += 8; i += 1
            elif t == 7:
                cp[i] = ("class", struct.unpack_from(">H", data, pos)[0]); pos += 2
            elif t == 8:
                cp[i] = ("string", struct.unpack_from(">H", data, pos)[0]); pos += 2
            elif t in (9, 10, 11):
                cp[i] = ({9: "field", 10: "method", 11: "imethod"}[t],) + struct.unpack_from(">HH", data, pos); pos += 4
            elif t == 12:
                cp[i] = ("nat",) + struct.unpack_from(">HH", data, pos); pos += 4
            elif t == 15:
                cp[i] = ("mh",); pos += 3
            elif t == 16:
                cp[i] = ("mt",); pos += 2
            elif t == 18:
                cp[i] = ("indy",); pos += 4
            else:
                raise ValueError(f"constant pool tag {t}")
            i += 1
        self.cp = cp
        self.access, this_c, super_c, n_if = struct.unpack_from(">HHHH", data, pos)
        pos += 8 + 2 * n_if
        self.name = self.utf(cp[this_c][1])
        self.super_name = self.utf(cp[super_c][1]) if super_c else None
        self.static_fields = {}
        self.field_descs = {}
        self.methods = {}
        (nf,) = struct.unpack_from(">H", data, pos)
        pos += 2
        for _ in range(nf):
            acc, ni, di, na = struct.unpack_from(">HHHH", data, pos)
            pos += 8
            fname, fdesc = self.utf(ni), self.utf(di)
            self.field_descs[fname] = (fdesc, acc)
            cv = None
            for _ in range(na):
                ani, aln = struct.unpack_from(">HI", data, pos)
                pos += 6
                if self.utf(ani) == "ConstantValue":
                    cv = self.const(struct.unpack_from(">H", data, pos)[0])
                pos += aln
            if acc & 0x0008:
                if cv is None:
                    cv = {"D": 0.0, "F": JFloat(0.0), "J": JLong(0)}.get(fdesc, 0 if fdesc[0] in "IZBCS" else None)
                self.static_fields[fname] = cv
        (nm,) = struct.unpack_from(">H", data, pos)
        pos += 2
        for _ in range(nm):
            acc, ni, di, na = struct.unpack_from(">HHHH", data, pos)
            pos += 8
            m = Method()
            m.name, m.desc, m.flags, m.code, m.cls = self.utf(ni), self.utf(di), acc, None, self
            m.max_locals = 0
            for _ in range(na):
                ani, aln = struct.unpack_from(">HI", data, pos)
                pos += 6
                if self.utf(ani) == "Code":
                    _ms, ml, cl = struct.unpack_from(">HHI", data, pos)
                    m.code = data[pos + 8:pos + 8 + cl]
                    m.max_locals = ml
                pos += aln
            m.argcats = parse_desc(m.desc)
            m.nargs = len(m.argcats)
            self.methods[(m.name, m.desc)] = m

    def utf(self, i):
        return self.cp[i][1]

    def const(self, i):
        e = self.cp[i]
        k = e[0]
        if k == "int":
            return e[1]
        if k == "float":
            return JFloat(e[1])
        if k == "long":
            return JLong(e[1])
        if k == "double":
            return e[1]
        if k == "string":
            return self.utf(e[1])
        if k == "class":
            return ("classref", self.utf(e[1]))
        raise ValueError(k)

    def memberref(self, i):
        _, ci, nti = self.cp[i]
        cname = self.utf(self.cp[ci][1])
        _, ni, di = self.cp[nti]
        return cname, self.utf(ni), self.utf(di)


def parse_desc(desc):
    """-> list of arg categories (1 or 2)."""
    out = []
    i = 1
    while desc[i] != ")":
        c = desc[i]
        if c in "DJ":
            out.append(2); i += 1
        elif c == "L":
            out.append(1); i = desc.index(";", i) + 1
        elif c == "[":
            while desc[i] == "[":
                i += 1
            if desc[i] == "L":
                i = desc.index(";", i) + 1
            else:
                i += 1
            out.append(1)
        else:
            out.append(1); i += 1
    return out


# ----------------------------------------------------------------------------- natives
def _raw_bits(d):
    return JLong(struct.unpack("<q", struct.pack("<d", d))[0])


def _bits_double(b):
    return struct.unpack("<d", struct.pack("<q", _i64(b)))[0]


NATIVES = {
    ("java/lang/Double", "doubleToRawLongBits", "(D)J"): lambda a: _raw_bits(a[0]),
    ("java/lang/Double", "doubleToLongBits", "(D)J"): lambda a: _raw_bits(math.nan if a[0] != a[0] else a[0]),
    ("java/lang/Double", "longBitsToDouble", "(J)D"): lambda a: _bits_double(a[0]),
    ("java/lang/Double", "isNaN", "(D)Z"): lambda a: int(a[0] != a[0]),
    ("java/lang/Double", "isInfinite", "(D)Z"): lambda a: int(a[0] in (math.inf, -math.inf)),
    ("java/lang/Float", "floatToRawIntBits", "(F)I"): lambda a: struct.unpack("<i", struct.pack("<f", a[0]))[0],
    ("java/lang/Float", "intBitsToFloat", "(I)F"): lambda a: JFloat(struct.unpack("<f", struct.pack("<i", a[0]))[0]),
    ("java/lang/Math", "sqrt", "(D)D"): lambda a: math.sqrt(a[0]) if a[0] >= 0 else math.nan,
    ("java/lang/Math", "abs", "(D)D"): lambda a: abs(a[0]),
    ("java/lang/Math", "abs", "(I)I"): lambda a: _i32(abs(a[0])),
    ("java/lang/Math", "max", "(II)I"): lambda a: max(a[0], a[1]),
    ("java/lang/Math", "min", "(II)I"): lambda a: min(a[0], a[1]),
    ("java/lang/Math", "floor", "(D)D"): lambda a: float(math.floor(a[0])) if math.isfinite(a[0]) else a[0],
    # FastMath.<clinit> only: LOG_MAX_VALUE = StrictMath.log(Double.MAX_VALUE) (used by cosh/sinh, not on our path)
    ("java/lang/StrictMath", "log", "(D)D"): lambda a: math.log(a[0]),
    ("java/lang/Object", "<init>", "()V"): lambda a: None,
    ("java/lang/Object", "getClass", "()Ljava/lang/Class;"): lambda a: ("classref", a[0].cls),
}


# ----------------------------------------------------------------------------- VM
class MiniJVM:
    def __init__(self, jar_path):
        self.zip = zipfile.ZipFile(jar_path)
        self.names = set(self.zip.namelist())
        self.classes = {}
        self.initialized = set()
        self.n_bytecodes = 0

    def has_class(self, name):
        return name + ".class" in self.names

    def load(self, name) -> ClassFile:
        c = self.classes.get(name)
        if c is None:
            c = ClassFile(self.zip.read(name + ".class"))
            self.classes[name] = c
        if name not in self.initialized:
            self.initialized.add(name)
            if c.super_name and self.has_class(c.super_name):
                self.load(c.super_name)
            m = c.methods.get(("<clinit>", "()V"))
            if m is not None:
                self.run(m, [])
        return c

    def find_method(self, cname, name, desc):
        while cname and self.has_class(cname):
            c = self.load(cname)
            m = c.methods.get((name, desc))
            if m is not None:
                return m
            cname = c.super_name
        return None

    def find_static_holder(self, cname, fname):
        while cname and self.has_class(cname):
            c = self.load(cname)
            if fname in c.static_fields:
                return c
            cname = c.super_name
        raise KeyError(fname)

    def call_static(self, cname, name, desc, *args):
        m = self.find_method(cname, name, desc)
        if m is None:
            raise KeyError((cname, name, desc))
        return self.run(m, list(args))

    def invoke(self, kind, cname, name, desc, args):
        """args includes receiver for non-static. Returns value or None."""
        nat = NATIVES.get((cname, name, desc))
        if nat is not None:
            return nat(args)
        if cname == "java/lang/Enum" and name == "<init>":  # Enum(String name, int ordinal)
            args[0].fields["$name"], args[0].fields["$ordinal"] = args[1], args[2]
            return None
        if name == "ordinal" and desc == "()I" and isinstance(args[0], JObject) and "$ordinal" in args[0].fields:
            return args[0].fields["$ordinal"]
        if kind == "virtual" or kind == "interface":
            recv = args[0]
            if isinstance(recv, JArray):
                if name == "clone":
                    return JArray(list(recv.data), recv.kind)
                raise NotImplementedError(("array", name))
            if recv is None:
                raise JavaThrow("java/lang/NullPointerException")
            m = self.find_method(recv.cls, name, desc)
        else:
            m = self.find_method(cname, name, desc) if self.has_class(cname) else None
        if m is None:
            if name == "<init>" and not self.has_class(cname):
                return None  # opaque JDK object constructor
            raise NotImplementedError(f"no method {cname}.{name}{desc}")
        if m.code is None:
            raise NotImplementedError(f"abstract/native {cname}.{name}{desc}")
        return self.run(m, args)

    def invokedynamic(self, cf, cp_index, st):
        raise NotImplementedError("invokedynamic")   # refjvm.RefJVM implements LambdaMetafactory / StringConcatFactory sites

    # ------------------------------------------------------------------ interpreter loop
    def run(self, m: Method, args):
        code = m.code
        cf = m.cls
        cp = cf.cp
        loc = [None] * (m.max_locals + 2)
        i = 0
        for a in args:
            loc[i] = a
            i += 2 if _is_cat2(a) else 1
        st = []
        push, pop = st.append, st.pop
        pc = 0
        u1 = lambda p: code[p]
        s1 = lambda p: code[p] - 256 if code[p] > 127 else code[p]
        u2 = lambda p: (code[p] << 8) | code[p + 1]
        s2 = lambda p: _s16((code[p] << 8) | code[p + 1])
        s4 = lambda p: struct.unpack_from(">i", code, p)[0]
        nb = 0
        while True:
            op = code[pc]
            nb += 1
            # ---- constants
            if op == 0x00:
                pc += 1
            elif op == 0x01:
                push(None); pc += 1
            elif 0x02 <= op <= 0x08:
                push(op - 3); pc += 1
            elif op in (0x09, 0x0A):
                push(JLong(op - 9)); pc += 1
            elif 0x0B <= op <= 0x0D:
                push(JFloat(op - 0x0B)); pc += 1
            elif op in (0x0E, 0x0F):
                push(float(op - 0x0E)); pc += 1
            elif op == 0x10:
                push(s1(pc + 1)); pc += 2
            elif op == 0x11:
                push(s2(pc + 1)); pc += 3
            elif op == 0x12:
                push(cf.const(u1(pc + 1))); pc += 2
            elif op in (0x13, 0x14):
                push(cf.const(u2(pc + 1))); pc += 3
            # ---- loads
            elif 0x15 <= op <= 0x19:
                push(loc[u1(pc + 1)]); pc += 2
            elif 0x1A <= op <= 0x2D:
                push(loc[(op - 0x1A) & 3]); pc += 1
            elif 0x2E <= op <= 0x35:
                idx = pop(); arr = pop()
                if arr is None:
                    raise JavaThrow("java/lang/NullPointerException")
                if idx < 0 or idx >= len(arr.data):
                    raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException")
                push(arr.data[idx]); pc += 1
            # ---- stores
            elif 0x36 <= op <= 0x3A:
                loc[u1(pc + 1)] = pop(); pc += 2
            elif 0x3B <= op <= 0x4E:
                loc[(op - 0x3B) & 3] = pop(); pc += 1
            elif 0x4F <= op <= 0x56:
                v = pop(); idx = pop(); arr = pop()
                if idx < 0 or idx >= len(arr.data):
                    raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException")
                if op == 0x54:
                    v = _s8(v) if arr.kind != "Z" else v & 1
                elif op == 0x55:
                    v &= 0xFFFF
                elif op == 0x56:
                    v = _s16(v)
                arr.data[idx] = v; pc += 1
            # ---- stack
            elif op == 0x57:
                pop(); pc += 1
            elif op == 0x58:
                v = pop()
                if not _is_cat2(v):
                    pop()
                pc += 1
            elif op == 0x59:
                push(st[-1]); pc += 1
            elif op == 0x5A:
                v1 = pop(); v2 = pop(); push(v1); push(v2); push(v1); pc += 1
            elif op == 0x5B:
                v1 = pop(); v2 = pop()
                if _is_cat2(v2):
                    push(v1); push(v2); push(v1)
                else:
                    v3 = pop(); push(v1); push(v3); push(v2); push(v1)
                pc += 1
            elif op == 0x5C:
                v1 = st[-1]
                if _is_cat2(v1):
                    push(v1)
                else:
                    v2 = st[-2]; push(v2); push(v1)
                pc += 1
            elif op == 0x5D:
                v1 = pop()
                if _is_cat2(v1):
                    v2 = pop(); push(v1); push(v2); push(v1)
                else:
                    v2 = pop(); v3 = pop(); push(v2); push(v1); push(v3); push(v2); push(v1)
                pc += 1
            elif op == 0x5E:
                v1 = pop()
                if _is_cat2(v1):
                    v2 = pop()
                    if _is_cat2(v2):
                        push(v1); push(v2); push(v1)
                    else:
                        v3 = pop(); push(v1); push(v3); push(v2); push(v1)
                else:
                    v2 = pop(); v3 = pop()
                    if _is_cat2(v3):
                        push(v2); push(v1); push(v3); push(v2); push(v1)
                    else:
                        v4 = pop(); push(v2); push(v1); push(v4); push(v3); push(v2); push(v1)
                pc += 1
            elif op == 0x5F:
                v1 = pop(); v2 = pop(); push(v1); push(v2); pc += 1
            # ---- arithmetic
            elif op == 0x60:
                b = pop(); push(_i32(pop() + b)); pc += 1
            elif op == 0x61:
                b = pop(); push(JLong(_i64(pop() + b))); pc += 1
            elif op == 0x62:
                b = pop(); push(_f32(pop() + b)); pc += 1
            elif op == 0x63:
                b = pop(); push(float(pop()) + float(b)); pc += 1
            elif op == 0x64:
                b = pop(); push(_i32(pop() - b)); pc += 1
            elif op == 0x65:
                b = pop(); push(JLong(_i64(pop() - b))); pc += 1
            elif op == 0x66:
                b = pop(); push(_f32(pop() - b)); pc += 1
            elif op == 0x67:
                b = pop(); push(float(pop()) - float(b)); pc += 1
            elif op == 0x68:
                b = pop(); push(_i32(pop() * b)); pc += 1
            elif op == 0x69:
                b = pop(); push(JLong(_i64(pop() * b))); pc += 1
            elif op == 0x6A:
                b = pop(); push(_f32(_dmul(float(pop()), float(b)))); pc += 1
            elif op == 0x6B:
                b = pop(); push(_dmul(float(pop()), float(b))); pc += 1
            elif op in (0x6C, 0x6D):
                b = pop(); a = pop()
                if b == 0:
                    raise JavaThrow("java/lang/ArithmeticException")
                q = abs(a) // abs(b)
                if (a < 0) != (b < 0):
                    q = -q
                push(_i32(q) if op == 0x6C else JLong(_i64(q))); pc += 1
            elif op == 0x6E:
                b = pop(); push(_f32(_ddiv(float(pop()), float(b)))); pc += 1
            elif op == 0x6F:
                b = pop(); push(_ddiv(float(pop()), float(b))); pc += 1
            elif op in (0x70, 0x71):
                b = pop(); a = pop()
                if b == 0:
                    raise JavaThrow("java/lang/ArithmeticException")
                r = abs(a) % abs(b)
                if a < 0:
                    r = -r
                push(_i32(r) if op == 0x70 else JLong(_i64(r))); pc += 1
            elif op in (0x72, 0x73):
                b = pop(); a = pop()
                r = math.fmod(a, b) if (b != 0 and math.isfinite(a) and b == b) else math.nan
                if math.isinf(b) and math.isfinite(a):
                    r = a
                push(_f32(r) if op == 0x72 else r); pc += 1
            elif op == 0x74:
                push(_i32(-pop())); pc += 1
            elif op == 0x75:
                push(JLong(_i64(-pop()))); pc += 1
            elif op == 0x76:
                push(JFloat(-pop())); pc += 1
            elif op == 0x77:
                push(-float(pop())); pc += 1
            elif op == 0x78:
                s = pop() & 31; push(_i32(pop() << s)); pc += 1
            elif op == 0x79:
                s = pop() & 63; push(JLong(_i64(pop() << s))); pc += 1
            elif op == 0x7A:
                s = pop() & 31; push(pop() >> s); pc += 1
            elif op == 0x7B:
                s = pop() & 63; push(JLong(pop() >> s)); pc += 1
            elif op == 0x7C:
                s = pop() & 31; push(_i32((pop() & 0xFFFFFFFF) >> s)); pc += 1
            elif op == 0x7D:
                s = pop() & 63; push(JLong(_i64((pop() & 0xFFFFFFFFFFFFFFFF) >> s))); pc += 1
            elif op == 0x7E:
                b = pop(); push(_i32(pop() & b)); pc += 1
            elif op == 0x7F:
                b = pop(); push(JLong(_i64(pop() & b))); pc += 1
            elif op == 0x80:
                b = pop(); push(_i32(pop() | b)); pc += 1
            elif op == 0x81:
                b = pop(); push(JLong(_i64(pop() | b))); pc += 1
            elif op == 0x82:
                b = pop(); push(_i32(pop() ^ b)); pc += 1
            elif op == 0x83:
                b = pop(); push(JLong(_i64(pop() ^ b))); pc += 1
            elif op == 0x84:
                idx = u1(pc + 1); loc[idx] = _i32(loc[idx] + s1(pc + 2)); pc += 3
            # ---- conversions
            elif op == 0x85:
                push(JLong(pop())); pc += 1
            elif op == 0x86:
                push(_f32(float(pop()))); pc += 1
            elif op == 0x87:
                push(float(pop())); pc += 1
            elif op == 0x88:
                push(_i32(pop())); pc += 1
            elif op == 0x89:
                push(_f32(float(int(pop())))); pc += 1
            elif op == 0x8A:
                push(float(int(pop()))); pc += 1
            elif op == 0x8B:
                push(_d2i(float(pop()), 32)); pc += 1
            elif op == 0x8C:
                push(JLong(_d2i(float(pop()), 64))); pc += 1
            elif op == 0x8D:
                push(float(pop())); pc += 1
            elif op == 0x8E:
                push(_d2i(pop(), 32)); pc += 1
            elif op == 0x8F:
                push(JLong(_d2i(pop(), 64))); pc += 1
            elif op == 0x90:
                push(_f32(pop())); pc += 1
            elif op == 0x91:
                push(_s8(pop())); pc += 1
            elif op == 0x92:
                push(pop() & 0xFFFF); pc += 1
            elif op == 0x93:
                push(_s16(pop())); pc += 1
            # ---- comparisons
            elif op == 0x94:
                b = pop(); a = pop(); push((a > b) - (a < b)); pc += 1
            elif op in (0x95, 0x96, 0x97, 0x98):
                b = pop(); a = pop()
                if a != a or b != b:
                    push(-1 if op in (0x95, 0x97) else 1)
                else:
                    push((a > b) - (a < b))
                pc += 1
            elif 0x99 <= op <= 0x9E:
                v = pop()
                t = (v == 0, v != 0, v < 0, v >= 0, v > 0, v <= 0)[op - 0x99]
                pc = pc + s2(pc + 1) if t else pc + 3
            elif 0x9F <= op <= 0xA4:
                b = pop(); a = pop()
                t = (a == b, a != b, a < b, a >= b, a > b, a <= b)[op - 0x9F]
                pc = pc + s2(pc + 1) if t else pc + 3
            elif op in (0xA5, 0xA6):
                b = pop(); a = pop()
                t = (a is b) if op == 0xA5 else (a is not b)
                pc = pc + s2(pc + 1) if t else pc + 3
            elif op == 0xA7:
                pc = pc + s2(pc + 1)
            elif op == 0xAA:  # tableswitch
                base = pc
                p = (pc + 4) & ~3
                dflt, lo, hi = s4(p), s4(p + 4), s4(p + 8)
                v = pop()
                if v < lo or v > hi:
                    pc = base + dflt
                else:
                    pc = base + s4(p + 12 + 4 * (v - lo))
            elif op == 0xAB:  # lookupswitch
                base = pc
                p = (pc + 4) & ~3
                dflt, npairs = s4(p), s4(p + 4)
                v = pop()
                tgt = dflt
                for k in range(npairs):
                    if s4(p + 8 + 8 * k) == v:
                        tgt = s4(p + 12 + 8 * k)
                        break
                pc = base + tgt
            elif 0xAC <= op <= 0xB0:
                self.n_bytecodes += nb
                return pop()
            elif op == 0xB1:
                self.n_bytecodes += nb
                return None
            # ---- fields
            elif op == 0xB2:
                cname, fname, fdesc = cf.memberref(u2(pc + 1))
                if not self.has_class(cname):
                    push(None)  # opaque JDK static (e.g. a logger); never dereferenced on this path
                else:
                    push(self.find_static_holder(cname, fname).static_fields[fname])
                pc += 3
            elif op == 0xB3:
                cname, fname, fdesc = cf.memberref(u2(pc + 1))
                self.find_static_holder(cname, fname).static_fields[fname] = pop(); pc += 3
            elif op == 0xB4:
                cname, fname, fdesc = cf.memberref(u2(pc + 1))
                o = pop()
                if o is None:
                    raise JavaThrow("java/lang/NullPointerException")
                if fname not in o.fields:
                    o.fields[fname] = {"D": 0.0, "F": JFloat(0.0), "J": JLong(0)}.get(fdesc, 0 if fdesc[0] in "IZBCS" else None)
                push(o.fields[fname]); pc += 3
            elif op == 0xB5:
                cname, fname, fdesc = cf.memberref(u2(pc + 1))
                v = pop(); o = pop(); o.fields[fname] = v; pc += 3
            # ---- invokes
            elif 0xB6 <= op <= 0xB9:
                cname, name, desc = cf.memberref(u2(pc + 1))
                cats = parse_desc(desc)
                nargs = len(cats) + (0 if op == 0xB8 else 1)
                args_ = st[len(st) - nargs:] if nargs else []
                if nargs:
                    del st[len(st) - nargs:]
                kind = {0xB6: "virtual", 0xB7: "special", 0xB8: "static", 0xB9: "interface"}[op]
                if op == 0xB8 and self.has_class(cname):
                    self.load(cname)
                r = self.invoke(kind, cname, name, desc, args_)
                if not desc.endswith(")V"):
                    push(r)
                pc += 5 if op == 0xB9 else 3
            elif op == 0xBA:
                self.invokedynamic(cf, u2(pc + 1), st); pc += 5   # hook: pops the call site's arguments, pushes its result
            elif op == 0xBB:
                cname = cf.utf(cp[u2(pc + 1)][1])
                if self.has_class(cname):
                    self.load(cname)
                push(JObject(cname)); pc += 3
            elif op == 0xBC:
                n = pop()
                kind = {4: "Z", 5: "C", 6: "F", 7: "D", 8: "B", 9: "S", 10: "I", 11: "J"}[u1(pc + 1)]
                zero = {"D": 0.0, "F": JFloat(0.0), "J": JLong(0)}.get(kind, 0)
                push(JArray([zero] * n, kind)); pc += 2
            elif op == 0xBD:
                n = pop(); push(JArray([None] * n, "L")); pc += 3
            elif op == 0xBE:
                arr = pop()
                if arr is None:
                    raise JavaThrow("java/lang/NullPointerException")
                push(len(arr.data)); pc += 1
            elif op == 0xBF:
                raise JavaThrow(pop())
            elif op == 0xC0:
                pc += 3  # checkcast: trust
            elif op == 0xC1:
                o = pop(); cname = cf.utf(cp[u2(pc + 1)][1])
                push(int(o is not None and self._instanceof(o, cname))); pc += 3
            elif op in (0xC2, 0xC3):
                pop(); pc += 1
            elif op == 0xC5:
                dims = u1(pc + 3)
                sizes = [pop() for _ in range(dims)][::-1]
                desc = cf.utf(cp[u2(pc + 1)][1])
                push(self._multi(sizes, desc.lstrip("["))); pc += 4
            elif op == 0xC6:
                pc = pc + s2(pc + 1) if pop() is None else pc + 3
            elif op == 0xC7:
                pc = pc + s2(pc + 1) if pop() is not None else pc + 3
            elif op == 0xC4:  # wide
                op2 = code[pc + 1]
                idx = u2(pc + 2)
                if op2 == 0x84:
                    loc[idx] = _i32(loc[idx] + s2(pc + 4)); pc += 6
                elif 0x15 <= op2 <= 0x19:
                    push(loc[idx]); pc += 4
                elif 0x36 <= op2 <= 0x3A:
                    loc[idx] = pop(); pc += 4
                else:
                    raise NotImplementedError("wide %x" % op2)
            else:
                raise NotImplementedError(f"opcode 0x{op:02x} in {cf.name}.{m.name}{m.desc} @ {pc}")

    def _instanceof(self, o, cname):
        c = getattr(o, "cls", None)
        while c:
            if c == cname:
                return True
            if not self.has_class(c):
                return False
            c = self.load(c).super_name
        return False

    def _multi(self, sizes, elem):
        if len(sizes) == 1:
            zero = {"D": 0.0, "F": JFloat(0.0), "J": JLong(0)}.get(elem, 0 if elem in "IZBCS" else None)
            return JArray([zero] * sizes[0], elem)
        return JArray([self._multi(sizes[1:], elem) for _ in range(sizes[0])], "L")


def _s8(v):
    v &= 0xFF
    return v - 256 if v & 0x80 else v


def _s16(v):
    v &= 0xFFFF
    return v - 65536 if v & 0x8000 else v


# ----------------------------------------------------------------------------- convenience
COMMONS_MATH_JAR = "/root/reference/compiled/commons-math3-3.6.1.jar"


class CommonsMath:
    """The five entry points the oracle needs, executed from the jar's bytecode."""

    def __init__(self, jar=COMMONS_MATH_JAR):
        self.vm = MiniJVM(jar)
        self._FM = "org/apache/commons/math3/util/FastMath"
        self._alt = None

    def exp(self, x):
        return self.vm.call_static(self._FM, "exp", "(D)D", float(x))

    def log(self, x):
        return self.vm.call_static(self._FM, "log", "(D)D", float(x))

    def log1p(self, x):
        return self.vm.call_static(self._FM, "log1p", "(D)D", float(x))

    def log_gamma(self, x):
        return self.vm.call_static("org/apache/commons/math3/special/Gamma", "logGamma", "(D)D", float(x))

    def log_beta(self, a, b):
        return self.vm.call_static("org/apache/commons/math3/special/Beta", "logBeta", "(DD)D", float(a), float(b))

    def regularized_beta(self, x, a, b):
        return self.vm.call_static("org/apache/commons/math3/special/Beta", "regularizedBeta", "(DDD)D",
                                   float(x), float(a), float(b))

    def binomial_test_greater(self, n, k, p):
        """BinomialTest.binomialTest(n, k, p, AlternativeHypothesis.GREATER_THAN) — call site HC.java:3050."""
        vm = self.vm
        if self._alt is None:
            alt_cls = "org/apache/commons/math3/stat/inference/AlternativeHypothesis"
            vm.load(alt_cls)
            self._alt = vm.find_static_holder(alt_cls, "GREATER_THAN").static_fields["GREATER_THAN"]
            self._bt = JObject("org/apache/commons/math3/stat/inference/BinomialTest")
        m = vm.find_method("org/apache/commons/math3/stat/inference/BinomialTest", "binomialTest",
                           "(IIDLorg/apache/commons/math3/stat/inference/AlternativeHypothesis;)D")
        return vm.run(m, [self._bt, int(n), int(k), float(p), self._alt])


if __name__ == "__main__":
    cm = CommonsMath()
    print("exp(1)      =", repr(cm.exp(1.0)), "  math:", repr(math.exp(1.0)))
    print("log(10)     =", repr(cm.log(10.0)), "  math:", repr(math.log(10.0)))
    print("log1p(-.38) =", repr(cm.log1p(-0.38)), "  math:", repr(math.log1p(-0.38)))
    print("logGamma(7.5) =", repr(cm.log_gamma(7.5)), "  math:", repr(math.lgamma(7.5)))
    print("binomialTest(9,9,0.379,>) =", repr(cm.binomial_test_greater(9, 9, 0.379)), " exact:", 0.379 ** 9)
    print("binomialTest(20,3,0.379,>) =", repr(cm.binomial_test_greater(20, 3, 0.379)))
    print("bytecodes executed:", cm.vm.n_bytecodes)
