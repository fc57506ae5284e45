"""Tiny bytecode lister (a `javap -c` stand-in; there is no JDK in the authoring container).

TEST INFRASTRUCTURE ONLY.  Used while restating the commons-math3 3.6.1 arithmetic chain
(the jar is the only form in which the reference ships it).

usage: python oracle/tools/javap_lite.py <jar> <class/name> [method-name-substring]
"""
import struct
import sys
import zipfile

from minijvm import ClassFile

NAMES = {
    0x00: "nop", 0x01: "aconst_null", 0x02: "iconst_m1", 0x03: "iconst_0", 0x04: "iconst_1", 0x05: "iconst_2",
    0x06: "iconst_3", 0x07: "iconst_4", 0x08: "iconst_5", 0x09: "lconst_0", 0x0A: "lconst_1", 0x0B: "fconst_0",
    0x0C: "fconst_1", 0x0D: "fconst_2", 0x0E: "dconst_0", 0x0F: "dconst_1", 0x10: "bipush", 0x11: "sipush",
    0x12: "ldc", 0x13: "ldc_w", 0x14: "ldc2_w", 0x15: "iload", 0x16: "lload", 0x17: "fload", 0x18: "dload",
    0x19: "aload", 0x2E: "iaload", 0x2F: "laload", 0x30: "faload", 0x31: "daload", 0x32: "aaload", 0x33: "baload",
    0x34: "caload", 0x35: "saload", 0x36: "istore", 0x37: "lstore", 0x38: "fstore", 0x39: "dstore", 0x3A: "astore",
    0x4F: "iastore", 0x50: "lastore", 0x51: "fastore", 0x52: "dastore", 0x53: "aastore", 0x54: "bastore",
    0x55: "castore", 0x56: "sastore", 0x57: "pop", 0x58: "pop2", 0x59: "dup", 0x5A: "dup_x1", 0x5B: "dup_x2",
    0x5C: "dup2", 0x5D: "dup2_x1", 0x5E: "dup2_x2", 0x5F: "swap", 0x60: "iadd", 0x61: "ladd", 0x62: "fadd",
    0x63: "dadd", 0x64: "isub", 0x65: "lsub", 0x66: "fsub", 0x67: "dsub", 0x68: "imul", 0x69: "lmul", 0x6A: "fmul",
    0x6B: "dmul", 0x6C: "idiv", 0x6D: "ldiv", 0x6E: "fdiv", 0x6F: "ddiv", 0x70: "irem", 0x71: "lrem", 0x72: "frem",
    0x73: "drem", 0x74: "ineg", 0x75: "lneg", 0x76: "fneg", 0x77: "dneg", 0x78: "ishl", 0x79: "lshl", 0x7A: "ishr",
    0x7B: "lshr", 0x7C: "iushr", 0x7D: "lushr", 0x7E: "iand", 0x7F: "land", 0x80: "ior", 0x81: "lor", 0x82: "ixor",
    0x83: "lxor", 0x84: "iinc", 0x85: "i2l", 0x86: "i2f", 0x87: "i2d", 0x88: "l2i", 0x89: "l2f", 0x8A: "l2d",
    0x8B: "f2i", 0x8C: "f2l", 0x8D: "f2d", 0x8E: "d2i", 0x8F: "d2l", 0x90: "d2f", 0x91: "i2b", 0x92: "i2c",
    0x93: "i2s", 0x94: "lcmp", 0x95: "fcmpl", 0x96: "fcmpg", 0x97: "dcmpl", 0x98: "dcmpg", 0x99: "ifeq", 0x9A: "ifne",
    0x9B: "iflt", 0x9C: "ifge", 0x9D: "ifgt", 0x9E: "ifle", 0x9F: "if_icmpeq", 0xA0: "if_icmpne", 0xA1: "if_icmplt",
    0xA2: "if_icmpge", 0xA3: "if_icmpgt", 0xA4: "if_icmple", 0xA5: "if_acmpeq", 0xA6: "if_acmpne", 0xA7: "goto",
    0xAA: "tableswitch", 0xAB: "lookupswitch", 0xAC: "ireturn", 0xAD: "lreturn", 0xAE: "freturn", 0xAF: "dreturn",
    0xB0: "areturn", 0xB1: "return", 0xB2: "getstatic", 0xB3: "putstatic", 0xB4: "getfield", 0xB5: "putfield",
    0xB6: "invokevirtual", 0xB7: "invokespecial", 0xB8: "invokestatic", 0xB9: "invokeinterface", 0xBB: "new",
    0xBC: "newarray", 0xBD: "anewarray", 0xBE: "arraylength", 0xBF: "athrow", 0xC0: "checkcast", 0xC1: "instanceof",
    0xC2: "monitorenter", 0xC3: "monitorexit", 0xC4: "wide", 0xC5: "multianewarray", 0xC6: "ifnull", 0xC7: "ifnonnull",
}
for base, nm in ((0x1A, "iload"), (0x1E, "lload"), (0x22, "fload"), (0x26, "dload"), (0x2A, "aload"),
                 (0x3B, "istore"), (0x3F, "lstore"), (0x43, "fstore"), (0x47, "dstore"), (0x4B, "astore")):
    for k in range(4):
        NAMES[base + k] = f"{nm}_{k}"


def disasm(cf: ClassFile, m):
    code = m.code
    pc = 0
    out = []
    while pc < len(code):
        op = code[pc]
        nm = NAMES.get(op, f"op_{op:02x}")
        arg = ""
        ln = 1
        if op in (0x10, 0xBC):
            arg = str(code[pc + 1] - 256 if (op == 0x10 and code[pc + 1] > 127) else code[pc + 1]); ln = 2
        elif op == 0x11:
            arg = str(struct.unpack_from(">h", code, pc + 1)[0]); ln = 3
        elif op == 0x12:
            arg = repr(cf.const(code[pc + 1])); ln = 2
        elif op in (0x13, 0x14):
            arg = repr(cf.const(struct.unpack_from(">H", code, pc + 1)[0])); ln = 3
        elif 0x15 <= op <= 0x19 or 0x36 <= op <= 0x3A:
            arg = str(code[pc + 1]); ln = 2
        elif op == 0x84:
            arg = f"{code[pc + 1]} {code[pc + 2] - 256 if code[pc + 2] > 127 else code[pc + 2]}"; ln = 3
        elif 0x99 <= op <= 0xA7 or op in (0xC6, 0xC7):
            arg = "-> " + str(pc + struct.unpack_from(">h", code, pc + 1)[0]); ln = 3
        elif 0xB2 <= op <= 0xB9:
            c, n, d = cf.memberref(struct.unpack_from(">H", code, pc + 1)[0])
            arg = f"{c.split('/')[-1]}.{n} {d}"; ln = 5 if op == 0xB9 else 3
        elif op in (0xBB, 0xBD, 0xC0, 0xC1):
            arg = cf.utf(cf.cp[struct.unpack_from(">H", code, pc + 1)[0]][1]); ln = 3
        elif op == 0xAA:
            p = (pc + 4) & ~3
            d, lo, hi = struct.unpack_from(">iii", code, p)
            tg = [pc + struct.unpack_from(">i", code, p + 12 + 4 * k)[0] for k in range(hi - lo + 1)]
            arg = f"lo={lo} hi={hi} default->{pc + d} targets->{tg}"; ln = p + 12 + 4 * (hi - lo + 1) - pc
        elif op == 0xAB:
            p = (pc + 4) & ~3
            d, n = struct.unpack_from(">ii", code, p)
            prs = [(struct.unpack_from(">i", code, p + 8 + 8 * k)[0], pc + struct.unpack_from(">i", code, p + 12 + 8 * k)[0]) for k in range(n)]
            arg = f"default->{pc + d} pairs={prs}"; ln = p + 8 + 8 * n - pc
        elif op == 0xC5:
            arg = cf.utf(cf.cp[struct.unpack_from(">H", code, pc + 1)[0]][1]) + f" dims={code[pc + 3]}"; ln = 4
        elif op == 0xC4:
            ln = 6 if code[pc + 1] == 0x84 else 4
            arg = f"{NAMES.get(code[pc + 1])} {struct.unpack_from('>H', code, pc + 2)[0]}"
        out.append(f"{pc:5d}: {nm} {arg}")
        pc += ln
    return out


if __name__ == "__main__":
    jar, cname = sys.argv[1], sys.argv[2]
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    cf = ClassFile(zipfile.ZipFile(jar).read(cname + ".class"))
    for fname, (fdesc, acc) in cf.field_descs.items():
        if acc & 8 and filt == "":
            print(f"static {fdesc} {fname} = {cf.static_fields.get(fname)!r}")
    for (name, desc), m in cf.methods.items():
        if filt in name and m.code is not None:
            print(f"\n=== {cf.name}.{name}{desc}  (locals={m.max_locals})")
            print("\n".join(disasm(cf, m)))
