#!/usr/bin/env python3
"""Minimal JVM .class disassembler -- a READING AID for SURVEY.md, not product code.

The reference hot path ships only as bytecode (Jar/NanoporeBC_UMI_finder-2.1.jar and
Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar); this container has no JDK/javap.
The classes were compiled with -g, so LineNumberTable + LocalVariableTable survive:
every `Lnnn` prefix printed below is the ORIGINAL .java source line, which is what
SURVEY.md cites as `<jar>!/<class> (<File>.java:Lnnn)`.

Usage:
  python3 tools/jdis.py /root/reference/Jar/NanoporeBC_UMI_finder-2.1.jar BarcodeMatchTester
  python3 tools/jdis.py <jar> <class-substring> -m doJob substitutions   # only these methods
  python3 tools/jdis.py <jar> <class-substring> --nocode                 # signatures only
"""
import struct, sys, zipfile, io

OPS = {}
def _d(code, name, fmt=''):
    OPS[code] = (name, fmt)
names0 = """0 nop;1 aconst_null;2 iconst_m1;3 iconst_0;4 iconst_1;5 iconst_2;6 iconst_3;7 iconst_4;8 iconst_5;9 lconst_0;10 lconst_1;
11 fconst_0;12 fconst_1;13 fconst_2;14 dconst_0;15 dconst_1;26 iload_0;27 iload_1;28 iload_2;29 iload_3;30 lload_0;31 lload_1;32 lload_2;33 lload_3;
34 fload_0;35 fload_1;36 fload_2;37 fload_3;38 dload_0;39 dload_1;40 dload_2;41 dload_3;42 aload_0;43 aload_1;44 aload_2;45 aload_3;
46 iaload;47 laload;48 faload;49 daload;50 aaload;51 baload;52 caload;53 saload;59 istore_0;60 istore_1;61 istore_2;62 istore_3;
63 lstore_0;64 lstore_1;65 lstore_2;66 lstore_3;67 fstore_0;68 fstore_1;69 fstore_2;70 fstore_3;71 dstore_0;72 dstore_1;73 dstore_2;74 dstore_3;
75 astore_0;76 astore_1;77 astore_2;78 astore_3;79 iastore;80 lastore;81 fastore;82 dastore;83 aastore;84 bastore;85 castore;86 sastore;
87 pop;88 pop2;89 dup;90 dup_x1;91 dup_x2;92 dup2;93 dup2_x1;94 dup2_x2;95 swap;96 iadd;97 ladd;98 fadd;99 dadd;100 isub;101 lsub;102 fsub;103 dsub;
104 imul;105 lmul;106 fmul;107 dmul;108 idiv;109 ldiv;110 fdiv;111 ddiv;112 irem;113 lrem;114 frem;115 drem;116 ineg;117 lneg;118 fneg;119 dneg;
120 ishl;121 lshl;122 ishr;123 lshr;124 iushr;125 lushr;126 iand;127 land;128 ior;129 lor;130 ixor;131 lxor;133 i2l;134 i2f;135 i2d;136 l2i;137 l2f;138 l2d;
139 f2i;140 f2l;141 f2d;142 d2i;143 d2l;144 d2f;145 i2b;146 i2c;147 i2s;148 lcmp;149 fcmpl;150 fcmpg;151 dcmpl;152 dcmpg;
172 ireturn;173 lreturn;174 freturn;175 dreturn;176 areturn;177 return;190 arraylength;191 athrow;194 monitorenter;195 monitorexit"""
for tok in names0.replace('\n', '').split(';'):
    c, n = tok.split()
    _d(int(c), n)
_d(16, 'bipush', 'b'); _d(17, 'sipush', 's'); _d(18, 'ldc', 'C1'); _d(19, 'ldc_w', 'C2'); _d(20, 'ldc2_w', 'C2')
for c, n in [(21, 'iload'), (22, 'lload'), (23, 'fload'), (24, 'dload'), (25, 'aload'), (54, 'istore'), (55, 'lstore'), (56, 'fstore'), (57, 'dstore'), (58, 'astore'), (169, 'ret')]:
    _d(c, n, 'L')
_d(132, 'iinc', 'Lb')
for c, n in [(153, 'ifeq'), (154, 'ifne'), (155, 'iflt'), (156, 'ifge'), (157, 'ifgt'), (158, 'ifle'), (159, 'if_icmpeq'), (160, 'if_icmpne'), (161, 'if_icmplt'), (162, 'if_icmpge'),
             (163, 'if_icmpgt'), (164, 'if_icmple'), (165, 'if_acmpeq'), (166, 'if_acmpne'), (167, 'goto'), (168, 'jsr'), (198, 'ifnull'), (199, 'ifnonnull')]:
    _d(c, n, 'J2')
_d(200, 'goto_w', 'J4'); _d(201, 'jsr_w', 'J4')
for c, n in [(178, 'getstatic'), (179, 'putstatic'), (180, 'getfield'), (181, 'putfield'), (182, 'invokevirtual'), (183, 'invokespecial'), (184, 'invokestatic'),
             (187, 'new'), (189, 'anewarray'), (192, 'checkcast'), (193, 'instanceof')]:
    _d(c, n, 'C2')
_d(185, 'invokeinterface', 'C2bb'); _d(186, 'invokedynamic', 'C2bb'); _d(188, 'newarray', 'A'); _d(197, 'multianewarray', 'C2b')
_d(170, 'tableswitch', 'T'); _d(171, 'lookupswitch', 'K'); _d(196, 'wide', 'W')
ATYPE = {4: 'boolean', 5: 'char', 6: 'float', 7: 'double', 8: 'byte', 9: 'short', 10: 'int', 11: 'long'}


class R:
    def __init__(self, b): self.b = b; self.p = 0
    def u1(self): v = self.b[self.p]; self.p += 1; return v
    def u2(self): v = struct.unpack_from('>H', self.b, self.p)[0]; self.p += 2; return v
    def u4(self): v = struct.unpack_from('>I', self.b, self.p)[0]; self.p += 4; return v
    def s1(self): v = struct.unpack_from('>b', self.b, self.p)[0]; self.p += 1; return v
    def s2(self): v = struct.unpack_from('>h', self.b, self.p)[0]; self.p += 2; return v
    def s4(self): v = struct.unpack_from('>i', self.b, self.p)[0]; self.p += 4; return v
    def raw(self, n): v = self.b[self.p:self.p + n]; self.p += n; return v


def mutf8(b):
    try:
        return b.decode('utf-8')
    except Exception:
        return b.decode('latin-1')


class ClassFile:
    def __init__(self, data):
        r = R(data)
        assert r.u4() == 0xCAFEBABE
        self.minor = r.u2(); self.major = r.u2()
        n = r.u2(); cp = [None] * n; i = 1
        while i < n:
            t = r.u1()
            if t == 1: l = r.u2(); cp[i] = ('Utf8', mutf8(r.raw(l)))
            elif t == 3: cp[i] = ('Int', r.s4())
            elif t == 4: cp[i] = ('Float', struct.unpack('>f', r.raw(4))[0])
            elif t == 5: cp[i] = ('Long', struct.unpack('>q', r.raw(8))[0]); i += 1
            elif t == 6: cp[i] = ('Double', struct.unpack('>d', r.raw(8))[0]); i += 1
            elif t == 7: cp[i] = ('Class', r.u2())
            elif t == 8: cp[i] = ('String', r.u2())
            elif t in (9, 10, 11): cp[i] = ({9: 'Field', 10: 'Method', 11: 'IMethod'}[t], r.u2(), r.u2())
            elif t == 12: cp[i] = ('NaT', r.u2(), r.u2())
            elif t == 15: cp[i] = ('MHandle', r.u1(), r.u2())
            elif t == 16: cp[i] = ('MType', r.u2())
            elif t in (17, 18): cp[i] = ('Dyn' if t == 17 else 'InDyn', r.u2(), r.u2())
            elif t in (19, 20): cp[i] = ('Module' if t == 19 else 'Package', r.u2())
            else: raise ValueError('cp tag %d' % t)
            i += 1
        self.cp = cp
        self.access = r.u2(); self.this = self.cls(r.u2()); sc = r.u2(); self.super = self.cls(sc) if sc else None
        self.ifaces = [self.cls(r.u2()) for _ in range(r.u2())]
        self.fields = [self.member(r) for _ in range(r.u2())]
        self.methods = [self.member(r) for _ in range(r.u2())]
        self.attrs = self.attributes(r)

    def utf(self, i): return self.cp[i][1]
    def cls(self, i): return self.utf(self.cp[i][1])
    def nat(self, i): e = self.cp[i]; return self.utf(e[1]), self.utf(e[2])
    def const(self, i):
        e = self.cp[i]; t = e[0]
        if t == 'Utf8': return repr(e[1])
        if t in ('Int', 'Long'): return '%s %d' % (t, e[1])
        if t in ('Float', 'Double'): return '%s %r' % (t, e[1])
        if t == 'Class': return 'class ' + self.utf(e[1])
        if t == 'String': return 'String ' + repr(self.utf(e[1]))
        if t in ('Field', 'Method', 'IMethod'):
            n, d = self.nat(e[2]); return '%s %s.%s:%s' % (t, self.cls(e[1]), n, d)
        if t == 'InDyn' or t == 'Dyn':
            n, d = self.nat(e[2]); return '%s #%d %s:%s%s' % (t, e[1], n, d, self.bsm(e[1]))
        if t == 'MHandle': return 'MHandle kind%d %s' % (e[1], self.const(e[2]))
        if t == 'MType': return 'MType ' + self.utf(e[1])
        return str(e)

    def bsm(self, idx):
        for name, data in self.attrs:
            if name == 'BootstrapMethods':
                r = R(data); n = r.u2(); out = None
                for k in range(n):
                    ref = r.u2(); na = r.u2(); args = [r.u2() for _ in range(na)]
                    if k == idx:
                        out = ' {bsm=%s args=[%s]}' % (self.const(ref).split(' ', 2)[-1].split(':')[0], '; '.join(self.const(a) for a in args))
                return out or ''
        return ''

    def attributes(self, r):
        out = []
        for _ in range(r.u2()):
            name = self.utf(r.u2()); l = r.u4(); out.append((name, r.raw(l)))
        return out

    def member(self, r):
        acc = r.u2(); name = self.utf(r.u2()); desc = self.utf(r.u2()); attrs = self.attributes(r)
        return (acc, name, desc, attrs)


def accstr(a, method=False):
    s = []
    for bit, n in [(1, 'public'), (2, 'private'), (4, 'protected'), (8, 'static'), (16, 'final'), (0x20, 'synchronized' if method else 'super'),
                   (0x40, 'bridge' if method else 'volatile'), (0x80, 'varargs' if method else 'transient'), (0x100, 'native'), (0x200, 'interface'),
                   (0x400, 'abstract'), (0x1000, 'synthetic'), (0x4000, 'enum')]:
        if a & bit and n != 'super': s.append(n)
    return ' '.join(s)


def disasm(cf, code, lines, lvt, out):
    r = R(code); n = len(code)
    def lv(idx, pc):
        for (s, l, nm, d, i) in lvt:
            if i == idx and s <= pc + 4 and pc < s + l + 1:
                return '%d(%s)' % (idx, nm)
        return str(idx)
    while r.p < n:
        pc = r.p; op = r.u1()
        if op not in OPS:
            out.append('   %5d: <op %d>' % (pc, op)); continue
        name, fmt = OPS[op]; args = []
        if fmt == 'b': args.append(str(r.s1()))
        elif fmt == 's': args.append(str(r.s2()))
        elif fmt == 'C1': args.append(cf.const(r.u1()))
        elif fmt == 'C2': args.append(cf.const(r.u2()))
        elif fmt == 'C2bb': args.append(cf.const(r.u2())); r.u1(); r.u1()
        elif fmt == 'C2b': args.append(cf.const(r.u2())); args.append('dims=%d' % r.u1())
        elif fmt == 'L': args.append(lv(r.u1(), pc))
        elif fmt == 'Lb': i = r.u1(); args.append(lv(i, pc)); args.append(str(r.s1()))
        elif fmt == 'J2': args.append('-> %d' % (pc + r.s2()))
        elif fmt == 'J4': args.append('-> %d' % (pc + r.s4()))
        elif fmt == 'A': args.append(ATYPE.get(r.u1(), '?'))
        elif fmt == 'W':
            op2 = r.u1(); nm2 = OPS[op2][0]; idx = r.u2(); name = 'wide ' + nm2; args.append(lv(idx, pc))
            if op2 == 132: args.append(str(r.s2()))
        elif fmt == 'T':
            while r.p % 4: r.u1()
            df = r.s4(); lo = r.s4(); hi = r.s4()
            tg = ['%d->%d' % (lo + k, pc + r.s4()) for k in range(hi - lo + 1)]
            args.append('{%s default->%d}' % (', '.join(tg), pc + df))
        elif fmt == 'K':
            while r.p % 4: r.u1()
            df = r.s4(); np_ = r.s4()
            tg = []
            for k in range(np_):
                m = r.s4(); o = r.s4(); tg.append('%d->%d' % (m, pc + o))
            args.append('{%s default->%d}' % (', '.join(tg), pc + df))
        if name.split('_')[0] in ('iload', 'lload', 'fload', 'dload', 'aload', 'istore', 'lstore', 'fstore', 'dstore', 'astore') and '_' in name and not fmt:
            idx = int(name.split('_')[1]); nm = lv(idx, pc)
            if '(' in nm: args.append('; ' + nm)
        ln = lines.get(pc)
        out.append('%s %5d: %s %s' % (('L%-4d' % ln) if ln is not None else '     ', pc, name, ' '.join(args)))


def dump(data, only=None, nocode=False):
    cf = ClassFile(data); out = []
    src = [cf.utf(struct.unpack('>H', d)[0]) for n, d in cf.attrs if n == 'SourceFile']
    out.append('=== class %s extends %s implements %s [%s] major=%d src=%s' % (cf.this, cf.super, ','.join(cf.ifaces), accstr(cf.access), cf.major, src))
    for n, d in cf.attrs:
        if n == 'Signature': out.append('  signature ' + cf.utf(struct.unpack('>H', d)[0]))
        if n == 'InnerClasses':
            r = R(d)
            for _ in range(r.u2()):
                a, b, c, e = r.u2(), r.u2(), r.u2(), r.u2()
                out.append('  inner %s' % cf.cls(a))
    for acc, name, desc, attrs in cf.fields:
        extra = ''
        for an, ad in attrs:
            if an == 'ConstantValue': extra += ' = ' + cf.const(struct.unpack('>H', ad)[0])
            if an == 'Signature': extra += ' sig=' + cf.utf(struct.unpack('>H', ad)[0])
        out.append('  field %s %s %s%s' % (accstr(acc), name, desc, extra))
    for acc, name, desc, attrs in cf.methods:
        if only and name not in only: continue
        sig = ''
        for an, ad in attrs:
            if an == 'Signature': sig = ' sig=' + cf.utf(struct.unpack('>H', ad)[0])
        out.append('  method %s %s%s%s' % (accstr(acc, True), name, desc, sig))
        if nocode: continue
        for an, ad in attrs:
            if an == 'Code':
                r = R(ad); ms = r.u2(); ml = r.u2(); cl = r.u4(); code = r.raw(cl)
                ex = [(r.u2(), r.u2(), r.u2(), r.u2()) for _ in range(r.u2())]
                cattrs = cf.attributes(r); lines = {}; lvt = []
                for cn, cd in cattrs:
                    rr = R(cd)
                    if cn == 'LineNumberTable':
                        for _ in range(rr.u2()):
                            s = rr.u2(); l = rr.u2(); lines[s] = l
                    if cn == 'LocalVariableTable':
                        for _ in range(rr.u2()):
                            s = rr.u2(); l = rr.u2(); nm = cf.utf(rr.u2()); de = cf.utf(rr.u2()); ix = rr.u2(); lvt.append((s, l, nm, de, ix))
                out.append('    stack=%d locals=%d codelen=%d' % (ms, ml, cl))
                if lvt:
                    out.append('    locals: ' + ', '.join('%d=%s:%s' % (ix, nm, de) for s, l, nm, de, ix in sorted(lvt, key=lambda x: (x[4], x[0]))))
                disasm(cf, code, lines, lvt, out)
                for s, e, h, t in ex:
                    out.append('    catch %s [%d,%d) -> %d' % (cf.cls(t) if t else 'any', s, e, h))
    return '\n'.join(out)


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('jar'); ap.add_argument('cls'); ap.add_argument('-m', nargs='*'); ap.add_argument('--nocode', action='store_true')
    a = ap.parse_args()
    z = zipfile.ZipFile(a.jar)
    for nm in z.namelist():
        if nm.endswith('.class') and (a.cls in nm):
            print(dump(z.read(nm), a.m, a.nocode))
