"""Second, independent restatement of the reference path (pure Python, object-for-object like the Java).

TEST INFRASTRUCTURE ONLY.  It exists so that the C oracle (slr_oracle.c) is checked by a differently
structured implementation of the same bytecode: Java objects become Python objects, `long` becomes a
masked int, java.util.HashMap is modelled with real bucket lists and a real resize().  Slow: use on small
lists / few windows.  Citations as in slr_oracle.c (F!/T! jars, original .java line numbers).
"""
from collections import deque

M64 = (1 << 64) - 1


def _shl(x, n):
    return (x << (n & 63)) & M64


def _ushr(x, n):
    return (x & M64) >> (n & 63)


def _sext8(b):
    return b & M64 if b >= 0 else (b + (1 << 64)) & M64


# ---- T!com/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase ------------------------------------------
BASE_TO_TWOBIT = [-2] * 254                                   # java:L78-L79
for ch, v in (("A", 0), ("a", 0), ("G", 1), ("g", 1), ("C", 2), ("c", 2), ("T", 3), ("t", 3)):
    BASE_TO_TWOBIT[ord(ch)] = v                               # java:L80-L87
REVERSE_COMP = [3, 2, 1, 0]                                   # java:L72-L76
CLEAR_BITS = []                                               # java:L89-L100
_l = (-4) & M64
for _i in range(32):
    CLEAR_BITS.append(_l)
    _l = ((_l << 2) | 3) & M64
SET_BITS = [[0, 1, 3, 2]]                                     # java:L105-L112
for _i in range(31):
    SET_BITS.append([(v << 2) & M64 for v in SET_BITS[-1]])

# ---- T!com/rw/nuc/encoding/NucleicAcidByteCodeBase --------------------------------------------------
ENCODE = [-1] * 254                                           # java:L45-L46
for chs, v in (("-", 0), ("Aa", 1), ("Gg", 2), ("Cc", 4), ("Tt", 8), ("Nn", 15), ("Hh", 13), ("Rr", 3), ("Yy", 12),
               ("Mm", 5), ("Kk", 10), ("Ss", 6), ("Ww", 9), ("Bb", 14), ("Vv", 7), ("Dd", 11)):
    for ch in chs:
        ENCODE[ord(ch)] = v                                   # java:L48-L78
ONEBYTE_RC = [-1] * 16                                        # java:L100-L133
for a, b in (("-", "-"), ("A", "T"), ("G", "C"), ("C", "G"), ("T", "A"), ("N", "N"), ("H", "D"), ("R", "Y"), ("Y", "R"),
             ("M", "K"), ("K", "M"), ("S", "S"), ("W", "W"), ("B", "V"), ("V", "B"), ("D", "H")):
    ONEBYTE_RC[ENCODE[ord(a)]] = ENCODE[ord(b)]
BYTE_TO_2BITLONG_0 = [0] * 16                                 # java:L92-L98, row 0
for ch in "AGCT":
    BYTE_TO_2BITLONG_0[ENCODE[ord(ch)]] = BASE_TO_TWOBIT[ord(ch)]


class JavaException(Exception):
    pass


def pack(s):                                                  # getLongHashForSeq java:L183-L187
    r = 0
    for ch in s:
        c = ord(ch) if isinstance(ch, str) else ch
        if c >= 254:
            raise JavaException("AIOOBE BASE_TO_TWOBIT_ARRAY")
        r = (_shl(r, 2) | _sext8(BASE_TO_TWOBIT[c])) & M64
    return r


def revcomp2(seq, L):                                         # java:L477-L484
    t = 0
    for _ in range(L):
        t = _shl(t, 2)
        t |= REVERSE_COMP[seq & 3]
        seq = _ushr(seq, 2)
    return t


def replace_deg(seq, pos, L):                                 # java:L228-L234
    seq &= CLEAR_BITS[L - pos - 1]
    shift = (L - (pos + 1)) << 1
    return [seq | _shl(b, shift) for b in range(4)]


def insert_deg(h, pos, L):                                    # java:L300-L310
    shift = (L - pos - 1) << 1
    upper = _shl(_ushr(h, shift), shift)
    shift = 64 - shift
    h = _shl(h, shift)
    h = _ushr(h, shift + 2)
    row = SET_BITS[L - (pos + 1) - 1]
    return [upper | h | row[b] for b in (0, 1, 3, 2)]


def delete_byte(h, code4, pos, L):                            # java:L321-L327
    shift = (L - pos) << 1
    upper = _shl(_ushr(h, shift), shift)
    shift = 64 - shift
    h = _shl(h, shift + 2)
    h = _ushr(h, shift)
    if not 0 <= code4 < 16:
        raise JavaException("AIOOBE BYTE_TO_2BITLONG_ARRAY")
    return upper | h | BYTE_TO_2BITLONG_0[code4]


# ---- F!com/rw/nuc/encoding/TwoBit/LongSeqMutated ------------------------------------------------------
class LongSeqMutated:
    __slots__ = ("seq", "L", "unmut", "nSub", "nIns", "nDel", "offset", "posPrev", "posCur", "level")

    def __init__(self, seq, L, offset):                       # java:L61 -> L44-L50
        self.seq, self.L, self.offset = seq, L, offset
        self.unmut = None
        self.nSub = self.nIns = self.nDel = 0
        self.posPrev = self.posCur = -1
        self.level = 0

    def copy(self):                                           # java:L68-L77
        c = LongSeqMutated(self.seq, self.L, self.offset)
        c.unmut = self.unmut
        c.nSub, c.nIns, c.nDel = self.nSub, self.nIns, self.nDel
        c.posPrev, c.posCur, c.level = self.posPrev, self.posCur, self.level
        return c


class OneMatch:                                               # BarcodeMatchTester$Matches$OneMatch
    def __init__(self, readSeq, ed, offset, bc, L, nSub, nIns, nDel):
        self.readSeq, self.ed, self.offset, self.bc = readSeq, ed, offset, bc
        self.L, self.nSub, self.nIns, self.nDel = L, nSub, nIns, nDel

    def jhash(self):                                          # java:L443
        h = (self.readSeq ^ (self.readSeq >> 32)) & 0xFFFFFFFF
        return h

    def jequals(self, o):                                     # java:L433-L436
        return o.readSeq == self.readSeq and o.ed == self.ed and o.offset == self.offset

    def key(self):                                            # compareTo java:L449-L461 is a weak order on this key
        return (self.ed, 0 if self.offset == 0 else 1)


class JHashSet:
    """java.util.HashSet / HashMap semantics that influence iteration order."""

    def __init__(self):
        self.table = None
        self.size = 0
        self.threshold = 0
        self.treeified = False

    def _resize(self):
        if self.table is None:
            self.table = [[] for _ in range(16)]
            self.threshold = 12
            return
        old = self.table
        ncap = len(old) * 2
        self.threshold *= 2
        self.table = [[] for _ in range(ncap)]
        for j, chain in enumerate(old):
            for (h, e) in chain:                              # split keeps relative order
                self.table[h & (ncap - 1)].append((h, e))

    def add(self, e):
        hc = e.jhash()
        h = (hc ^ (hc >> 16)) & 0xFFFFFFFF
        if self.table is None:
            self._resize()
        chain = self.table[h & (len(self.table) - 1)]
        for (h2, e2) in chain:
            if h2 == h and e2.jequals(e):
                return False
        chain.append((h, e))
        if len(chain) >= 9:                                   # binCount >= TREEIFY_THRESHOLD - 1
            if len(self.table) < 64:
                self._resize()
            else:
                self.treeified = True
        self.size += 1
        if self.size > self.threshold:
            self._resize()
        return True

    def __iter__(self):
        if self.table is None:
            return
        for chain in self.table:
            for (_, e) in chain:
                yield e


# ---- F!com/rw/nanoporereadscanner/analyzers/BarcodeMatchTester ---------------------------------------
class BarcodeMatchTester:
    def __init__(self, seq, L, ed, skipFullMatches, allowIndels, searchSet, offset, post, doNext):
        self.seq, self.L, self.ed = seq, L, ed
        self.skipFull, self.allowIndels, self.set = skipFullMatches, allowIndels, searchSet
        self.offset, self.post, self.doNext = offset, post, doNext
        self.deque = deque()
        self.matches = JHashSet()
        self.probes = 0
        # NucTwoBitPerBaseEDtesterBase ctor java:L82-L95
        if ed >= 2:
            self.tested = set()
            self.use64 = L >= 14 and L > 16
        else:
            self.tested = None

    def _key(self, s):
        return s if self.use64 else (s & 0xFFFFFFFF)

    def already(self, s):                                     # java:L120
        return self.tested is not None and self._key(s) in self.tested

    def mark(self, s):                                        # java:L105-L112
        if self.tested is not None:
            self.tested.add(self._key(s))

    def check(self, n):                                       # java:L367-L374
        if self.skipFull and n.unmut == n.seq:
            return None
        self.probes += 1
        if n.seq in self.set:
            return OneMatch(n.unmut, n.level, n.offset, n.seq, self.L, n.nSub, n.nIns, n.nDel)
        return None

    def goNext(self, n):                                      # NucTwoBitPerBaseEDtesterBase java:L133-L144
        if self.ed > n.level:
            c = n.copy()
            c.posPrev = n.posCur
            c.posCur = -1
            c.level = n.level + 1
            self.deque.append(c)

    def doJob(self):                                          # java:L198-L244
        parent = LongSeqMutated(self.seq, self.L, self.offset)
        parent.unmut = self.seq
        r = self.check(parent)
        if r is not None:
            self.matches.add(r)
        if self.ed == 0:
            return self.matches
        parent.level = 1
        self.deque.append(parent)
        Lm1 = self.L - 1
        while self.deque:
            cur = self.deque.pop()                            # pollLast
            cur.posCur += 1
            if cur.posCur < Lm1:
                self.deque.append(cur.copy())
            if cur.posPrev == cur.posCur:
                continue
            # substitutions java:L257-L273
            for s in replace_deg(cur.seq, cur.posCur, self.L):
                if s != cur.seq and not self.already(s):
                    n = cur.copy()
                    n.nSub += 1
                    n.seq = s
                    r = self.check(n)
                    if r is not None:
                        self.matches.add(r)
                    if r is not None or self.doNext:
                        self.goNext(n)
            if self.allowIndels and cur.posCur < Lm1:
                # insertions java:L284-L300
                for s in insert_deg(cur.seq, cur.posCur, self.L):
                    if not self.already(s):
                        n = cur.copy()
                        n.seq = s
                        n.nDel += 1
                        r = self.check(n)
                        if r is not None:
                            self.matches.add(r)
                        if r is None or self.doNext:
                            self.goNext(n)
                # deletions java:L313-L357
                if not (self.post is not None and cur.nDel + 1 > len(self.post)):
                    last = self.post[cur.nDel] if self.post is not None else 0
                    m = delete_byte(cur.seq, last, cur.posCur, self.L)
                    cands = [m] if self.post is not None else [m, m | 1, m | 2, m | 3]
                    for s in cands:
                        if not self.already(s):
                            n = cur.copy()
                            n.seq = s
                            n.nIns += 1
                            r = self.check(n)
                            if r is not None:
                                self.matches.add(r)
                            if r is None or self.doNext:
                                self.goNext(n)
            self.mark(cur.seq)
        return self.matches


# ---- F!com/rw/nanoporereadscanner/analyzers/Parser.assignBarcode (java:L195-L315) ---------------------
def assign_barcode(read, adapterpos, search, ranks, ed_max, plusminus=2, three_prime=True, L=16):
    """`read` = stranded read string, `adapterpos` as the Java (1-based adapter end).  Returns a dict
    (assigned False => only ed / ed_second are meaningful) or raises JavaException."""
    def substring(b, e):
        if b < 0 or e > len(read) or b > e:
            raise JavaException("StringIndexOutOfBounds")
        return read[b:e]

    matches = JHashSet()
    probes = 0
    for off in sorted(range(-plusminus, plusminus + 1), key=abs):          # L198-L200 (stable)
        if three_prime:
            bcStart = adapterpos - L + off                                  # L206
            bcEnd = adapterpos - 1 + off                                    # L207
        else:
            bcStart = adapterpos + 1 + off                                  # L209
            bcEnd = adapterpos + L + off                                    # L210
        bc = pack(substring(bcStart - 1, bcEnd))                            # L214
        if three_prime:
            codes = []
            for ch in substring(bcStart - 5, bcStart):                      # L218
                c = ENCODE[ord(ch)] if ord(ch) < 254 else None
                if c is None:
                    raise JavaException("AIOOBE ENCODE_MATRIX")
                codes.append(c & 0xFF)
            post = []
            for c in reversed(codes):                                       # reverseComplement()
                if c > 15:
                    raise JavaException("AIOOBE ONEBYTE_REVERSECOMP_MATRIX")
                post.append(ONEBYTE_RC[c])
            bc = revcomp2(bc, L)                                            # L221
        else:
            post = []
            for ch in substring(bcEnd, bcEnd + 5):                          # L219
                if ord(ch) >= 254:
                    raise JavaException("AIOOBE ENCODE_MATRIX")
                c = ENCODE[ord(ch)]
                post.append(c if c >= 0 else 255)
        t = BarcodeMatchTester(bc, L, ed_max, False, True, search, off, post, True)   # L223-L238
        m = t.doJob()
        probes += t.probes
        for e in m:                                                         # addAll (L240)
            matches.add(e)
    res = dict(assigned=False, bc=0, ed=-1, ed_second=2147483647, offset=0, n_ins=0, n_del=0, n_sub=0, rank=-1,
               probes=probes, tie_unpinned=matches.treeified)
    lst = list(matches)
    if not lst:                                                             # L244
        return res
    lst.sort(key=OneMatch.key)                                              # stable (L247)
    best = lst[0]
    second = None
    for e in lst[1:]:
        if e.bc != best.bc:                                                 # distinctByKey(matchingBC)
            second = e
            break
    res["ed"] = best.ed
    res["ed_second"] = second.ed if second is not None else 2147483647
    if best.ed <= ed_max and (second is None or best.ed < second.ed):      # L251-L252
        res.update(assigned=True, bc=best.bc, offset=best.offset, n_ins=best.nIns, n_del=best.nDel, n_sub=best.nSub,
                   rank=ranks.get(best.bc, -1) if ranks is not None else -1)
        # L273-L280
        if three_prime:
            start = adapterpos - 1 + best.offset
            end = start - (L - 1) - (best.nIns - best.nDel)
        else:
            start = adapterpos + 1 + best.offset
            end = start + (L - 1) + (best.nIns - best.nDel)
        res["bcStart"], res["bcEnd"] = start, end
    return res


# ---- UMI: apachemod/LevenshteinDistance.limitedCompare (java:L220-L283) --------------------------------
IMAX = 2147483647


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def limited_compare(left, right, threshold):
    n, m = len(left), len(right)
    p = [0] * (n + 1)
    d = [0] * (n + 1)
    boundary = threshold + 1
    for i in range(boundary):
        p[i] = i
    for i in range(boundary, n + 1):
        p[i] = IMAX
    for i in range(n + 1):
        d[i] = IMAX
    for j in range(1, m + 1):
        rj = right[j - 1]
        d[0] = j
        mn = max(1, j - threshold)
        mx = n if j > IMAX - threshold else min(n, j + threshold)
        if mn > 1:
            d[mn - 1] = IMAX
        lower = IMAX
        for i in range(mn, mx + 1):
            if left[i - 1] == rj:
                d[i] = p[i - 1]
            else:
                d[i] = _i32(1 + min(min(d[i - 1], p[i]), p[i - 1]))
            lower = min(lower, d[i])
        if lower > threshold:
            return -1
        p, d = d, p
    return p[n] if p[n] <= threshold else -1


def umi_best9(a, b, umi_len=12):
    """ClusteringEditDistanceBase.calcEditDistances + calcBestEditDistance (java:L297-L350, L67-L80)."""
    eds = [[0] * 3 for _ in range(3)]
    for i in (-1, 0, 1):
        s1 = list(a[1 + i:1 + i + umi_len])
        for j in (-1, 0, 1):
            s2 = list(b[1 + j:1 + j + umi_len])
            if s1 == s2:
                eds[i + 1][j + 1] = 0
            else:
                dd = limited_compare(s1, s2, 4)
                eds[i + 1][j + 1] = 5 if dd == -1 else dd
    best = (127, 0, 0)
    for i in (1, 2, 0):                                       # EnumSet order ZERO, PLUSONE, MINUSONE (getValue 1,2,0)
        for v in (1, 2, 0):
            if eds[i][v] < best[0]:
                best = (eds[i][v], i, v)
    return _i32((best[0] & 0xFFFFFF) | (0x08000000 << best[1]) | (0x01000000 << best[2]))


# =====================================================================================================
# Illumina-guided search: BCUMIEDtesterBase + UMInuc/BCnuc testers + the consumers' reduction.
# Written object-for-object: one Python object per Java object, the matchingList holds the very objects the Java
# list holds (so the aliasing of the root node and the flag inheritance through the copy constructor are real).
# =====================================================================================================
FLAG_GENE, FLAG_ALL, FLAG_EMPTY = 1, 2, 4        # stand-ins for the BarcodeFindingFlag bits the tester sets
FOURBIT_TO_TWOBIT = [0] * 15                     # NucleicAcidInmutableOneBytePerBase.java:L29-L37
for _ch in "AGCT":
    FOURBIT_TO_TWOBIT[ENCODE[ord(_ch)]] = BASE_TO_TWOBIT[ord(_ch)]


class GuidedNode:
    """LongSeqMutated incl. the NucTwoBitPerBaseWithErrors fields the guided mode uses (findingErrorFlag)."""
    __slots__ = ("seq", "L", "nSub", "nIns", "nDel", "offset", "posPrev", "posCur", "level", "flag")

    def __init__(self, seq, L, level, offset):                # LongSeqMutated.java:L61 (nDeletions 0, pos -1/-1)
        self.seq, self.L, self.offset, self.level = seq, L, offset, level
        self.nSub = self.nIns = self.nDel = 0
        self.posPrev = self.posCur = -1
        self.flag = 0

    def copy(self):                                           # LongSeqMutated.java:L68-L77 + NucTwoBitPerBaseWithErrors.java:L46-L56
        c = GuidedNode(self.seq, self.L, self.level, self.offset)
        c.nSub, c.nIns, c.nDel = self.nSub, self.nIns, self.nDel
        c.posPrev, c.posCur = self.posPrev, self.posCur
        c.flag = self.flag                                    # L55: findingErrorFlag is copied
        return c

    def n_errors(self):                                       # getNErrors java:L100
        return self.nDel + self.nIns + self.nSub


class GuidedTester:
    """BCUMIEDtesterBase with either checkMatchWithTestSets flavour.
    umis: set (UMI flavour) — or gene/all/empty sets (BC flavour, gene may be None)."""

    def __init__(self, ed, L, post4, bailout=None, umis=None, gene=None, all_bcs=None, all_ed=0, empty=None, empty_ed=0,
                 bc_flavour=False, allow_indels=True):
        self.ed, self.L, self.post4, self.bailout, self.allow_indels = ed, L, post4, bailout, allow_indels
        self.umis, self.gene, self.all_bcs, self.all_ed, self.empty, self.empty_ed = umis, gene, all_bcs, all_ed, empty, empty_ed
        self.bc_flavour = bc_flavour
        self.matching = []                                    # BCUMIEDtesterBase.java:L53
        self.deque = deque()                                  # NucTwoBitPerBaseEDtesterBase.java:L62
        # ctor L82-L95: Optional<Boolean> use64bitHash
        if L >= 14 and ed >= 2:
            self.use64 = L > 16
        elif L < 14 and ed >= 2:
            self.use64 = False
        else:
            self.use64 = None
        self.tested = set()
        self.probes = 0

    def add_tested(self, seq):                                # L105-L112
        if self.use64 is not None:
            self.tested.add(seq if self.use64 else seq & 0xFFFFFFFF)

    def already_tested(self, seq):                            # L120
        if self.use64 is None:
            return False
        return (seq if self.use64 else seq & 0xFFFFFFFF) in self.tested

    def check(self, node):
        if not self.bc_flavour:                               # UMInucTwoBitPerBaseEDtester.java:L60-L62
            self.probes += 1
            return node if node.seq in self.umis else None
        # BCnucTwoBitPerBaseEDtester.java:L72-L92
        if self.gene is not None:
            self.probes += 1
            if node.seq in self.gene:
                node.flag |= FLAG_GENE                        # L76: on the node itself
                return node
        if self.all_bcs is not None and node.level <= self.all_ed:
            self.probes += 1
            if node.seq in self.all_bcs:
                es = node.copy()                              # L80: new NucTwoBitPerBaseWithErrors(seq)
                es.flag |= FLAG_ALL
                return es
        if self.empty is not None and node.level <= self.empty_ed:
            self.probes += 1
            if node.seq in self.empty:
                es = node.copy()
                es.flag |= FLAG_EMPTY
                return es
        return None

    def go_next(self, node):                                  # NucTwoBitPerBaseEDtesterBase.java:L133-L144
        if self.ed > node.level:
            if self.bailout is None or node.level < self.bailout or not self.matching:
                d = node.copy()
                d.posPrev = node.posCur
                d.posCur = -1
                d.level = node.level + 1
                self.deque.append(d)

    def run(self, seq, offset):                               # matchesSeqEditDistance BCUMIEDtesterBase.java:L82-L124
        L = self.L
        parent = GuidedNode(seq, L, 1, offset)                # L82: (seq, nDeletions 0, currentlevel 1, offset)
        r = self.check(parent)
        if r is not None:
            self.matching.append(r)
        if self.ed == 0:
            return self.matching
        self.deque.append(parent)
        while self.deque:
            cur = self.deque.pop()                            # pollLast
            cur.posCur += 1
            if cur.posCur < L - 1:
                self.deque.append(cur.copy())
            if cur.posPrev == cur.posCur:
                continue
            for s in replace_deg(cur.seq, cur.posCur, L):     # substitutions L136-L151
                if s != cur.seq and not self.already_tested(s):
                    m = cur.copy()
                    m.nSub += 1
                    m.seq = s
                    r = self.check(m)
                    if r is not None:
                        self.matching.append(r)
                    self.go_next(m)
            if self.allow_indels:
                if cur.posCur < L - 1:                        # insertions L161-L175
                    for s in insert_deg(cur.seq, cur.posCur, L):
                        if not self.already_tested(s):
                            m = cur.copy()
                            m.seq = s
                            m.nDel += 1
                            r = self.check(m)
                            if r is not None:
                                self.matching.append(r)
                            self.go_next(m)
                if cur.nDel + 1 <= len(self.post4):           # deletions L187-L203
                    code = self.post4[cur.nDel]               # getByteAt(nDeletions+1)
                    s = delete_byte(cur.seq, code, cur.posCur, L)
                    if not self.already_tested(s):
                        m = cur.copy()
                        m.seq = s
                        m.nIns += 1
                        r = self.check(m)
                        if r is not None:
                            self.matching.append(r)
                        self.go_next(m)
            self.add_tested(cur.seq)
        return self.matching


def score_where_found(flag):                                  # BarcodeFindingFlag.java:L119-L131
    if flag & FLAG_GENE:
        return 3
    if flag & FLAG_ALL:
        return 2
    if flag & FLAG_EMPTY:
        return 1
    return 0


def guided_query(slice_ascii, anchor, L, ed, plusminus, post_len, bailout=None, bc_flavour=False, **sets):
    """Offset loop of findUMI / testBarcodes + sorted().distinct() of getBestAndSecondBCorUMI.
    Returns (raw list of dicts, sorted distinct list of dicts).  Raises JavaException like the Java would."""
    raw = []
    for i in sorted(range(-plusminus, plusminus + 1), key=abs):          # stable: 0,-1,1,-2,2
        ws = anchor + i
        if ws < 0 or ws + L + post_len > len(slice_ascii):
            raise JavaException("getSubSequence out of range")
        codes = []
        for ch in slice_ascii[ws:ws + L + post_len]:
            c = ENCODE[ch] if ch < 254 else -1
            if c < 0:
                raise JavaException("unknown base")
            codes.append(c)
        w = 0
        for c in codes[:L]:                                              # getLongHashForBytes T!...java:L197-L201
            if c >= 15:
                raise JavaException("AIOOBE FOURBIT_TO_TWOBIT_MATRIX")
            w = (_shl(w, 2) | FOURBIT_TO_TWOBIT[c]) & M64
        t = GuidedTester(ed, L, codes[L:], bailout=bailout, bc_flavour=bc_flavour, **sets)
        raw.extend(t.run(w, i))
    as_dict = lambda n: dict(seq=n.seq, n_sub=n.nSub, n_ins=n.nIns, n_del=n.nDel, offset=n.offset, where=n.flag, level=n.level)
    lst = raw
    if len(raw) > 1:
        if bc_flavour:
            key = lambda n: (n.n_errors(), score_where_found(n.flag), abs(n.offset))
        else:
            key = lambda n: (n.n_errors(), abs(n.offset))
        lst, seen = [], set()
        for n in sorted(raw, key=key):                                   # list.sort / Stream.sorted are stable, like Python's
            if n.seq not in seen:
                seen.add(n.seq)
                lst.append(n)
    return [as_dict(n) for n in raw], [as_dict(n) for n in lst]


# ---- neighbour-set clustering: ClusterOne_MyClustering.clusterLocal (ClusterOne_MyClustering.java:L175-L219) -----------
def _i8(x):
    x &= 0xFF
    return x - 256 if x >= 128 else x


def fastutil_mix(x):
    """it.unimi.dsi.fastutil.HashCommon.mix(int) (fastutil 8.2.2, jar absent from the mount: restated from the
    published source): h = x * 0x9E3779B9; h ^ (h >>> 16)."""
    h = (x * 0x9E3779B9) & 0xFFFFFFFF
    return h ^ (h >> 16)


def fastutil_key_order(keys_in_insertion_order):
    """Iteration order of an Int2ObjectOpenHashMap (fastutil 8.2.2 defaults: 16 expected, load factor .75) filled by
    put() in the given order: key 0 first (it lives outside the table), then table slots from the last to the first.
    Restated from the published source, NOT pinned (the jar is a missing blob); the tests use it only as one example
    of an order the caller may pass as `rank`."""
    n, size, has_zero = 32, 0, False
    table = [0] * n

    def place(tab, mask, k):
        pos = fastutil_mix(k) & mask
        while tab[pos] != 0:
            pos = (pos + 1) & mask
        tab[pos] = k

    for k in keys_in_insertion_order:
        if k == 0:
            if has_zero:
                continue
            has_zero = True
        else:
            pos = fastutil_mix(k) & (n - 1)
            dup = False
            while table[pos] != 0:
                if table[pos] == k:
                    dup = True
                    break
                pos = (pos + 1) & (n - 1)
            if dup:
                continue
            table[pos] = k
        size += 1
        if size - 1 >= int(n * 0.75):          # if (size++ >= maxFill) rehash(arraySize(size + 1, f))
            need = -(-(size + 1) * 4 // 3)      # ceil((size + 1) / .75)
            new_n = 2
            while new_n < need:
                new_n *= 2
            new = [0] * new_n
            for i in range(n - 1, -1, -1):      # rehash walks the old table downwards
                if table[i] != 0:
                    place(new, new_n - 1, table[i])
            table, n = new, new_n
    order = [0] if has_zero else []
    order += [table[i] for i in range(n - 1, -1, -1) if table[i] != 0]
    return order


def cluster_local(matrix, indices, ed, key_order=None):
    """clusterLocal, stream by stream.  matrix[a][v] = packed BestEditDistance ints; indices = the Collection<Integer>;
    key_order(keys) -> iteration order of possibleClusters (default: ascending).  Returns None (Optional.empty) or the
    set of clusters (frozensets of indices)."""
    indices = list(indices)
    possible = {}
    for a in indices:                                                            # L179-L185
        s = {v for v in indices if _i8(matrix[a][v] & 0xFFFFFF) <= ed}
        if len(s) > 1 and a not in possible:                                     # merge function keeps the first
            possible[a] = s
    order = list(key_order(list(possible))) if key_order else sorted(possible)
    id_map = {}
    for c in order:                                                              # L190-L199
        best = None
        for l in order:
            if c in possible[l]:
                if best is None or not (len(possible[best]) >= len(possible[l])):   # compare(a, b) >= 0 ? a : b
                    best = l
        id_map.setdefault(best, set()).add(c)
    if not id_map:                                                               # L219
        return None
    return {frozenset(v) for v in id_map.values()}


# ---- ClusterOneHierarchical.call: LingPipe complete / single link + centres + per-read values ------------------------------------------
# Second, object-for-object restatement (the C oracle works on index arrays): F!…/clustering/ClusterOneHierarchical.java:L61-L217,
# A!com/aliasi/cluster/{CompleteLinkClusterer (L146-L237), SingleLinkClusterer (L198-L268), Dendrogram (L205-L215), LinkDendrogram},
# A!com/aliasi/util/{BoundedPriorityQueue (L131-L153, L342-L346, L458-L464), ObjectToSet}, F!com/rw/clustering/{DistanceMatrix, OneUmiCluster}.
class _JInt:
    """java.lang.Integer as a HashSet element"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v)

    def jhash(self):
        return self.v

    def jequals(self, o):
        return self.v == o.v


def jdk_int_set(values_in_insertion_order):
    """iteration order of a java.util.HashSet<Integer> filled by add() in the given order"""
    hs = JHashSet()
    for v in values_in_insertion_order:
        hs.add(_JInt(v))
    return [e.v for e in hs]


def fastutil_intset_order(keys_in_insertion_order):
    """fastutil IntOpenHashSet (OneUmiCluster's superclass): same open-addressing layout as Int2ObjectOpenHashMap"""
    return fastutil_key_order(keys_in_insertion_order)


class _Dendro:
    def __init__(self, left=None, right=None, score=0.0, obj=None):
        self.left, self.right, self.score, self.obj, self.parent = left, right, float(score), obj, None
        if left is not None:
            left.parent = self
            right.parent = self

    def dereference(self):                                     # Dendrogram.dereference
        d = self
        while d.parent is not None:
            d = d.parent
        return d

    def member_list(self):                                     # addMembers: dendrogram1 first, then dendrogram2; into a HashSet
        if self.left is None:
            return [self.obj]
        return self.left.member_list() + self.right.member_list()

    def partition_distance(self, max_distance):                # Dendrogram.java:L205-L215 (stack via addFirst / removeFirst)
        out, stack = [], [self]
        while stack:
            cur = stack.pop(0)
            if cur.score <= max_distance:
                out.append(jdk_int_set(cur.member_list()))     # memberSet(): new HashSet + addMembers
            elif cur.left is not None:
                stack.insert(0, cur.left)
                stack.insert(0, cur.right)
        return out


class _Pair:
    __slots__ = ("d1", "d2", "score", "entry_id", "step")

    def __init__(self, d1, d2, score):
        self.d1, self.d2, self.score, self.entry_id, self.step = d1, d2, float(score), None, 0


class _Queue:
    """BoundedPriorityQueue(ScoredObject.reverseComparator(), MAX): a TreeSet ordered by EntryComparator — ascending score, among equal
    scores the entry offered LAST comes first"""

    def __init__(self):
        self.items, self.next_id = [], 0

    def offer(self, p):
        p.entry_id = self.next_id
        self.next_id += 1
        self.items.append(p)

    def poll(self):
        if not self.items:
            return None
        best = min(self.items, key=lambda p: (p.score, -p.entry_id))
        self.items.remove(best)
        return best

    def remove_all(self, ps):
        s = set(id(p) for p in ps)
        self.items = [p for p in self.items if id(p) not in s]


def complete_link(m, distance, max_distance=float("inf"), stop_above=None):
    """CompleteLinkClusterer.hierarchicalCluster over elements 0..m-1.  Returns (root or forest roots, tie_seen).  ObjectToSet's HashSet<PairScore>
    iterates in identity-hash order in the JVM; here: insertion order (Python dicts), the canonical order of oracle and GPU kernel.
    stop_above: stop once the cheapest pair costs more (the links above the cut do not change partitionDistance(stop_above))."""
    leafs = [_Dendro(obj=i) for i in range(m)]
    if m == 1:
        return [leafs[0]], False
    queue, index = _Queue(), {}
    for i in range(m):
        for j in range(i + 1, m):
            ps = _Pair(leafs[i], leafs[j], distance(i, j))
            queue.offer(ps)
            index.setdefault(leafs[i], {})[ps] = None
            index.setdefault(leafs[j], {})[ps] = None
    tie_seen, step, last = False, 0, None
    while queue.items:
        nxt = queue.poll()
        if stop_above is not None and nxt.score > stop_above:
            queue.items.append(nxt)
            break
        if nxt.step > 0 and any(p.score == nxt.score and p.step == nxt.step for p in queue.items):
            tie_seen = True
        step += 1
        d1, d2 = nxt.d1.dereference(), nxt.d2.dereference()
        d12 = _Dendro(d1, d2, nxt.score)
        last = d12
        buf = {}
        ps3set = index.pop(d1)
        queue.remove_all(ps3set)
        for ps3 in ps3set:
            d3 = ps3.d2 if ps3.d1 is d1 else ps3.d1
            index[d3].pop(ps3, None)
            buf[d3] = ps3.score
        ps3set = index.pop(d2)
        queue.remove_all(ps3set)
        for ps3 in ps3set:
            d3 = ps3.d2 if ps3.d1 is d2 else ps3.d1
            index[d3].pop(ps3, None)
            if d3 not in buf:
                continue
            ps = _Pair(d12, d3, max(buf[d3], ps3.score))
            ps.step = step
            queue.offer(ps)
            index.setdefault(d12, {})[ps] = None
            index.setdefault(d3, {})[ps] = None
        if not queue.items:
            return [d12], tie_seen
    roots = [d for d in leafs if d.parent is None]
    seen = set()
    for d in leafs:
        r = d.dereference()
        if id(r) not in seen:
            seen.add(id(r))
            if r.left is not None:
                roots.append(r)
    return roots, tie_seen


def single_link(m, distance, max_distance):
    leafs = [_Dendro(obj=i) for i in range(m)]
    pairs = [(float(distance(i, j)), i, j) for i in range(m) for j in range(i + 1, m)]
    pairs.sort(key=lambda t: t[0])                              # stable
    clusters = m
    for sc, i, j in pairs:
        if clusters <= 1 or sc > max_distance:
            break
        d1, d2 = leafs[i].dereference(), leafs[j].dereference()
        if d1 is d2:
            continue
        _Dendro(d1, d2, sc)
        clusters -= 1
    roots, seen = [], set()
    for d in leafs:
        r = d.dereference()
        if id(r) not in seen:
            seen.add(id(r))
            roots.append(r)
    return roots


def assign_hier(matrix, ed_complete=2, ed_single=1, single_threshold=3000, fold_depth=50, qv01=False):
    """ClusterOneHierarchical.call on one job's packed matrix (list of lists / 2-D array).  Returns one dict per read:
    center (-1 = none), u1, u2 (-1 = absent), pos2, off_mean, assigned, skipped, cluster_size, n_clusters, tie_unpin."""
    n = len(matrix)
    ED = lambda a, b: _i8(int(matrix[a][b]) & 0xFFFFFF)
    rec = [dict(center=-1, u1=0, u2=-1, pos2=0, off_mean=0, assigned=False, skipped=False, cluster_size=0, n_clusters=0, tie_unpin=False)
           for _ in range(n)]
    iwn = [i for i in range(n) if sum(1 for j in range(n) if i != j and ED(i, j) <= ed_complete) > 0]     # DistanceMatrix.java:L87-L90
    if len(iwn) <= 1:
        return rec
    single = len(iwn) > single_threshold
    cut = ed_single if single else ed_complete
    dist = lambda a, b: float(ED(iwn[a], iwn[b]))
    m = len(iwn)
    if single:
        roots, tie_seen = single_link(m, dist, float(cut)), False
    else:
        roots, tie_seen = complete_link(m, dist, stop_above=float(cut))
    reduced = []
    for r in roots:
        reduced += r.partition_distance(float(cut))
    reduced = [c for c in reduced if len(c) > 1]                                                         # L101
    chain_dep = False
    full = []
    for c in reduced:                                                                                    # DistanceMatrix.java:L145
        b = jdk_int_set([iwn[k] for k in c])
        cap = 16
        while len(c) > cap * 3 // 4:
            cap *= 2
        if len({k & (cap - 1) for k in c}) < len(c) or len({k & (cap - 1) for k in b}) < len(b):
            chain_dep = True
        if len(b) > 1:
            full.append(b)
    if not full:
        return rec
    thr = lambda a, b: ED(iwn[a], iwn[b]) <= cut
    cluster_graph = all((thr(a, c) == thr(b, c)) for a in range(m) for b in range(m) if a != b and thr(a, b) for c in range(m) if c != a and c != b)
    unpin = tie_seen and (not cluster_graph or chain_dep)
    maxdepth = max(len(c) for c in full)
    cluster_list = []
    for c in full:
        if len(c) * fold_depth > maxdepth:
            cluster_list.append(fastutil_intset_order(c))                                                 # OneUmiCluster
        else:
            for x in c:
                rec[x]["skipped"], rec[x]["cluster_size"] = True, len(c)
    for it in cluster_list:
        k = len(it)
        if k == 2:
            center = it[0] if qv01 else it[1]
        else:
            sums = [(sum(int(float(ED(s, w)) ** 2.0) for w in it if w != s), i) for i, s in enumerate(it)]
            center = it[min(sums)[1]]                                                                    # stable sort, first
        offs = [(-1 if matrix[center][v] & 0x08000000 else 0 if matrix[center][v] & 0x10000000 else 1 if matrix[center][v] & 0x20000000 else 0)
                for v in it if v != center]
        import math
        off_mean = int(math.floor(sum(offs) / len(offs) + 0.5))
        for x in it:
            r = rec[x]
            r.update(center=center, assigned=True, cluster_size=k, off_mean=off_mean, u1=ED(center, x))
            p = int(matrix[center][x])
            r["pos2"] = 0 if p & 0x01000000 else 1 if p & 0x02000000 else 2 if p & 0x04000000 else 1
            if len(cluster_list) > 1:
                outside = [ED(x, y) for y in range(n) if y not in it]
                if outside:
                    r["u2"] = min(outside)
    for r in rec:
        r["n_clusters"], r["tie_unpin"] = len(cluster_list), unpin
    return rec


# ---- F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering (jobs of more than 100 reads) ------------------------------------------
# Every stream of this class is parallel above 30 reads (ClusterOne_MyClustering.java:L176-L177, L187-L189); the restatement follows the SEQUENTIAL
# semantics (what a JVM with one worker thread produces).  Containers whose order reaches the result are modelled explicitly.
class FuIntSet:
    """fastutil 8.2.2 IntOpenHashSet as OneUmiCluster uses it: add, iteration, and java.util.AbstractCollection.removeAll(Collection) — the set's
    iterator walks the slots downwards and removes through SetIterator.remove (backward-shift deletion, entries that wrap around the table end are
    remembered in `wrapped` and visited last).  Restated from the published source (the jar is not in the mount)."""

    def __init__(self):
        self.n, self.size, self.has_zero = 32, 0, False
        self.key = [0] * 32
        self.min_n = 32

    def _max_fill(self):
        return min(int(-(-self.n * 3 // 4)), self.n - 1)                      # HashCommon.maxFill(n, .75f)

    def add(self, k):
        if k == 0:
            if self.has_zero:
                return False
            self.has_zero = True
        else:
            mask = self.n - 1
            pos = fastutil_mix(k) & mask
            while self.key[pos] != 0:
                if self.key[pos] == k:
                    return False
                pos = (pos + 1) & mask
            self.key[pos] = k
        self.size += 1
        if self.size - 1 >= self._max_fill():                                 # if (size++ >= maxFill) rehash(arraySize(size + 1, f))
            need = -(-(self.size + 1) * 4 // 3)
            nn = 2
            while nn < need:
                nn *= 2
            self._rehash(nn)
        return True

    def _rehash(self, nn):
        new = [0] * nn
        for i in range(self.n - 1, -1, -1):
            if self.key[i] != 0:
                pos = fastutil_mix(self.key[i]) & (nn - 1)
                while new[pos] != 0:
                    pos = (pos + 1) & (nn - 1)
                new[pos] = self.key[i]
        self.key, self.n = new, nn

    def order(self):
        return ([0] if self.has_zero else []) + [self.key[i] for i in range(self.n - 1, -1, -1) if self.key[i] != 0]

    def __len__(self):
        return self.size

    def _shift(self, pos, wrapped=None):
        key, mask = self.key, self.n - 1
        while True:
            last = pos
            pos = (pos + 1) & mask
            while True:
                curr = key[pos]
                if curr == 0:
                    key[last] = 0
                    return
                slot = fastutil_mix(curr) & mask
                if (last >= slot or slot > pos) if last <= pos else (last >= slot and slot > pos):
                    break
                pos = (pos + 1) & mask
            if wrapped is not None and pos < last:
                wrapped.append(key[pos])
            key[last] = curr

    def remove(self, k):                                                       # IntOpenHashSet.remove(int)
        if k == 0:
            if not self.has_zero:
                return False
            self.has_zero = False
            self.size -= 1
        else:
            mask = self.n - 1
            pos = fastutil_mix(k) & mask
            while self.key[pos] != k:
                if self.key[pos] == 0:
                    return False
                pos = (pos + 1) & mask
            self.size -= 1
            self._shift(pos)
        if self.n > self.min_n and self.size < self._max_fill() // 4 and self.n > 16:
            self._rehash(self.n // 2)
        return True

    def remove_all(self, victims):                                             # AbstractCollection.removeAll: iterate THIS, it.remove() on a hit
        victims = set(victims)
        pos, c, must_null, wrapped = self.n, self.size, self.has_zero, []
        while c != 0:
            c -= 1
            if must_null:
                must_null = False
                if 0 in victims:
                    self.has_zero = False
                    self.size -= 1
                continue
            cur = None
            while cur is None:
                pos -= 1
                if pos < 0:
                    cur = wrapped[-pos - 1]
                    if cur in victims:
                        self.remove(cur)                                       # "we're removing wrapped entries": the set's own remove
                    break
                if self.key[pos] != 0:
                    cur = self.key[pos]
                    if cur in victims:
                        self._shift(pos, wrapped)
                        self.size -= 1


class _JSetOfInts:
    """a java.util.Set<Integer> as an element of another HashSet: AbstractSet.hashCode = sum of the elements, equals by content"""

    def __init__(self, items):
        self.items = list(items)
        self.fs = frozenset(items)

    def jhash(self):
        s = sum(self.items) & 0xFFFFFFFF
        return s

    def jequals(self, o):
        return self.fs == o.fs


def chm_key_order(keys_in_insertion_order):
    """iteration order of a java.util.concurrent.ConcurrentHashMap<Integer, ?> filled by computeIfAbsent in the given order by ONE thread:
    table of 16, doubled when the count reaches .75 of it (addCount), bins are chains in insertion order, a transfer keeps the last run of a
    bin and PREPENDS the nodes before it (so their order reverses).  Second value: a bin reached the treeify threshold (not modelled)."""
    cap, sc, count, long_bin = 16, 12, 0, False
    table = [[] for _ in range(cap)]
    spread = lambda h: (h ^ (h >> 16)) & 0x7FFFFFFF
    for k in keys_in_insertion_order:
        h = spread(k & 0xFFFFFFFF)
        b = table[h & (cap - 1)]
        if any(x == k for (_, x) in b):
            continue
        if len(b) >= 8:
            long_bin = True
        b.append((h, k))
        count += 1
        while count >= sc:
            new = [[] for _ in range(2 * cap)]
            for i, chain in enumerate(table):
                if not chain:
                    continue
                run_bit, last_run = chain[0][0] & cap, 0
                for j in range(1, len(chain)):
                    bb = chain[j][0] & cap
                    if bb != run_bit:
                        run_bit, last_run = bb, j
                lo = list(chain[last_run:]) if run_bit == 0 else []
                hi = list(chain[last_run:]) if run_bit != 0 else []
                for j in range(last_run):
                    if chain[j][0] & cap == 0:
                        lo.insert(0, chain[j])
                    else:
                        hi.insert(0, chain[j])
                new[i], new[i + cap] = lo, hi
            table, cap = new, 2 * cap
            sc = cap - (cap >> 2)
    return [k for chain in table for (_, k) in chain], long_bin


def _my_cluster_local(indices, ED, ed):
    """ClusterOne_MyClustering.clusterLocal (…java:L175-L219).  Returns (list of clusters as HashSet<Integer> iteration orders, in the iteration
    order of the resulting HashSet<Set<Integer>>; harmful_tie; long_bin)."""
    nbr = {}
    for a in indices:                                                          # L179-L184
        s = [v for v in indices if ED(a, v) <= ed]
        if len(s) > 1:
            nbr[a] = s
    keys = fastutil_key_order(list(nbr.keys()))                                # Int2ObjectOpenHashMap filled in the order of `indices` (L185)
    nset = {a: frozenset(s) for a, s in nbr.items()}
    chosen, harmful = {}, False
    for c in keys:                                                             # L190-L196: Stream.max keeps the FIRST of equal maxima
        best = None
        for e in keys:
            if c in nset[e]:
                if best is None or len(nset[e]) > len(nset[best]):
                    best = e
        for e in keys:
            if c in nset[e] and len(nset[e]) == len(nset[best]) and nset[e] != nset[best]:
                harmful = True
        chosen[c] = best
    groups, first_seen = {}, []
    for c in keys:                                                             # L199: groupingByConcurrent(right, mapping(left, toSet()))
        e = chosen[c]
        if e not in groups:
            groups[e] = JHashSet()
            first_seen.append(e)
        groups[e].add(_JInt(c))
    order, long_bin = chm_key_order(first_seen)
    outer = JHashSet()                                                         # L219: idMap.values().stream().collect(toSet())
    for e in order:
        long_bin |= groups[e].treeified
        outer.add(_JSetOfInts([x.v for x in groups[e]]))
    long_bin |= outer.treeified
    return [s.items for s in outer], harmful, long_bin


def assign_myclust(matrix, ed=2, fold_depth=50, qv01=False):
    """ClusterOne_MyClustering.call (…java:L59-L166) on one job's packed matrix, sequential-stream semantics.  Same records as assign_hier."""
    import math
    n = len(matrix)
    ED = lambda a, b: _i8(int(matrix[a][b]) & 0xFFFFFF)
    rec = [dict(center=-1, u1=0, u2=-1, pos2=0, off_mean=0, assigned=False, skipped=False, cluster_size=0, n_clusters=0, tie_unpin=False)
           for _ in range(n)]

    def make_cluster(members):                                                 # toCollection(OneUmiCluster::new) + setClusterCenter (L86-L87)
        s = FuIntSet()
        for x in members:
            s.add(x)
        return dict(set=s, center=center_of(s))

    def center_of(s):                                                          # OneUmiCluster.setClusterCenterNotPreGrouped (OneUmiCluster.java:L49-L65)
        it = s.order()
        if len(it) == 2:
            return it[0] if qv01 else it[1]
        sums = [sum(int(float(ED(a, w)) ** 2.0) for w in it if w != a) for a in it]
        return it[sums.index(min(sums))]                                       # sorted() is stable, findFirst

    full, harmful, long_bin = _my_cluster_local(list(range(n)), ED, ed)        # L72-L73
    if not full:                                                               # Optional.empty (L219): idMap is empty
        return rec
    maxdepth = max(len(c) for c in full)                                       # L77
    cluster_list = []
    for c in full:                                                             # L78-L88
        if len(c) * fold_depth > maxdepth:
            cluster_list.append(make_cluster(c))
        else:
            for x in c:
                rec[x]["skipped"], rec[x]["cluster_size"] = True, len(c)
    clustered = set(x for cl in cluster_list for x in cl["set"].order())       # L90
    unclustered = [d for d in range(n) if d not in clustered]                  # L91
    removed = []
    for cl in cluster_list:                                                    # L102, L60-L65: removeOffCenter
        rem = [s_ for s_ in cl["set"].order() if ED(s_, cl["center"]) > ed]
        if rem:                                                                # OneUmiCluster.removeEntries (L114-L119)
            cl["set"].remove_all(rem)
            cl["center"] = center_of(cl["set"])
        removed += rem
    unclustered += removed                                                     # L104
    if removed:                                                                # L106-L112
        extra, h2, l2 = _my_cluster_local(unclustered, ED, ed)
        harmful |= h2
        long_bin |= l2
        for c in extra:
            if len(c) > 1:
                cluster_list.append(make_cluster(c))
    for cl in cluster_list:                                                    # L116-L164
        it = cl["set"].order()
        if len(it) <= 1:
            continue
        center = cl["center"]
        offs = [(-1 if matrix[center][v] & 0x08000000 else 0 if matrix[center][v] & 0x10000000 else 1 if matrix[center][v] & 0x20000000 else 0)
                for v in it if v != center]
        off_mean = int(math.floor(sum(offs) / len(offs) + 0.5))
        filtered = [s_ for s_ in it if ED(s_, center) <= ed]
        if len(filtered) <= 1:
            continue
        inside = set(it)
        for x in filtered:
            r = rec[x]
            if r["skipped"]:                                                   # ClusterOneBase.java:L122-L123: UMI_CLUSTERING_SKIPPED_HIGHCOMPLEXITY set in round 1
                continue
            r.update(center=center, assigned=True, cluster_size=len(it), off_mean=off_mean, u1=ED(center, x))
            p = int(matrix[center][x])
            r["pos2"] = 0 if p & 0x01000000 else 1 if p & 0x02000000 else 2 if p & 0x04000000 else 1
            if len(cluster_list) > 1:
                outside = [ED(x, y) for y in range(n) if y not in inside]
                if outside:
                    r["u2"] = min(outside)
    for r in rec:
        r["n_clusters"], r["tie_unpin"] = len(cluster_list), bool(harmful or long_bin)
    return rec
