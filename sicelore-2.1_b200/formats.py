"""Wire / on-disk formats either side of the hot path, host side only: the read-name extension that carries the scan
and barcode results from `scanfastq` to `assignumis`, and BarcodesAssigned.tsv.  Mirrors of
  F!com/rw/nanoporereadscanner/readerwriter/FastqRecordExt.class   getRecordForWriting (FastqRecordExt.java:L209-L311),
                                                                   getScanDatFromReadName (L395-L493), getMeanQV (L56-L59)
  F!com/rw/nanoporereadscanner/stats/ParseStatsHtmlPrinter.class   writeAssignedTSV (ParseStatsHtmlPrinter.java:L294-L340)
  /root/reference/src/main/java/org/ipmc/sicelore/programs/SelectValidCellBarcode.java:57-80   (the TSV's consumer)
so that a Java-free driver can write what the reference writes from the GPU records (slr_bc_result + table counters)."""
from decimal import Decimal, ROUND_HALF_EVEN

import numpy as np

_BASES = "AGCT"                      # TWOBIT_TO_BASE_ARRAY: A=0 G=1 C=2 T=3


def unpack2bit(h, length=16):
    """NucleicAcidTwoBitPerBase.toString()"""
    h = int(h)
    return "".join(_BASES[(h >> (2 * (length - 1 - i))) & 3] for i in range(length))


class ReadNameTags:
    """ReadScannerParameters prefixes (config.xml:40-52, 66) + the fixed barcode tags (ReadScannerParameters ctor: bc= ed= ed_sec= bcStart=
    bcEnd=, barcodeRankPrefix rk=)"""
    paStartPrefix, paEndPrefix, adapterPosPrefix, tsoPosPrefix, seqPrefix, qvPrefix = "PS=", "PE=", "AE=", "T=", "X=", "Q="
    barcodeSeqPrefix, barcodeEdPrefix, barcodeEdSecondaryPrefix = "bc=", "ed=", "ed_sec="
    barcodeStartPrefix, barcodeEndPrefix, barcodeRankPrefix = "bcStart=", "bcEnd=", "rk="
    nbasesOfAdapterSeqInReadname = 3
    N_BASES_AFTER_ADAPTER_FOR_SEQ = 40


def _dec_format_1(x):
    """DecimalFormat("##.#").format(Float): at most one fraction digit, HALF_EVEN, no trailing zero (FastqRecordExt.java:L36)"""
    d = Decimal(float(np.float32(x))).quantize(Decimal("0.1"), rounding=ROUND_HALF_EVEN)
    s = format(d, "f")
    return s[:-2] if s.endswith(".0") else s


def mean_qv(quals, start, stop):
    """getMeanQV (L56-L59): chars().skip(start - 1).limit(stop - start + 1).map(c - 33).average() as float"""
    q = quals[start - 1:start - 1 + (stop - start + 1)]
    return np.float32(sum(ord(c) - 33 for c in q) / len(q))


def convert_int(number):
    """FastqRecordExt$NumberToAndFromAscii.convertInt (L524): Integer.toString(number, 36)"""
    digits, n, out = "0123456789abcdefghijklmnopqrstuvwxyz", abs(int(number)), ""
    while True:
        out = digits[n % 36] + out
        n //= 36
        if n == 0:
            break
    return ("-" if number < 0 else "") + out


def convert_string(s):
    """NumberToAndFromAscii.convertString (L535): Integer.parseInt(s, 36); ValueError = NumberFormatException"""
    body = s[1:] if s[:1] in "+-" else s
    if not body or any(not (ch.isascii() and ch.isalnum()) for ch in body):
        raise ValueError("NumberFormatException: For input string: %r" % s)
    v = int(s, 36)
    if not -(1 << 31) <= v < (1 << 31):
        raise ValueError("NumberFormatException: For input string: %r" % s)
    return v


def read_name_extension(reversed_read, stranded_seq, stranded_quals, adapter_end=None, polya_start=None, polya_end=None, tso_end=None,
                        bc=None, ed=None, ed_second=None, bc_start=None, bc_end=None, rank=None, is5p=False, tags=ReadNameTags, bc_len=16,
                        read_id=None):
    """The `add` StringBuilder of getRecordForWriting for a read that passed (L226-L277); read_id (optional) is appended in base 36 (L272-L273).
    bc = 2-bit barcode (int) or string; None = no barcode found.  Positions are the 1-based positions on the stranded read."""
    add = "_REV" if reversed_read else "_FWD"
    add += "_"
    if polya_start is not None:
        add += "%s%d_" % (tags.paStartPrefix, polya_start)                       # L230
    if polya_start is not None and polya_end is not None:
        add += "%s%d_" % (tags.paEndPrefix, polya_end)                           # L231
    if adapter_end is not None:
        add += "%s%d_" % (tags.adapterPosPrefix, adapter_end)                    # L232
    if tso_end is not None:
        add += "%s%d_" % (tags.tsoPosPrefix, tso_end)                            # L233
    if bc is not None:                                                           # L234-L243
        bcs = bc if isinstance(bc, str) else unpack2bit(bc, bc_len)
        add += "%s%s_%s%d_" % (tags.barcodeSeqPrefix, bcs, tags.barcodeEdPrefix, ed)
        if ed_second is not None:
            add += "%s%d_" % (tags.barcodeEdSecondaryPrefix, ed_second)
        add += "%s%d_%s%d_" % (tags.barcodeStartPrefix, bc_start, tags.barcodeEndPrefix, bc_end)
        if rank is not None:
            add += "%s%d_" % (tags.barcodeRankPrefix, rank)
    # L247-L259: the extension is appended to the read name at L298 only on the path through the X= / Q= block — a read without adapter, or
    # whose adapter end is too close to the read start for the X= slice ("Beginrange inconsistent"), keeps its bare name
    if adapter_end is None:
        return ""
    if is5p:
        begin = adapter_end - tags.nbasesOfAdapterSeqInReadname                   # L250-L251
        end = adapter_end + tags.N_BASES_AFTER_ADAPTER_FOR_SEQ - 1
    else:
        begin = adapter_end - tags.N_BASES_AFTER_ADAPTER_FOR_SEQ - 1               # L253-L254
        end = adapter_end + tags.nbasesOfAdapterSeqInReadname - 1
    if begin < 0:
        return ""
    if begin == 0:
        # the range check lets begin == 0 through (L247: ifge), and getMeanQV then asks for IntStream.skip(begin - 1) = skip(-1) (L58): the
        # reference dies here with an IllegalArgumentException — a 3' read whose adapter ends at base 41, a 5' read whose adapter ends at base 3
        raise ValueError("IllegalArgumentException: -1 (IntStream.skip in getMeanQV, FastqRecordExt.java:L58)")
    if end > len(stranded_seq):
        raise IndexError("StringIndexOutOfBoundsException: substring(%d, %d)" % (begin, end))
    add += "%s%s_" % (tags.seqPrefix, stranded_seq[begin:end])                   # L261-L264
    add += "%s%s" % (tags.qvPrefix, _dec_format_1(mean_qv(stranded_quals, begin, end)))     # L270
    add += "_"                                                                   # L271
    if read_id is not None:
        add += convert_int(read_id)                                              # L272-L273
    if bc is not None:
        add += " cellBC=" + (bc if isinstance(bc, str) else unpack2bit(bc, bc_len))                # L276-L277
    return add


def _extract(s, tag):
    """lambda$getScanDatFromReadName$4 (L397-L408): text between the first occurrence of `tag` and the next '_'"""
    i = s.find(tag)
    if i < 0:
        return None
    i += len(tag)
    j = s.find("_", i)
    return s[i:j] if j >= 0 else s[i:]


def parse_read_name(name, tags=ReadNameTags, max_bc_ed=None):
    """getScanDatFromReadName (L411-L489).  Returns None when neither _REV_ nor _FWD_ is present (Optional.absent, L418); raises
    KeyError when the adapter tag is missing (AdapterInfoNotFoundInReadException, L442-L443) and ValueError when the text after the last
    '_' is not a base-36 number (NumberFormatException from the read-id parse, L492-L494: scanfastq always writes '_' after Q=).  max_bc_ed = the assignumis limit on the
    barcode edit distance: a larger ed= drops the whole barcode block (L450-L459)."""
    k = name.find("_REV_")
    if k >= 0:
        reverse = True
    else:
        k = name.find("_FWD_")
        if k < 0:
            return None
        reverse = False
    sub = name[k:]
    out = {"reversed": reverse}
    g = lambda tag: _extract(sub, tag)
    pa_s, pa_e, ae, tso = g(tags.paStartPrefix), g(tags.paEndPrefix), g(tags.adapterPosPrefix), g(tags.tsoPosPrefix)
    bc_ed, bc_ed2, bc_seq = g(tags.barcodeEdPrefix), g(tags.barcodeEdSecondaryPrefix), g(tags.barcodeSeqPrefix)
    bc_s, bc_e, rk, seq, qv = g(tags.barcodeStartPrefix), g(tags.barcodeEndPrefix), g(tags.barcodeRankPrefix), g(tags.seqPrefix), g(tags.qvPrefix)
    if pa_s is not None:
        out["polya_start"] = int(pa_s)
    if pa_e is not None:
        out["polya_end"] = int(pa_e)
    if ae is None:
        raise KeyError("AdapterInfoNotFoundInReadException")
    out["adapter_end"] = int(ae)
    if tso is not None:
        out["tso_end"] = int(tso)
    do_bc = False
    if bc_ed is not None:
        e = int(bc_ed)
        do_bc = max_bc_ed is None or e <= max_bc_ed
        if do_bc:
            out["ed"] = e
    if do_bc:
        if bc_seq is not None:
            out["bc"] = bc_seq
        if bc_ed2 is not None:
            out["ed_second"] = int(bc_ed2)
        if bc_s is not None:
            out["bc_start"] = int(bc_s)
        if bc_e is not None:
            out["bc_end"] = int(bc_e)
        if rk is not None:
            out["rank"] = int(rk)
    if seq is not None:
        out["seq"] = seq
    if qv is not None:
        out["mean_qv"] = float(np.float32(qv))
    last = sub.rfind("_")                                                        # L492-L494: whatever follows the last '_' is the base-36 read id
    if last < len(sub) - 1:
        out["read_id"] = convert_string(sub[last + 1:])                          # a name that does not end in '_' [+ id] throws, like the Java
    return out


def write_assigned_tsv(path, barcodes2bit, counts_ed, assign_ed, bc_len=16):
    """writeAssignedTSV: header `Barcode\\tn Reads with ED<=E match\\tED=0..\\tED=E`, then one line per barcode with at least one assigned
    read, sorted by descending read count (L309-L310), numbers with DecimalFormat("###,###,###,###") grouping (dec_FORMATTER).
    counts_ed: [n, 3] per-barcode counters of the device table (slr_bc_counts_read).  Equal counts: the reference's order is the
    ConcurrentHashMap's iteration order (unpinned); here the 2-bit value ascending."""
    counts_ed = np.asarray(counts_ed).reshape(len(barcodes2bit), -1)
    total = counts_ed[:, :assign_ed + 1].sum(axis=1)
    keys = np.asarray(barcodes2bit, dtype=np.uint64)
    idx = np.nonzero(total > 0)[0]
    idx = idx[np.lexsort((keys[idx], -total[idx]))]
    with open(path, "w") as f:
        f.write("Barcode\tn Reads with ED<=%d match" % assign_ed + "".join("\tED=%d" % i for i in range(assign_ed + 1)) + "\n")
        for i in idx:
            f.write(unpack2bit(keys[i], bc_len) + "\t" + format(int(total[i]), ",") +
                    "".join("\t" + (format(int(counts_ed[i, e]), ",") if counts_ed[i, e] else "0") for e in range(assign_ed + 1)) + "\n")
    return len(idx)


def read_assigned_tsv(path):
    """What SelectValidCellBarcode.java:57-72 reads back: (barcode, total, ed0, ed1) per line, header skipped, ',' removed"""
    rows = []
    with open(path) as f:
        f.readline()
        for line in f:
            tab = line.rstrip("\n").replace(",", "").split("\t")
            if len(tab) < 3:
                continue
            rows.append((tab[0], int(tab[1]), int(tab[2] or 0), int(tab[3] or 0) if len(tab) > 3 else 0))
    return rows
