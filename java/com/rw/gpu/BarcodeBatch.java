package com.rw.gpu;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * One ReadChunk worth of barcode searches: the packing and unpacking halves of the shadow Parser.call() of INTEGRATION.md section 2, free of any
 * jar type so that it compiles on its own (javac -d out java/com/rw/gpu/*.java).  The shadow Parser keeps one instance per worker thread:
 *
 *   batch.clear();
 *   for every read with adapterFound():  int i = batch.add(fq.getStrandedSeq(), adapterresult.getEnd());
 *   batch.run(ctx, table, bcEditDistance, testPlusMinusPos);
 *   for every such read:  if (batch.threw(i)) throw ...;  if (batch.assigned(i)) { br.setStart(batch.start(i)); ... }
 *
 * Replaces Parser.assignBarcode (F!com/rw/nanoporereadscanner/analyzers/Parser.class, Parser.java:L195-L315) for the whole chunk.
 */
public final class BarcodeBatch {
    public static final int SLICE = 32, REC = 32, FLANK = 8, BC_LEN = 16;
    private final boolean threePrime;
    private final int capacity;
    private final ByteBuffer slices, lens, anchor, out;
    private final int[] adapterPos;
    private int n;

    public BarcodeBatch(int capacity, boolean threePrime) {
        this.capacity = capacity;
        this.threePrime = threePrime;
        slices = direct(capacity * SLICE);
        lens = direct(capacity * 4);
        anchor = direct(capacity * 4);
        out = direct(capacity * REC);
        adapterPos = new int[capacity];
    }

    private static ByteBuffer direct(int bytes) { return ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder()); }

    public void clear() { n = 0; }
    public int size() { return n; }

    /** strandedSeq = FastqRecordExt.getStrandedSeq() (L196), adapterEnd = AdapterResult.getEnd(), 1-based (L195); returns the index of the read in the batch */
    public int add(String strandedSeq, int adapterEnd) {
        if (n == capacity) throw new IllegalStateException("BarcodeBatch is full");
        int ws0 = threePrime ? adapterEnd - BC_LEN - 1 : adapterEnd;       // 0-based start of the offset-0 window (L206-L210)
        int s0 = Math.max(0, ws0 - FLANK);
        int len = Math.max(0, Math.min(SLICE, strandedSeq.length() - s0));
        int base = n * SLICE;
        for (int k = 0; k < len; k++) slices.put(base + k, (byte) strandedSeq.charAt(s0 + k));    // ASCII as it stands: case and N handling are the library's
        for (int k = len; k < SLICE; k++) slices.put(base + k, (byte) 0);
        lens.putInt(n * 4, len);
        anchor.putInt(n * 4, ws0 - s0);
        adapterPos[n] = adapterEnd;
        return n++;
    }

    /** one native call for the chunk; throws when the library reports an error (there is no CPU fallback) */
    public void run(long ctx, long table, int bcEditDistance, int testPlusMinusPos) {
        if (n == 0) return;
        int rc = Native.bcAssign(ctx, table, bcEditDistance, testPlusMinusPos, threePrime, slices, SLICE, SLICE, lens, anchor, n, out);
        if (rc != 0) throw new IllegalStateException("slr_bc_assign: " + rc + " " + Native.lastError());
    }

    /** the same through all GPUs of the box (slr_multi_bc_assign) */
    public void runMulti(long multi, long multiTable, int bcEditDistance, int testPlusMinusPos) {
        if (n == 0) return;
        int rc = Native.multiBcAssign(multi, multiTable, bcEditDistance, testPlusMinusPos, threePrime, slices, SLICE, SLICE, lens, anchor, n, out);
        if (rc != 0) throw new IllegalStateException("slr_multi_bc_assign: " + rc + " " + Native.lastError());
    }

    // slr_bc_result: u64 bc | i32 ed | i32 ed_second | i8 offset | i8 n_ins | i8 n_del | i8 n_sub | i32 rank | u32 flags | 4 bytes of padding
    public long bc(int i) { return out.getLong(i * REC); }
    public int ed(int i) { return out.getInt(i * REC + 8); }
    /** Integer.MAX_VALUE = no second-best barcode (L288-L289) */
    public int edSecond(int i) { return out.getInt(i * REC + 12); }
    public int offset(int i) { return out.get(i * REC + 16); }
    public int insertions(int i) { return out.get(i * REC + 17); }
    public int deletions(int i) { return out.get(i * REC + 18); }
    public int substitutions(int i) { return out.get(i * REC + 19); }
    public int rank(int i) { return out.getInt(i * REC + 20); }
    public int flags(int i) { return out.getInt(i * REC + 24); }
    /** BC_FOUND (L251-L253) */
    public boolean assigned(int i) { return (flags(i) & 1) != 0; }
    /** the Java would have thrown for this read (slice too short / non-IUPAC character): the shadow Parser rethrows */
    public boolean threw(int i) { return (flags(i) & 2) != 0; }
    /** BarcodeResult.start (L273-L276) */
    public int start(int i) { return threePrime ? adapterPos[i] - 1 + offset(i) : adapterPos[i] + 1 + offset(i); }
    /** BarcodeResult.end (L277-L280): OneMatch.getOffsetForReadEnd() = insertions - deletions */
    public int end(int i) {
        int d = insertions(i) - deletions(i);
        return threePrime ? start(i) - (BC_LEN - 1) - d : start(i) + (BC_LEN - 1) + d;
    }
}
