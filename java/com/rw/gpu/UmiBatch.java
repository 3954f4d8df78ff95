package com.rw.gpu;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * The (cell, gene-region) jobs of one BAM chunk for the UMI side: packing of the UMI strings into the CSR buffers of slr_umi_dist / slr_umi_cluster /
 * slr_umi_assign and typed access to the per-read records.  Jar-independent (compiles with Native.java alone); the shadow UmiClustering fills it
 * from OneCellOneGeneRegionData.getData() (one addJob per Callable that UmiClustering.java:L45-L60 would create) and finishes the
 * NanoporeResultClusteringInfo objects from the records.
 *
 *   batch.clear();
 *   for every job:  batch.beginJob(qv(read0) > qv(read1));  for every read:  batch.addRead(s.getNaData(), bcEnd - 1);  batch.endJob();
 *   batch.assign(ctx, null);                              // ClusterOneHierarchical.call for every job of <= 100 reads
 *   for every read r of job j:  batch.center(r), batch.u1(r), batch.u2(r), ...
 */
public final class UmiBatch {
    public static final int STRIDE = 16, REC = 16;
    public static final int ASSIGNED = 1, SKIPPED = 2, TIE_UNPIN = 4, DEEP = 8;
    private ByteBuffer umis, jobOffsets, qv01, rec;
    private int nReads, nJobs, umiLen;

    public UmiBatch(int readCapacity, int jobCapacity, int umiLen) {
        this.umiLen = umiLen;
        umis = direct(readCapacity * STRIDE);
        jobOffsets = direct((jobCapacity + 1) * 8);
        qv01 = direct(jobCapacity);
        rec = direct(readCapacity * REC);
        clear();
    }

    private static ByteBuffer direct(int bytes) { return ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder()); }

    public void clear() { nReads = 0; nJobs = 0; jobOffsets.putLong(0, 0L); }
    public int reads() { return nReads; }
    public int jobs() { return nJobs; }

    /** firstReadHasHigherQv = mean_qv(read 0) > mean_qv(read 1), the only quality value the clusterers look at (OneUmiCluster.java:L53) */
    public void beginJob(boolean firstReadHasHigherQv) {
        if ((nJobs + 2) * 8 > jobOffsets.capacity()) grow(false);
        qv01.put(nJobs, (byte) (firstReadHasHigherQv ? 1 : 0));
    }

    /** codes = NucleicAcidInmutableOneBytePerBase.getNaData() of the strand-corrected X= mini sequence (A=1 G=2 C=4 T=8 N=15 ...,
     *  NucleicAcidByteCodeBase.java:L45-L78), from = bcEnd - 1: the umiLen + 2 codes of getSubSequence(bcEnd, umiLen + 2), i.e. the predicted UMI window
     *  widened by one base on each side for the -1 / 0 / +1 shifts (ClusteringEditDistanceBase.java:L312-L329) */
    public void addRead(byte[] codes, int from) {
        if ((nReads + 1) * STRIDE > umis.capacity()) grow(true);
        int base = nReads * STRIDE, len = Math.max(0, Math.min(umiLen + 2, codes.length - from));
        for (int k = 0; k < len; k++) umis.put(base + k, codes[from + k]);
        for (int k = len; k < STRIDE; k++) umis.put(base + k, (byte) 0);
        nReads++;
    }

    /** the same from text (tests, tools): A C G T and N only */
    public void addReadAscii(CharSequence window) {
        byte[] codes = new byte[window.length()];
        for (int k = 0; k < codes.length; k++) {
            switch (Character.toUpperCase(window.charAt(k))) {
                case 'A': codes[k] = 1; break;
                case 'G': codes[k] = 2; break;
                case 'C': codes[k] = 4; break;
                case 'T': codes[k] = 8; break;
                default: codes[k] = 15;
            }
        }
        addRead(codes, 0);
    }

    public void endJob() { nJobs++; jobOffsets.putLong(nJobs * 8, (long) nReads); }

    private void grow(boolean readSide) {
        if (readSide) {
            ByteBuffer u = direct(umis.capacity() * 2), r = direct(rec.capacity() * 2);
            for (int k = 0; k < nReads * STRIDE; k++) u.put(k, umis.get(k));
            umis = u; rec = r;
        } else {
            ByteBuffer j = direct(jobOffsets.capacity() * 2), q = direct(qv01.capacity() * 2 + 2);
            for (int k = 0; k <= nJobs; k++) j.putLong(k * 8, jobOffsets.getLong(k * 8));
            for (int k = 0; k < nJobs; k++) q.put(k, qv01.get(k));
            jobOffsets = j; qv01 = q;
        }
    }

    /** params = null (UMIparameters defaults) or {completeLinkED, singleLinkED, singleLinkThreshold, foldDepthBelowMax, maxHier[, deep]} */
    public void assign(long ctx, int[] params) {
        if (nJobs == 0) return;
        int rc = Native.umiAssign(ctx, umis, STRIDE, umiLen, jobOffsets, nJobs, params, qv01, null, null, rec);
        if (rc != 0) throw new IllegalStateException("slr_umi_assign: " + rc + " " + Native.lastError());
    }

    public void assignMulti(long multi, int[] params) {
        if (nJobs == 0) return;
        int rc = Native.multiUmiAssign(multi, umis, STRIDE, umiLen, jobOffsets, nJobs, params, qv01, rec);
        if (rc != 0) throw new IllegalStateException("slr_multi_umi_assign: " + rc + " " + Native.lastError());
    }

    // slr_umi_assign_rec: i32 center | i8 u1 | i8 u2 | i8 pos2 | i8 offset_center_mean | u16 flags | u16 cluster_size | i32 n_clusters
    /** index (inside the job) of the cluster's centre read, -1 = the read is in no cluster */
    public int center(int read) { return rec.getInt(read * REC); }
    public int u1(int read) { return rec.get(read * REC + 4); }
    public int u2(int read) { return rec.get(read * REC + 5); }
    public int pos2(int read) { return rec.get(read * REC + 6); }
    public int offsetCenterMean(int read) { return rec.get(read * REC + 7); }
    public int flags(int read) { return rec.getShort(read * REC + 8) & 0xFFFF; }
    public int clusterSize(int read) { return rec.getShort(read * REC + 10) & 0xFFFF; }
    public int nClusters(int read) { return rec.getInt(read * REC + 12); }
    public boolean assigned(int read) { return (flags(read) & ASSIGNED) != 0; }
    /** job of more than maxHier reads: the record comes from the large-job path (ClusterOne_MyClustering.call, sequential-stream semantics) — or, with
     *  params[5] == 0, the job was only flagged and is left to the jar's own class / Native.umiCluster */
    public boolean deep(int read) { return (flags(read) & DEEP) != 0; }
    /** equal-score merge order depended on identity hash codes in the reference JVM: re-run this job on the CPU if bit-for-bit agreement with one
     *  particular JVM run matters (the reference itself is not reproducible across runs for such a job) */
    public boolean tieUnpinned(int read) { return (flags(read) & TIE_UNPIN) != 0; }
    public long jobStart(int job) { return jobOffsets.getLong(job * 8); }
}
