package com.rw.gpu;

import java.nio.ByteBuffer;

/**
 * JNI binding of libsicelore_gpu.so (include/sicelore_gpu.h) for NanoporeBC_UMI_finder-2.1.jar.
 * Not compiled in the build image (no JDK there); the glue is java/sicelore_gpu_jni.c, compile-checked against a stub jni.h by
 * tests/test_abi.py.  All buffers are DIRECT ByteBuffers in native byte order; records are documented in INTEGRATION.md.
 * Every int-returning method returns 0 or a negative SLR_E_* code; lastError() has the message.  There is no CPU fallback.
 */
public final class Native {
    static { System.loadLibrary("sicelore_gpu_jni"); }
    private Native() {}

    public static native long ctxCreate(int device, int nStreams);                         // slr_ctx_create  (0 = failed)
    public static native void ctxDestroy(long ctx);                                        // slr_ctx_destroy
    /** barcodes2bit = key set of the Long2ObjectOpenHashMap (WorkerReadscanner.java:L264-L269), rank = CountsRank.rank or null */
    public static native long bcTableCreate(long ctx, long[] barcodes2bit, int[] rank);    // slr_bc_table_create (0 = failed)
    public static native void bcTableDestroy(long table);
    public static native long bcTableSize(long table);
    /** Parser.assignBarcode for a whole ReadChunk (Parser.java:L198-L252); lens may be null */
    public static native int bcAssign(long ctx, long table, int edMax, int plusMinus, boolean threePrime, ByteBuffer slices, int stride,
                                      int sliceLen, ByteBuffer lens, ByteBuffer anchor, long n, ByteBuffer out);
    /** UsedCellBCListGenerator$Worker exact lookup (UsedCellBCListGenerator.java:L206-L232) */
    public static native int bcExact(long ctx, long table, boolean threePrime, ByteBuffer slices, int stride, int sliceLen,
                                     ByteBuffer lens, ByteBuffer anchor, long n, ByteBuffer out);
    /** assignedBarcodes2ndPass / unfilteredUsedBarcodeMap counters: countsOut.length = 3 * number of barcodes */
    public static native int bcCountsRead(long ctx, long table, long[] countsOut);
    public static native int bcCountsReset(long ctx, long table);
    /** BarcodeDatasetColissionTester.submitSeq loop (BarcodeDatasetColissionTester.java:L212-L229); out: n * 24 bytes */
    public static native int bcCollide(long ctx, long table, int edMax, long[] barcodes, ByteBuffer out);
    /** ClusteringEditDistanceBase.generateDistanceMatrix for all jobs of a BAM chunk (…java:L168-L259) */
    public static native int umiDist(long ctx, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, ByteBuffer out,
                                     ByteBuffer outOffsets);
    /** the two O(n^2) steps of ClusterOne_MyClustering.clusterLocal (…java:L175-L219) behind the same matrices: rec = 16 bytes per read
     *  {|N(c)|, chosen entry or -1, |N(entry)|, entries tied for the maximum}; member / rank / out / outOffsets nullable */
    public static native int umiCluster(long ctx, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, int ed,
                                        ByteBuffer member, ByteBuffer rank, ByteBuffer out, ByteBuffer outOffsets, ByteBuffer rec);
    /** the same with the matrices kept on the device between calls: create (distance kernels run once), cluster with rank = null, fill the
     *  map with the keys, cluster again with their iteration ranks (and once more with the unclustered reads as member), fetch the matrices
     *  if the distances are needed, destroy  (0 = failed) */
    public static native long umiSessionCreate(long ctx, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs);
    public static native int umiSessionCluster(long session, int ed, ByteBuffer member, ByteBuffer rank, ByteBuffer rec);
    public static native long umiSessionCells(long session);
    public static native int umiSessionMatrices(long session, ByteBuffer out, long nCells);
    public static native void umiSessionDestroy(long session);
    /** UmiClustering$Submitter's two clusterers behind the same matrices: ClusterOneHierarchical.call for every job of at most 100 reads
     *  (ClusterOneHierarchical.java:L61-L217) and ClusterOne_MyClustering.call for the larger ones (ClusterOne_MyClustering.java:L59-L219, sequential-stream
     *  semantics): rec = 16 bytes per read {int center, byte u1, byte u2, byte pos2, byte offsetCenterMean, short flags, short clusterSize, int nClusters};
     *  params = null (config.xml / UMIparameters defaults) or int[5..6] {completeLinkED, singleLinkED, singleLinkThreshold, foldDepthBelowMax, maxHier,
     *  deep (1 = cluster the jobs above maxHier too, the default; 0 = only flag them)}; jobQv01 nullable: one byte per job,
     *  mean_qv(read 0) > mean_qv(read 1) (OneUmiCluster.java:L53) */
    public static native int umiAssign(long ctx, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, int[] params,
                                       ByteBuffer jobQv01, ByteBuffer out, ByteBuffer outOffsets, ByteBuffer rec);
    public static native int umiSessionAssign(long session, int[] params, ByteBuffer jobQv01, ByteBuffer rec);
    /** all GPUs of the box from this one JVM (slr_multi_*): nDevices <= 0 = every visible device, deviceIds nullable  (0 = failed) */
    public static native long multiCreate(int nDevices, int[] deviceIds, int nStreams);
    public static native void multiDestroy(long multi);
    public static native int multiDevices(long multi);
    public static native long multiBcTableCreate(long multi, long[] barcodes2bit, int[] rank);                     // replicated on every device (0 = failed)
    public static native void multiBcTableDestroy(long table);
    public static native int multiBcAssign(long multi, long table, int edMax, int plusMinus, boolean threePrime, ByteBuffer slices, int stride,
                                           int sliceLen, ByteBuffer lens, ByteBuffer anchor, long n, ByteBuffer out);
    public static native int multiBcExact(long multi, long table, boolean threePrime, ByteBuffer slices, int stride, int sliceLen, ByteBuffer lens,
                                          ByteBuffer anchor, long n, ByteBuffer out);
    /** assignedBarcodes2ndPass summed over the devices on the GPU (peer loads over NVLink), countsOut.length = 3 * number of barcodes */
    public static native int multiBcCountsRead(long multi, long table, long[] countsOut);
    public static native int multiBcCountsReset(long multi, long table);
    public static native int multiUmiDist(long multi, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, ByteBuffer out,
                                          ByteBuffer outOffsets);
    public static native int multiUmiCluster(long multi, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, int ed,
                                             ByteBuffer member, ByteBuffer rank, ByteBuffer rec);
    public static native int multiUmiAssign(long multi, ByteBuffer umis, int stride, int umiLen, ByteBuffer jobOffsets, long nJobs, int[] params,
                                            ByteBuffer jobQv01, ByteBuffer rec);
    /** candidate sets of the Illumina-guided search: groupKeys / groupOffsets = CSR of the per-(gene, cell) UMIs (IlluminaOneGeneOneCellData) or of the
     *  per-gene cell barcodes (BarcodesMap); allKeys = All10xselectedCells, emptyKeys = EmptyDropBarcodes (BC flavour, nullable)  (0 = failed) */
    public static native long guidedSetsCreate(long ctx, long[] groupKeys, long[] groupOffsets, long[] allKeys, int allEd, long[] emptyKeys,
                                               int emptyEd, boolean bcFlavour, int seqLen);
    public static native void guidedSetsDestroy(long sets);
    /** offset loop of IlluminaUMIanalyzer.findUMI (…java:L89-L136) / IlluminaBarcodeAnalyzer.testBarcodes (…java:L272-L304) + sorted().distinct()
     *  (IlluminaBarcodeUMIAnalyzerBase.java:L52-L60) for n reads; out: n * 40 bytes (slr_guided_result); rawOut nullable: n * rawCap * 16 bytes */
    public static native int guidedMatch(long ctx, long sets, int plusMinus, int postLen, int bailout, ByteBuffer slices, int stride, int sliceLen,
                                         ByteBuffer anchor, ByteBuffer groupId, ByteBuffer ed, long n, ByteBuffer out, ByteBuffer rawOut, int rawCap);
    /** DynamicEditDistances.getmaxED (DynamicEditDistances.java:L93-L98); cap < 0 = null; -1 = NoSuchElementException */
    /** nMismatchDiffBestvsSecondBest per record of guidedMatch (IlluminaBarcodeUMIAnalyzerBase.java:L66-L86): 0 = MORE_THAN_ONE_MATCH,
     *  Integer.MIN_VALUE = no second-best entry; scores7 = NeedlemanScores fields in declaration order or null = defaults */
    public static native int guidedMismatchDiff(ByteBuffer records, long n, ByteBuffer slices, int stride, int sliceLen, ByteBuffer anchor,
                                                int seqLen, int[] scores7, ByteBuffer diffOut);
    public static native int dynMaxEd(long[] maxCandidates, int count, int plusMinus, int cap);
    public static native String lastError();
    public static native int abiVersion();
}
