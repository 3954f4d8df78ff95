"""TEST INFRASTRUCTURE — second, independent restatement (pure Python) of the read grouper: `ReadGrouper.groupSams`
(F!com/rw/umifinder/bamreaders/ReadGrouper.class, ReadGrouper.java:L82-L260; `$Cluster` L455-L667, `$ClusterList` L675-L785,
`$NanoporeReadWithOrderedPosition` L429-L447), called by `BamReader.run` on every chunk of SAM records (BamReader.java:L134-L145).
The product's implementation is the C++ behind `slr_grouper_*` (csrc/slr_group.cpp, mirror `sicelore_b200.grouping`); this module is what
`oracle/make_ref_grouper.py` checks the interpreter-run vectors against while it generates them, and what `tests/test_grouper.py` mutates to
show that the vectors pin every quirk.  Only tests/ and oracle/ import it.

It restates the class literally, including the behaviour a cleaner implementation would not have:

* a run of <= 2 reads is not closed at a gap: the reads after the gap keep joining it (ReadGrouper.java:L247);
* the read that opens a gap is added to no run (L243-L250);
* `removeOffCenterLeft` leaves the cluster's centre stale (computed before the removal) for `removeOffCenterRight` (L627-L636, L648-L659);
* centres are `Math.round((float) average)`: above 2^24 the float cast quantises the genome position (L615-L616);
* region numbers are consumed by every `Cluster` ever constructed, kept or not (L522, L534);
* `keepDataEnd` reads the `center` FIELD of the last cluster (L173): when only one cluster survives and its centre cache is empty the
  reference throws a NullPointerException — mirrored as `NullCenterError`;
* `indexInList` counts only the reads that HAVE a position but is used as an index into the whole chunk (L122, L179, L196).

Pinned against the reference's own bytecode by `oracle/make_ref_grouper.py` -> `tests/golden/ref_grouper.npz` (`tests/test_grouper.py`)."""
import math

import numpy as np


class NullCenterError(RuntimeError):
    """java.lang.NullPointerException at ReadGrouper.java:L173 (`clusters.list.get(size - 1).center.intValue()` on an empty cache)"""


def java_round_f32(x):
    """Math.round((float) x): nearest int, ties toward +infinity, on the value after the double -> float cast (ReadGrouper.java:L615-L616)"""
    f = float(np.float32(x))
    return max(-(1 << 31), min((1 << 31) - 1, math.floor(f + 0.5)))


class _Cluster:
    """ReadGrouper$Cluster: member list (indices into the position-sorted array) + lazily cached centre / min / max BAM index"""
    __slots__ = ("g", "list", "id", "center", "max_idx", "min_idx")

    def __init__(self, g, members=None):
        self.g = g
        self.center = self.max_idx = self.min_idx = None
        # Cluster(List): stream().sorted(by position) — stable (ReadGrouper.java:L530-L531)
        self.list = [] if members is None else sorted(members, key=lambda m: g._pos[m])
        self.id = g.next_region_id                             # CURRENT_GENOMIC_REGION_ID++ (L522, L534)
        g.next_region_id += 1

    def add(self, m):                                          # L557-L562
        self.center = self.max_idx = self.min_idx = None
        self.list.append(m)

    def add_all(self, ms):                                     # L570-L573
        self.list.extend(ms)
        self.center = self.max_idx = self.min_idx = None

    def remove_all(self, ms):                                  # L581-L584
        drop = set(ms)
        self.list = [m for m in self.list if m not in drop]
        self.center = self.max_idx = self.min_idx = None

    def get_center(self):                                      # getCenter -> setCenter (L491, L614-L618): recomputed only when the cache is empty
        if self.center is None and self.list:
            p = self.g._pos
            self.center = java_round_f32(sum(int(p[m]) for m in self.list) / len(self.list))
        return self.center

    def get_max_index(self):                                   # L592-L595
        if self.max_idx is None and self.list:
            self.max_idx = max(int(self.g._idx[m]) for m in self.list)
        return self.max_idx

    def _remove_off_center(self, pred):                        # the BiFunction of L626-L638
        hits = [pred(m) for m in self.list]                    # count(): the predicate runs on every member (and fills the centre cache)
        if not any(hits):
            return None
        self.max_idx = self.min_idx = self.center = None
        out = [m for m in self.list if pred(m)]                # L634: the centre is recomputed on the still complete list
        drop = set(out)
        self.list = [m for m in self.list if m not in drop]    # L635: ... and NOT cleared after the removal
        return _Cluster(self.g, out)

    def remove_off_center_left(self):                          # L648-L649
        g = self.g
        return self._remove_off_center(lambda m: int(g._pos[m]) < self.get_center() - g.max_dist)

    def remove_off_center_right(self):                         # L658-L659
        g = self.g
        return self._remove_off_center(lambda m: int(g._pos[m]) > self.get_center() + g.max_dist)


def _sorted_nonempty(clusters):
    """ClusterList.sortAndRemoveEmpty (L703): non-empty clusters, stable sort by getCenter().  A single cluster is never compared, so its
    centre cache is NOT filled (what makes the NullPointerException of L173 reachable)."""
    live = [c for c in clusters if c.list]
    return sorted(live, key=lambda c: c.get_center()) if len(live) > 1 else live


class ReadGrouper:
    """One instance = the process-wide state of the reference class: `MAX_GENOME_DISTANCE_FOR_SAME_GENOMIC_REGION`
    (config.xml:247 max_GenomeDistance_forGrouping, default 500) and the static region counter."""

    def __init__(self, max_genome_distance=500, first_region_id=0):
        self.max_dist = int(max_genome_distance)
        self.next_region_id = int(first_region_id)
        self._pos = self._idx = None

    # ---------------------------------------------------------------------------------------------------- ClusterList.refineClusters (L711-L785)
    def _refine(self, clusters):
        outliers = []
        current = clusters
        while current:                                         # L729-L731: passes of off-centre removal, each on the clusters the last one split off
            nxt = []
            for c in current:
                for o in (c.remove_off_center_left(), c.remove_off_center_right()):
                    if o is not None:
                        nxt.append(o)
            current = nxt
            outliers.extend(nxt)
        clusters = _sorted_nonempty(list(clusters) + outliers)  # L734, L752
        keep_merging = True
        while keep_merging:                                    # L756-L781
            keep_merging = False
            for i in range(len(clusters) - 1):
                left, right = clusters[i], clusters[i + 1]
                if not left.list:
                    continue
                if right.get_center() - left.get_center() < 2 * self.max_dist:
                    left_bigger = len(left.list) > len(right.list)
                    src, dst = (right, left) if left_bigger else (left, right)
                    move = [m for m in src.list if abs(int(self._pos[m]) - dst.center) <= self.max_dist]
                    if move:
                        keep_merging = True
                        dst.add_all(move)
                        src.remove_all(move)
            clusters = [c for c in clusters if c.list]
        return [c for c in clusters if len(c.list) > 1]        # L783

    # ---------------------------------------------------------------------------------------------------- doClusteringOneStrand (L234-L260)
    def _cluster_one_strand(self, indices):
        if len(indices) <= 1:
            return []
        pos, md = self._pos, self.max_dist
        clusters = []
        cur = _Cluster(self)
        if int(pos[indices[1]]) - int(pos[indices[0]]) < md:
            cur.add(indices[0])
        for i in range(1, len(indices)):
            if int(pos[indices[i]]) - int(pos[indices[i - 1]]) < md:
                cur.add(indices[i])
            elif len(cur.list) > 2:
                clusters.append(cur)
                cur = _Cluster(self)
        if len(cur.list) > 2:
            clusters.append(cur)
        return self._refine(clusters)

    # ---------------------------------------------------------------------------------------------------- groupSams (L82-L230)
    def group_sams(self, position, flags, region, keep_data_end, has_position=None):
        """One chunk of SAM records in BAM order.  position[i] = positionOnGenomeForClustering (ignored where has_position[i] is false),
        flags[i] = SAM flags (bit 16 = reverse strand), region[i] (int64, in / out) = genomicRegionNmber, -1 = absent: the reads of every
        surviving cluster get its number, the others keep what they had (a carried-over read keeps the number of the previous round).
        Returns last_index: reads [0, last_index] are the grouped chunk handed to the clustering stage, reads (last_index, n) are carried
        into the next chunk when keep_data_end (else dropped from it: the reference returns an empty chunk).  An empty chunk returns None."""
        position = np.asarray(position, dtype=np.int64)
        flags = np.asarray(flags, dtype=np.int64)
        n = len(position)
        if n == 0:
            return None
        has = np.ones(n, dtype=bool) if has_position is None else np.asarray(has_position, dtype=bool)
        filt = np.nonzero(has)[0]                              # L119-L123: indexInList counts the reads WITH a position
        order = np.argsort(position[filt], kind="stable")      # Arrays.parallelSort(Comparable[]) is stable (L128)
        self._pos = position[filt][order]
        self._idx = order.astype(np.int64)
        chunk_index = filt[order]
        rev = (flags[chunk_index] & 16) != 0
        fwd_idx = [int(i) for i in np.nonzero(~rev)[0]]
        rev_idx = [int(i) for i in np.nonzero(rev)[0]]
        clusters = self._cluster_one_strand(fwd_idx)
        clusters = _sorted_nonempty(clusters + self._cluster_one_strand(rev_idx))      # L135-L167
        last_index = n - 1
        if keep_data_end and clusters:                         # L171-L186
            most_right = int(self._pos[-1])
            while clusters:
                if clusters[-1].center is None:
                    raise NullCenterError("ReadGrouper.java:L173: centre cache of the only surviving cluster is empty")
                if clusters[-1].center <= most_right - 3 * self.max_dist:
                    break
                clusters.pop()
            if clusters:
                last_index = clusters[-1].get_max_index()
                if last_index < n // 3:
                    last_index = n // 3
        for c in clusters:                                     # L189-L191
            for m in c.list:
                region[chunk_index[m]] = c.id
        self._pos = self._idx = None
        return int(last_index)
