"""ReadGrouper.groupSams and the job former (the grouping that produces the clustering jobs): the library's native host code
(csrc/slr_group.cpp through `sicelore_b200.grouping`) and the Python restatement `oracle/pyref_group.py` against vectors the reference's own
bytecode produced (oracle/make_ref_grouper.py -> tests/golden/ref_grouper.npz, oracle/make_ref_jobs.py -> ref_jobs.npz), and the properties
of the chunk loop around it.  No GPU needed: this is host code on both sides (the reference runs it on its BAM reader thread)."""
import importlib
import os

import numpy as np
import pytest

import __graft_entry__ as g

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_grouper.npz")


@pytest.fixture(scope="module")
def grouping():
    g.load_package().build()
    return importlib.import_module("sicelore_b200.grouping")


@pytest.fixture(scope="module")
def pyg():
    from oracle import pyref_group
    return pyref_group


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def replay(grouping, z, make_grouper=None):
    """run every recorded groupSams call through the mirror; returns the indices of the calls whose outcome differs"""
    off = z["offsets"]
    bad = []
    for c in range(len(off) - 1):
        a, b = int(off[c]), int(off[c + 1])
        G = (make_grouper or grouping.ReadGrouper)(int(z["max_dist"][c]), int(z["id_before"][c]))
        reg = z["region_in"][a:b].copy()
        try:
            li = G.group_sams(z["position"][a:b], z["flags"][a:b], reg, bool(z["keep_data_end"][c]), z["has_position"][a:b].astype(bool))
            thrown = ""
        except grouping.NullCenterError:
            li, thrown = None, "java/lang/NullPointerException"
        ok = thrown == str(z["thrown"][c]) and G.next_region_id == int(z["id_after"][c])
        if ok and not thrown:
            n_done = 0 if li is None else li + 1
            ok = n_done == int(z["n_done"][c]) and np.array_equal(reg, z["region_out"][a:b])
            if ok and z["keep_data_end"][c]:
                ok = (b - a) - n_done == int(z["n_carried"][c])
        if not ok:
            bad.append(c)
    return bad


def test_native_grouper_matches_reference_bytecode(grouping, pyg, gold):
    z = gold
    assert len(z["offsets"]) - 1 >= 400 and int(z["offsets"][-1]) >= 15000
    assert replay(grouping, z) == []                                        # slr_grouper_group_sams
    assert replay(pyg, z) == []                                             # the independent Python restatement
    # what the vectors contain: an exception case, carried-over reads with a region number from the previous round, reads without a position
    assert (z["thrown"] != "").sum() >= 1
    assert (z["region_in"] >= 0).sum() > 100 and (z["has_position"] == 0).sum() > 10 and (z["n_carried"] > 0).sum() > 50
    assert (z["position"] > (1 << 24)).sum() > 1000


def test_vectors_pin_every_quirk(pyg, gold):
    """each behaviour a cleaner implementation would not have (oracle/pyref_group.py docstring) changes at least one recorded outcome"""
    grouping = pyg
    G0, C0 = grouping.ReadGrouper, grouping._Cluster

    def variant(**patch):
        class C(C0):
            __slots__ = ()
        class G(G0):
            pass
        for k, v in patch.items():
            setattr(C if hasattr(C0, k) else G, k, v)
        return C, G

    def run(C, G, module_patch=None):
        old = {k: getattr(grouping, k) for k in ("_Cluster", "java_round_f32", "_sorted_nonempty")}
        try:
            grouping._Cluster = C
            for k, v in (module_patch or {}).items():
                setattr(grouping, k, v)
            return replay(grouping, gold, G)
        finally:
            for k, v in old.items():
                setattr(grouping, k, v)

    # (1) centre cleared after an off-centre removal (the reference keeps the stale one for the right-hand test)
    def fresh_remove(self, pred):
        out = C0._remove_off_center(self, pred)
        if out is not None:
            self.center = None
        return out
    C, G = variant(_remove_off_center=fresh_remove)
    assert run(C, G), "stale centre"
    # (2) centre rounded in double precision instead of through the float cast
    C, G = variant()
    assert run(C, G, {"java_round_f32": lambda x: int(np.floor(x + 0.5))}), "float cast of the centre"
    # (3) sortAndRemoveEmpty filling the centre of a single cluster: the NullPointerException disappears
    C, G = variant()
    assert run(C, G, {"_sorted_nonempty": lambda cl: sorted([c for c in cl if c.list], key=lambda c: c.get_center())}), "single-cluster sort"
    # (4) a short run closed at a gap
    def strict_strand(self, indices):
        if len(indices) <= 1:
            return []
        pos, md, clusters, cur = self._pos, self.max_dist, [], grouping._Cluster(self)
        if int(pos[indices[1]]) - int(pos[indices[0]]) < md:
            cur.add(indices[0])
        for i in range(1, len(indices)):
            if int(pos[indices[i]]) - int(pos[indices[i - 1]]) < md:
                cur.add(indices[i])
            else:
                if len(cur.list) > 2:
                    clusters.append(cur)
                cur = grouping._Cluster(self)
        if len(cur.list) > 2:
            clusters.append(cur)
        return self._refine(clusters)
    C, G = variant(_cluster_one_strand=strict_strand)
    assert run(C, G), "short runs survive a gap"
    # (5) region numbers handed out only to surviving clusters
    assert len(set(int(x) for x in gold["id_after"] - gold["id_before"])) > 3


def test_region_numbers_follow_the_static_counter(grouping):
    G = grouping.ReadGrouper(500)
    pos = np.array([100, 120, 130, 140, 9000, 9010, 9020, 9030], dtype=np.int64)
    reg = np.full(8, -1, dtype=np.int64)
    assert G.group_sams(pos, np.zeros(8, dtype=np.int64), reg, False) == 7
    first = G.next_region_id
    assert reg.tolist() == [0, 0, 0, 0, -1, 1, 1, 1] and first == 2     # the read that opens the gap joins no run (ReadGrouper.java:L243-L250)
    reg2 = np.full(8, -1, dtype=np.int64)
    G.group_sams(pos, np.full(8, 16, dtype=np.int64), reg2, False)          # the next chunk continues the numbering
    assert reg2.tolist() == [2, 2, 2, 2, -1, 3, 3, 3]
    assert G.group_sams(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), True) is None


def test_strands_are_grouped_separately(grouping):
    G = grouping.ReadGrouper(500)
    pos = np.arange(1000, 1012, dtype=np.int64)
    flags = np.array([0, 16] * 6, dtype=np.int64)
    reg = np.full(12, -1, dtype=np.int64)
    G.group_sams(pos, flags, reg, False)
    assert len(set(reg[flags == 0].tolist())) == 1 and len(set(reg[flags == 16].tolist())) == 1 and reg[0] != reg[1]


def test_group_stream_emits_every_record_once(grouping):
    rng = np.random.default_rng(5)
    n = 5000
    chrom = np.sort(rng.integers(0, 3, n))
    pos = np.concatenate([np.sort(rng.integers(0, 400_000, int((chrom == c).sum()))) for c in range(3)])
    flags = np.where(rng.random(n) < 0.5, 16, 0)
    G = grouping.ReadGrouper(500)
    region, emitted = grouping.group_stream(G, pos, flags, chrom, 700)
    allidx = np.concatenate(emitted)
    assert np.array_equal(allidx, np.arange(n))                             # BAM order is preserved, nothing is lost or doubled
    for ids in emitted:                                                     # a chunk never spans two reference sequences
        assert len(set(chrom[ids].tolist())) == 1
    # a region never mixes strands or chromosomes, and all its reads lie within a few max_dist
    for r in np.unique(region[region >= 0]):
        m = region == r
        assert len(set((flags[m] & 16).tolist())) == 1 and len(set(chrom[m].tolist())) == 1
    assert (region >= 0).mean() > 0.5
    # chunking changes nothing for records far from a chunk end: one big chunk per chromosome gives the same partition of most reads
    G2 = grouping.ReadGrouper(500)
    region2, _ = grouping.group_stream(G2, pos, flags, chrom, 10 ** 9)
    same = 0
    for r in np.unique(region2[region2 >= 0]):
        m = region2 == r
        same += int(len(set(region[m].tolist())) == 1)
    assert same > 0.8 * len(np.unique(region2[region2 >= 0]))


def test_grouper_feeds_the_job_former(grouping):
    """ReadGrouper -> groupDataByCellAndRegion -> CSR jobs: the chain a Java-free driver runs before slr_umi_assign"""
    pkg = g.load_package()
    rng = np.random.default_rng(9)
    n = 3000
    pos = np.sort(rng.choice(np.arange(0, 200_000, 4000), n) + rng.integers(0, 60, n))
    flags = np.zeros(n, dtype=np.int64)
    cell = rng.integers(1, 40, n).astype(np.uint64)
    G = grouping.ReadGrouper(500)
    region, _ = grouping.group_stream(G, pos, flags, np.zeros(n, dtype=np.int64), 1000)
    order, off = grouping.group_jobs(cell, region, region >= 0)
    o2, off2 = pkg.group_by_cell_and_region(cell, region, region >= 0)
    assert np.array_equal(order, o2) and np.array_equal(off, off2)
    assert len(off) > 100 and off[-1] == len(order)
    for j in range(len(off) - 1):
        ids = order[off[j]:off[j + 1]]
        assert len(ids) >= 2 and len(set(cell[ids].tolist())) == 1 and len(set(region[ids].tolist())) == 1


def test_job_former_matches_reference_bytecode(grouping):
    """UmiClustering.cluster up to its Submitter (groupDataByCellAndRegion, size filter, split of oversized groups; UmiClustering.java:L97-L145)
    run from the class files by oracle/make_ref_jobs.py: the host mirrors form the same set of jobs, reads inside a job in input order"""
    pkg = g.load_package()
    z = np.load(os.path.join(os.path.dirname(GOLD), "ref_jobs.npz"))
    off, joff = z["offsets"], z["job_offsets"]
    assert len(off) - 1 >= 80 and len(joff) - 1 >= 2500
    n_split = 0
    for c in range(len(off) - 1):
        bc, region = z["barcode"][off[c]:off[c + 1]], z["region"][off[c]:off[c + 1]]
        want = sorted(z["job_reads"][joff[j]:joff[j + 1]].tolist() for j in np.nonzero(z["job_case"] == c)[0])
        order, o = pkg.group_by_cell_and_region(bc.astype(np.uint64), region, (bc >= 0) & (region >= 0))
        got = []
        for j in range(len(o) - 1):
            ids = order[o[j]:o[j + 1]].tolist()
            a = 0
            parts = pkg.split_oversized_group(len(ids), int(z["ram"][c]))
            n_split += len(parts) > 1
            for sz in parts:
                got.append(ids[a:a + sz])
                a += sz
        assert sorted(got) == want, c
        order, o = grouping.group_jobs(bc.astype(np.uint64), region, (bc >= 0) & (region >= 0), 2, int(z["ram"][c]))      # slr_group_jobs
        assert sorted(order[o[j]:o[j + 1]].tolist() for j in range(len(o) - 1)) == want, c
    assert n_split > 20
