"""GATK-style report of the covariate tables (the intent of the reference's unused
include/kbbq/gatkreport.hh): every observed row is written once, with the table's own numbers."""
import os
import tempfile

import numpy as np
import pytest

from kbbq_b200 import api


def _parse(path):
    tables, cur, head = {}, None, None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("#:GATKTable:") and line.count(":") == 3 and line.endswith(":"):
            cur, head = line.split(":")[2], None
            tables[cur] = []
        elif line.startswith("#") or not line.strip():
            continue
        elif head is None:
            head = line.split()
        else:
            tables[cur].append(dict(zip(head, line.split())))
    return tables


def _tables(R=3, C=12, seed=5):
    rng = np.random.default_rng(seed)
    cyc = np.zeros((R, 94, 2, C, 2), np.uint64)
    for q in (2, 11, 25, 37, 40):
        cyc[:, q, :, :, 1] = rng.integers(1, 100, (R, 2, C))
        cyc[:, q, :, :, 0] = rng.integers(0, 2, (R, 2, C))
    cyc[1] = 0  # a read group nobody used: no rows
    qc = cyc.sum(axis=(2, 3))
    rgc = qc.sum(axis=1)
    dic = np.zeros((R, 94, 16, 2), np.uint64)
    dic[:, 37, :, 1] = rng.integers(0, 9, (R, 16))
    dic[:, 37, :, 0] = dic[:, 37, :, 1] // 3
    dic[1] = 0
    return rgc, qc, cyc, dic


def _check(path, rgc, qc, cyc, dic, names):
    t = _parse(path)
    assert open(path).readline().strip() == "#:GATKReport.v1.1:3"
    rows0 = {r["ReadGroup"]: r for r in t["RecalTable0"]}
    assert set(rows0) == {names[r] for r in range(rgc.shape[0]) if rgc[r, 1]}
    for r in range(rgc.shape[0]):
        if rgc[r, 1]:
            assert int(rows0[names[r]]["Observations"]) == int(rgc[r, 1]) and float(rows0[names[r]]["Errors"]) == float(rgc[r, 0])
            emp = -10 * np.log10((float(rgc[r, 0]) + 1) / (float(rgc[r, 1]) + 2))
            assert abs(float(rows0[names[r]]["EmpiricalQuality"]) - emp) < 1e-4
    assert len(t["RecalTable1"]) == int(np.count_nonzero(qc[..., 1]))
    for row in t["RecalTable1"]:
        r, q = names.index(row["ReadGroup"]), int(row["QualityScore"])
        assert int(row["Observations"]) == int(qc[r, q, 1]) and float(row["Errors"]) == float(qc[r, q, 0])
    assert len(t["RecalTable2"]) == int(np.count_nonzero(cyc[..., 1])) + int(np.count_nonzero(dic[..., 1]))
    for row in t["RecalTable2"]:
        r, q = names.index(row["ReadGroup"]), int(row["QualityScore"])
        if row["CovariateName"] == "Cycle":
            v = int(row["CovariateValue"])
            c = cyc[r, q, 1 if v < 0 else 0, abs(v) - 1]
        else:
            c = dic[r, q, "ACGT".index(row["CovariateValue"][0]) * 4 + "ACGT".index(row["CovariateValue"][1])]
        assert int(row["Observations"]) == int(c[1]) and float(row["Errors"]) == float(c[0])


def test_report_of_dense_tables_round_trips():
    rgc, qc, cyc, dic = _tables()
    names = ["lane1", "lane2", "lane3"]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "recal.table")
        api.report_tables(path, rgc, qc, cyc, dic, names)
        _check(path, rgc, qc, cyc, dic, names)
        api.report_tables(path, rgc, qc, cyc, dic)  # default names
        _check(path, rgc, qc, cyc, dic, ["rg0", "rg1", "rg2"])
        with pytest.raises(api.KbbqError):
            api.report_tables(os.path.join(d, "no/such/dir/x"), rgc, qc, cyc, dic)


@pytest.mark.gpu
def test_report_from_a_context_is_the_report_of_its_tables():
    import common
    from kbbq_b200 import synth

    G, cov, L, k = 40000, 20, 100, 32
    rs = synth.make_reads(G, cov, L, seed=3, n_rg=2)
    ctx = api.Context(k, 2, L)
    pb = api.PackedBatch(rs)
    db = ctx.upload(pb)
    flags = (rs.seq != rs.clean_seq).astype(np.uint8)
    ctx.consume_errors(db, pb.pack_bits(flags))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "recal.table")
        ctx.covariates_report(path, ["a", "b"])
        _check(path, *ctx.covariates(), ["a", "b"])
    db.free()
    ctx.close()
