"""round and utf8 match / like on the device against the oracle (same cases as the host emulation of the generated code,
tests/test_jit_host_emulation.py::test_generated_round_and_match_patterns), as projections and as filter predicates."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal
from oracle import expr_eval as E

pytestmark = pytest.mark.gpu


def test_round_and_match_like(gpu):
    c = gpu.col
    rng = np.random.default_rng(21)
    n = 5000
    x = np.concatenate([rng.standard_normal(n - 8) * 1000, [0.5, -0.5, 2.5, 0.125, -0.125, 1e300, np.nan, 0.49999999999999994]])
    words = ["special requests", "the special packages requests", "requests special", "", "Customer Complaints", "xCustomeryComplaintsz", "special",
             "specialrequests", "PROMO BRUSHED", "MEDIUM POLISHED TIN", "abc", "ab", "abcabc"]
    with np.errstate(all="ignore"):
        tbl = pa.table({"f64": pa.array(x, mask=rng.random(n) < 0.1), "f32": pa.array(x.astype(np.float32), mask=rng.random(n) < 0.1),
                        "s": pa.array([words[i] for i in rng.integers(0, len(words), n)], type=pa.large_utf8(), mask=rng.random(n) < 0.1)})
    exprs = [c("f64").round(2).alias("r2"), c("f64").round(0).alias("r0"), c("f32").round(1).alias("r1f"), c("f64").round(-2).alias("rm2"),
             c("s").match(".*special.*requests.*").alias("m1"), c("s").match("special").alias("m2"), c("s").match("^special.*requests$").alias("m3"),
             c("s").like("%Customer%Complaints%").alias("l1"), c("s").like("PROMO%").alias("l2"), c("s").like("%TIN").alias("l3"),
             c("s").like("abc").alias("l4"), c("s").like("ab%c").alias("l5"), c("s").like("%").alias("l6"), (~c("s").like("MEDIUM POLISHED%")).alias("l7")]
    rb = gpu.RecordBatch.from_arrow_table(tbl)
    assert_tables_equal(rb.eval_expression_list(exprs).to_arrow_table(), E.eval_expression_list(tbl, exprs))
    for pred in ([~c("s").match(".*special.*requests.*")], [c("s").like("%Customer%Complaints%"), c("f64").round(0) > 0.0]):
        assert_tables_equal(rb.filter(pred).to_arrow_table(), E.filter(tbl, pred))
