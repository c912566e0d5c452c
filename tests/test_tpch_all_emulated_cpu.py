"""All 22 TPC-H queries with the code the library GENERATES for their projections, filters and aggregates executed on the
host (tests/jit_emu.py: the generated source compiled unchanged with g++), joins / sorts / dedup by the oracle, against the
all-oracle run of the same query.  No GPU: this is the CPU-side evidence that every plan the 22 queries submit binds,
type-checks, generates code and computes what the oracle computes; tests/test_gpu_zzz_tpch_all.py adds the launches."""
import pyarrow as pa
import pytest

from daft_b200 import tpch
from daft_b200.recordbatch import RecordBatch
from helpers import assert_tables_equal
from oracle.recordbatch import OracleRecordBatch
from test_jit_host_emulation import _EmulatedBatch


class EmulatedQueryBatch(OracleRecordBatch):
    """OracleRecordBatch whose eval_expression_list / filter / agg run the generated code; the composite aggregates
    (count_distinct, string min / max) go through the product's own host-side composition (daft_b200/recordbatch.py)."""

    def _new(self, tbl):
        return EmulatedQueryBatch(tbl)

    def num_columns(self):
        return self.tbl.num_columns

    def dtypes(self):
        from daft_b200 import datatypes as dt
        return [dt.from_arrow(f.type) for f in self.tbl.schema]

    def union(self, other):
        t = self.tbl
        for n in other.tbl.column_names:
            t = t.append_column(n, other.tbl.column(n))
        return self._new(t)

    def eval_expression_list(self, exprs):
        return self._new(_EmulatedBatch(self.tbl).eval_expression_list(list(exprs)).to_arrow_table())

    def filter(self, exprs):
        return self._new(_EmulatedBatch(self.tbl).filter(list(exprs)).to_arrow_table()) if exprs else self

    _is_string_minmax = RecordBatch._is_string_minmax

    def agg(self, to_agg, group_by=None, predicate=None):
        from daft_b200.recordbatch import _inner
        if any(_inner(a).op == "count_distinct" for a in to_agg):
            return RecordBatch._agg_with_count_distinct(self, to_agg, group_by or [], predicate)
        if any(self._is_string_minmax(a) for a in to_agg):
            return RecordBatch._agg_with_string_minmax(self, to_agg, group_by or [], predicate)
        if any(_inner(a).op == "percentile" for a in to_agg):
            return RecordBatch._agg_with_percentile(self, to_agg, group_by or [], predicate)
        from daft_b200.expressions import col as _col, stage_moment_aggs
        staged = stage_moment_aggs(to_agg)   # stddev / var / skew: first-stage sums, then the final projection (as RecordBatch.agg)
        if staged is not None:
            first, final = staged
            return self.agg(first, group_by, predicate).eval_expression_list([_col(g.name()) for g in (group_by or [])] + final)
        return self._new(_EmulatedBatch(self.tbl).agg(list(to_agg), list(group_by or []), predicate).to_arrow_table())

    _resolve_keys = RecordBatch._resolve_keys

    def window(self, partition_by, order_by=(), descending=None, nulls_first=None, functions=()):
        from daft_b200.window import window
        return window(self, partition_by, order_by, descending, nulls_first, functions)

    def rename(self, names):
        return self._new(self.tbl.rename_columns(list(names)))

    def select_columns(self, idx):
        return self._new(self.tbl.select(list(idx)))

    # the two native window calls (daft_b200/csrc/sort.cu op_window_ranks / op_window_offset), restated in numpy for the host run
    def _window_ranks(self, partition_cols, order_cols):
        import numpy as np
        from oracle import oracle as O
        n = self.tbl.num_rows
        rows = lambda cols: list(zip(*[O.normalize(self.tbl.column(c)).to_pylist() for c in cols])) if cols else [()] * n
        from oracle.window import _same
        P, Q = rows(list(partition_cols)), rows(list(order_cols))
        rn, rk, dr, pb, pid = [], [], [], [], []
        for i in range(n):
            newp = i == 0 or not all(_same(a, b) for a, b in zip(P[i - 1], P[i]))
            newq = newp or not all(_same(a, b) for a, b in zip(Q[i - 1], Q[i]))
            pb.append(i if newp else pb[-1])
            pid.append((pid[-1] + 1 if pid else 0) if newp else pid[-1])
            rn.append(i - pb[-1] + 1)
            rk.append(rn[-1] if newq else rk[-1])
            dr.append(1 if newp else (dr[-1] + 1 if newq else dr[-1]))
        return self._new(pa.table({"__row_number": pa.array(rn, pa.uint64()), "__rank": pa.array(rk, pa.uint64()), "__dense_rank": pa.array(dr, pa.uint64()),
                                   "__partition_begin": pa.array(pb, pa.int64()), "__partition_id": pa.array(pid, pa.int64())}))

    def _window_offset(self, partition_ids, offset):
        import numpy as np
        from oracle import oracle as O
        pid = partition_ids.tbl.column(0).to_pylist()
        n = len(pid)
        src = [j if 0 <= (j := i + offset) < n and pid[j] == pid[i] else -1 for i in range(n)]
        t = O._take_nullable(self.tbl, np.array(src, dtype=np.int64))
        return self._new(t.append_column("__src", pa.array(src, pa.int64())))

    def _window_frame(self, partition_begin, agg, start, end, min_periods, out_dtype, kind=1, order=None, descending=False):
        """b200_window_frame through the host build of the very state machine the device kernel runs (tests/wf_host.cpp)"""
        import numpy as np
        import wf_host
        vals = self.tbl.column(0).combine_chunks() if self.tbl.num_columns else None
        pb = np.asarray(partition_begin.tbl.column(0).combine_chunks().to_numpy(zero_copy_only=False), dtype=np.int64)
        name = self.tbl.column_names[0] if self.tbl.num_columns else "count"
        ordv = order.tbl.column(0).combine_chunks() if order is not None else None
        return self._new(pa.table({name: wf_host.run(vals, pb, agg, start, end, min_periods, kind, ordv, descending)}))

    # the oracle's verbs return OracleRecordBatch: re-wrap
    def sort(self, *a, **k): return self._new(super().sort(*a, **k).tbl)
    def top_n(self, *a, **k): return self._new(super().top_n(*a, **k).tbl)
    def dedup(self, *a, **k): return self._new(super().dedup(*a, **k).tbl)
    def head(self, *a, **k): return self._new(super().head(*a, **k).tbl)
    def hash_join(self, right, *a, **k): return self._new(super().hash_join(right, *a, **k).tbl)


SF = {2: 0.05, 8: 0.05, 16: 0.05, 17: 0.05, 18: 0.05, 20: 0.05}


@pytest.mark.parametrize("q", list(range(1, 23)))
def test_query_plans_through_generated_code(q):
    name = f"q{q}"
    sf = SF.get(q, 0.01)
    if q == 1:
        cols = {"lineitem": tpch.Q1_COLUMNS}
    elif q == 6:
        cols = {"lineitem": tpch.Q1_COLUMNS}
    elif q == 5:
        cols = tpch.Q5_COLUMNS
    else:
        cols = tpch.QUERY_COLUMNS[name]
    host = {t: tpch.generate_host(t, sf, c) for t, c in cols.items()}
    fn = getattr(tpch, name)
    arg = (lambda cls: cls(host["lineitem"])) if q in (1, 6) else (lambda cls: {t: cls(v) for t, v in host.items()})
    got = fn(arg(EmulatedQueryBatch)).to_arrow_table()
    want = fn(arg(OracleRecordBatch)).to_arrow_table()
    assert_tables_equal(got, want, float_rtol=1e-9)
