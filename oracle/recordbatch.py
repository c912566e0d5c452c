"""Oracle twin of daft_b200.RecordBatch over host pyarrow Tables (TEST INFRASTRUCTURE ONLY).

Same method surface, so the query pipelines in daft_b200/tpch.py run unchanged on the CPU oracle and on the
CUDA path and can be compared table for table."""
from __future__ import annotations

from typing import Sequence

import pyarrow as pa

from daft_b200.expressions import Expression

from . import expr_eval as E
from . import oracle as O


class OracleRecordBatch:
    def __init__(self, tbl: pa.Table):
        self.tbl = tbl

    @staticmethod
    def from_arrow_table(tbl: pa.Table) -> "OracleRecordBatch":
        return OracleRecordBatch(tbl)

    def to_arrow_table(self) -> pa.Table:
        return self.tbl

    def __len__(self):
        return self.tbl.num_rows

    def column_names(self):
        return self.tbl.column_names

    def num_columns(self):
        return self.tbl.num_columns

    def dtypes(self):
        from daft_b200 import datatypes as dt
        return [dt.from_arrow(f.type) for f in self.tbl.schema]

    def slice(self, start: int, end: int) -> "OracleRecordBatch":
        return OracleRecordBatch(self.tbl.slice(start, max(end - start, 0)))

    def eval_expression_list(self, exprs: Sequence[Expression]) -> "OracleRecordBatch":
        return OracleRecordBatch(E.eval_expression_list(self.tbl, exprs))

    def filter(self, exprs: Sequence[Expression]) -> "OracleRecordBatch":
        return OracleRecordBatch(E.filter(self.tbl, exprs) if exprs else self.tbl)

    def agg(self, to_agg, group_by=None, predicate=None) -> "OracleRecordBatch":
        from daft_b200.expressions import col as _col, stage_moment_aggs
        staged = stage_moment_aggs(to_agg)  # stddev / var: the reference's staged rewrite (src/daft-local-plan/src/agg.rs:190-258)
        if staged is not None:
            first, final = staged
            keys = [_col(g.name()) for g in (group_by or [])]
            return self.agg(first, group_by, predicate).eval_expression_list(keys + final)
        tbl = E.filter(self.tbl, predicate) if predicate else self.tbl
        specs = []
        for a in to_agg:
            inner = a
            while inner.kind == "alias":
                inner = inner.args[0]
            specs.append((inner.op, E.eval_expression(tbl, inner.args[0]), a.name(), inner.value))
        keys = [(g.name(), E.eval_expression(tbl, g)) for g in (group_by or [])]
        return OracleRecordBatch(O.agg(tbl, specs, keys))

    def dedup(self, exprs) -> "OracleRecordBatch":
        """first row of every distinct key, in first-occurrence order (src/daft-recordbatch/src/ops/groups.rs make_groups)"""
        _, first, _ = O.group_ids([E.eval_expression(self.tbl, e) for e in exprs])
        import numpy as np
        return OracleRecordBatch(self.tbl.take(pa.array(np.sort(first.astype(np.int64)))))

    def head(self, num: int) -> "OracleRecordBatch":
        return OracleRecordBatch(self.tbl.slice(0, num))

    def sort(self, sort_keys, descending=None, nulls_first=None) -> "OracleRecordBatch":
        n = len(sort_keys)
        desc = [descending] * n if isinstance(descending, bool) else (descending or [False] * n)
        nf = [nulls_first] * n if isinstance(nulls_first, bool) else (nulls_first if nulls_first is not None else list(desc))
        idx = O.argsort([E.eval_expression(self.tbl, k) for k in sort_keys], desc, nf)
        return OracleRecordBatch(self.tbl.take(pa.array(idx, type=pa.int64())))

    def top_n(self, sort_keys, limit: int, descending=None, nulls_first=None, offset: int = 0) -> "OracleRecordBatch":
        """RecordBatch::top_n = argsort, then slice (src/daft-recordbatch/src/ops/sort.rs:47-61)."""
        s = self.sort(sort_keys, descending, nulls_first).tbl
        return OracleRecordBatch(s.slice(offset, limit))

    def hash_join(self, right: "OracleRecordBatch", left_on, right_on, how="inner", null_equals_null=None,
                  build_on_left=None) -> "OracleRecordBatch":
        lt, rt = self.tbl, right.tbl
        lk, rk = [], []
        for i, (le, re_) in enumerate(zip(left_on, right_on)):
            assert le.kind == "col" and re_.kind == "col", "oracle join supports plain key columns"
            lk.append(le.value)
            rk.append(re_.value)
        # clashing non-key right names get the `right.` prefix (as the product does)
        common = [l for l, r in zip(lk, rk) if l == r]
        new_names = [n if (n in common or n not in lt.column_names) else f"right.{n}" for n in rt.column_names]
        rmap = dict(zip(rt.column_names, new_names))
        rt2 = rt.rename_columns(new_names)
        return OracleRecordBatch(O.hash_join(lt, rt2, lk, [rmap[k] for k in rk], how, null_equals_null, build_on_left))
