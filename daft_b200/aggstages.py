"""Two-stage aggregation rewrite and the grouped-aggregate sink on top of the C ABI.

`populate_aggregation_stages` is the host mirror of populate_aggregation_stages_bound
(src/daft-local-plan/src/agg.rs:13-416): an aggregation list becomes a first-stage list (run per morsel / per scan
partition), a second-stage list (run over the concatenated partials) and a final projection.  Equivalent aggregates
are added to a stage once (the reference's IndexSet), later references reuse the column.

`GroupedAggregateSink` is the operator protocol of src/daft-local-execution/src/sinks/grouped_aggregate.rs:146-488 over
the b200_groupby_* entry points: make_state -> sink(morsel)* -> finalize.  The partials stay in HBM; nothing is
concatenated on the host.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

from . import _lib
from . import datatypes as dt
from .expressions import Expression, _agg_inner, col, encode_agg, encode_project, lit


class _Stage:
    """An ordered set of aggregate expressions keyed by structure (IndexSet<BoundAggExpr>)."""

    def __init__(self, prefix: str):
        self.prefix, self.exprs, self.index = prefix, [], {}

    @staticmethod
    def _key(e: Expression) -> str:
        """Structural identity, including what repr() leaves out (count mode, ddof, cast target, literal dtype)."""
        return f"{e.kind}:{e.op}:{e.value!r}:{e.dtype}:{e._name}(" + ",".join(_Stage._key(a) for a in e.args) + ")"

    def add(self, agg: Expression) -> Expression:
        key = self._key(agg)
        if key not in self.index:
            name = f"{self.prefix}{len(self.exprs)}"
            self.index[key] = name
            self.exprs.append(agg.alias(name))
        return col(self.index[key])


def populate_aggregation_stages(aggregations: Sequence[Expression], group_by: Sequence[Expression]):
    """-> (first_stage_aggs, second_stage_aggs, final_exprs).  The first stage groups by `group_by`; its output columns
    are the key columns (named after the key expressions) followed by the first-stage aggregates; the second stage
    groups by those key columns; `final_exprs` (key columns first) name the outputs after the original aggregates."""
    first, second = _Stage("__s1_"), _Stage("__s2_")
    final = [col(g.name()) for g in group_by]
    for a in aggregations:
        inner = _agg_inner(a)
        if inner.kind != "agg":
            raise ValueError(f"expected an aggregation expression, got {a!r}")
        x, out = inner.args[0], a.name()
        op = inner.op
        if op == "count":            # agg.rs:113-117
            final.append(second.add(first.add(x.count(inner.value)).sum()).alias(out))
        elif op == "sum":            # :128-132
            final.append(second.add(first.add(x.sum()).sum()).alias(out))
        elif op == "mean":           # :158-166: Sum + Count(Valid), merge_mean = sum / count (functions/agg/mod.rs:29-53)
            s = second.add(first.add(x.sum()).sum())
            c = second.add(first.add(x.count("valid")).sum())
            final.append((s / c).alias(out))
        elif op in ("min", "max"):   # :259-268 (bool_and / bool_or are min / max over Boolean: :269-278)
            g = first.add(getattr(x, op)())
            final.append(second.add(getattr(g, op)()).alias(out))
        elif op == "any_value":      # :279-283
            ign = inner.value == "ignore_nulls"
            g = first.add(x.any_value(ign))
            final.append(second.add(g.any_value(ign)).alias(out))
        elif op == "list":           # :286-290: List per partial, Concat of the partial lists
            final.append(second.add(first.add(x.agg_list()).agg_concat()).alias(out))
        elif op == "concat":         # :296-301
            final.append(second.add(first.add(x.agg_concat()).agg_concat()).alias(out))
        elif op in ("stddev", "var"):  # :190-258
            xf = x.cast(dt.F64)
            s = second.add(first.add(xf.sum()).sum())
            q = second.add(first.add((xf * xf).sum()).sum())
            c = second.add(first.add(xf.count("valid")).sum())
            n = c.cast(dt.F64)
            mean = s / n
            pop_var = q / n - mean * mean
            ddof = lit(float(inner.value))
            adjusted = pop_var * n / (n - ddof)
            final.append((n <= ddof).if_else(lit(None), adjusted.sqrt() if op == "stddev" else adjusted).alias(out))
        elif op == "skew":           # :302-348
            xf = x.cast(dt.F64)
            n = second.add(first.add(xf.count("valid")).sum())
            s = second.add(first.add(xf.sum()).sum())
            s2 = second.add(first.add((xf * xf).sum()).sum())
            s3 = second.add(first.add((xf * xf * xf).sum()).sum())
            denom_base = (s2 - s * s / n) / n
            denom = (denom_base * denom_base * denom_base).sqrt()
            numerator = s3 - lit(3) * s2 * s / n + lit(2) * s * s * s / n / n
            final.append((numerator / denom / n).alias(out))
        else:
            raise NotImplementedError(f"two-stage rewrite of {op} (agg.rs: Set needs list_distinct; sketches) is not built")
    return first.exprs, second.exprs, final


class GroupedAggregateSink:
    """make_state -> sink(batch)* -> finalize().  The sink runs the first stage on every pushed batch (the fused
    Filter -> Project -> HashAggregate kernel), keeps the partial results in HBM, and finalize() runs the second stage
    and the final projection over them (grouped_aggregate.rs:355-488; ungrouped: sinks/aggregate.rs:54-123)."""

    def __init__(self, aggregations: Sequence[Expression], group_by: Sequence[Expression] | None = None,
                 predicate: Sequence[Expression] | None = None):
        self.aggs, self.group_by, self.predicate = list(aggregations), list(group_by or []), predicate
        self.first, self.second, self.final = populate_aggregation_stages(self.aggs, self.group_by)
        self._h = None
        self._keep = None

    def _create(self, batch):
        names, dts = batch.column_names(), batch.dtypes()
        p1 = encode_agg(self.first, self.group_by, names, dts, self.predicate)
        # schema of the partials: key columns, then the first-stage aggregates (dtypes come back from a zero-row run)
        probe = batch.slice(0, 0).agg(self.first, self.group_by, self.predicate)
        n1, d1 = probe.column_names(), probe.dtypes()
        keys = [col(g.name()) for g in self.group_by]
        p2 = encode_agg(self.second, keys, n1, d1)
        probe2 = probe.agg(self.second, keys)
        n2, d2 = probe2.column_names(), probe2.dtypes()
        p3, self._onames, self._odt = encode_project(self.final, n2, d2)
        self._h = ctypes.c_void_p()
        _lib.call("b200_groupby_create", p1, ctypes.c_size_t(len(p1)), p2, ctypes.c_size_t(len(p2)), p3, ctypes.c_size_t(len(p3)),
                  ctypes.byref(self._h))
        self._keep = (p1, p2, p3)

    def sink(self, batch) -> None:
        if self._h is None:
            self._create(batch)
        _lib.call("b200_groupby_push", self._h, batch._h)

    def finalize(self):
        from .recordbatch import RecordBatch
        if self._h is None:
            raise ValueError("GroupedAggregateSink.finalize() before any batch was pushed")
        out = ctypes.c_void_p()
        _lib.call("b200_groupby_finalize", self._h, ctypes.byref(out))
        return RecordBatch._wrap(out, names=self._onames, dtypes=self._odt)

    def __del__(self):
        try:
            if self._h is not None and self._h.value:
                _lib.lib().b200_groupby_free(self._h)
                self._h = None
        except Exception:
            pass
