"""Device-resident RecordBatch: the host-side mirror of the reference's RecordBatch kernel surface.

Method names, argument meaning and error behaviour follow daft.recordbatch.RecordBatch
(daft/recordbatch/recordbatch.py:235-470 -> src/daft-recordbatch/src/python.rs:38-190) so the parity
tests read like the reference's own tests/recordbatch/*.py.  Every operator is one call through the
C ABI of libdaftb200.so into CUDA kernels; results stay in HBM until `to_arrow()`.
"""
from __future__ import annotations

import ctypes
from typing import Any, Mapping, Sequence

import pyarrow as pa

from . import _lib
from . import datatypes as dt
from .expressions import Expression, col, encode_agg, encode_filter, encode_project, stage_moment_aggs

JOIN_TYPES = {"inner": 0, "left": 1, "right": 2, "outer": 3, "semi": 4, "anti": 5}


def _i32(xs: Sequence[int]):
    return (ctypes.c_int32 * len(xs))(*xs)


def _u8(xs: Sequence[Any]):
    return (ctypes.c_uint8 * len(xs))(*[1 if x else 0 for x in xs])


def _inner(a: Expression) -> Expression:
    while a.kind == "alias":
        a = a.args[0]
    return a


class RecordBatch:
    __slots__ = ("_h", "_schema", "_names", "_dtypes", "__weakref__")

    def __init__(self, handle: int, schema: pa.Schema | None = None):
        self._h = ctypes.c_void_p(handle)
        self._schema = schema  # logical Arrow types (e.g. timestamp units) for round-tripping
        self._names = self._dtypes = None  # batches are immutable: schema queried once

    def __del__(self):
        try:
            if self._h and self._h.value:
                _lib.lib().b200_batch_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # ---- construction ---------------------------------------------------------------------------
    @staticmethod
    def from_arrow_table(tbl: pa.Table) -> "RecordBatch":
        _lib.ensure_init()
        if isinstance(tbl, pa.RecordBatch):
            tbl = pa.Table.from_batches([tbl])
        tbl = tbl.combine_chunks()
        arrays = []
        for c in tbl.columns:
            a = c.chunk(0) if c.num_chunks else pa.array([], type=c.type)
            if pa.types.is_dictionary(a.type):
                a = a.dictionary_decode()
            arrays.append(a)
        rb = pa.RecordBatch.from_arrays(arrays, names=tbl.column_names) if arrays else None
        if rb is None:
            raise ValueError("cannot create a RecordBatch with no columns")
        c_arr, c_sch = _lib.ArrowArray(), _lib.ArrowSchema()
        rb._export_to_c(ctypes.addressof(c_arr), ctypes.addressof(c_sch))
        out = ctypes.c_void_p()
        _lib.call("b200_batch_import", ctypes.byref(c_arr), ctypes.byref(c_sch), ctypes.byref(out))
        return RecordBatch(out.value, tbl.schema)

    from_arrow = from_arrow_table

    @staticmethod
    def from_pydict(data: Mapping[str, Any]) -> "RecordBatch":
        cols = {}
        for k, v in data.items():
            cols[k] = v if isinstance(v, (pa.Array, pa.ChunkedArray)) else pa.array(v)
        return RecordBatch.from_arrow_table(pa.table(cols))

    @staticmethod
    def _wrap(handle, like: "RecordBatch | None" = None, names=None, dtypes=None) -> "RecordBatch":
        """`names` / `dtypes`: the result's schema when the caller already knows it (saves one C call per column)."""
        rb = RecordBatch(handle.value if isinstance(handle, ctypes.c_void_p) else handle, None)
        if like is not None:
            names, dtypes = like._names, like._dtypes
        if names is not None and dtypes is not None:
            rb._names, rb._dtypes = list(names), list(dtypes)
        return rb

    # ---- inspection -------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(_lib.lib().b200_batch_num_rows(self._h))

    def num_columns(self) -> int:
        return int(_lib.lib().b200_batch_num_columns(self._h))

    def size_bytes(self) -> int:
        return int(_lib.lib().b200_batch_size_bytes(self._h))

    def _describe(self, i: int) -> _lib.ColumnDesc:
        d = _lib.ColumnDesc()
        _lib.call("b200_batch_describe", self._h, ctypes.c_int32(i), ctypes.byref(d))
        return d

    def _load_schema(self):
        ds = [self._describe(i) for i in range(self.num_columns())]
        self._names = [d.name.decode() for d in ds]
        self._dtypes = [d.dtype for d in ds]

    def column_names(self) -> list[str]:
        if self._names is None:
            self._load_schema()
        return list(self._names)

    def dtypes(self) -> list[int]:
        if self._dtypes is None:
            self._load_schema()
        return list(self._dtypes)

    def schema(self) -> list[tuple[str, int]]:
        return list(zip(self.column_names(), self.dtypes()))

    # ---- export -----------------------------------------------------------------------------------
    def to_arrow_table(self) -> pa.Table:
        c_arr, c_sch = _lib.ArrowArray(), _lib.ArrowSchema()
        _lib.call("b200_batch_export", self._h, ctypes.byref(c_arr), ctypes.byref(c_sch))
        rb = pa.RecordBatch._import_from_c(ctypes.addressof(c_arr), ctypes.addressof(c_sch))
        return pa.Table.from_batches([rb])

    to_arrow = to_arrow_table

    def to_pydict(self) -> dict[str, list[Any]]:
        return self.to_arrow_table().to_pydict()

    # ---- structural -------------------------------------------------------------------------------
    def select_columns(self, idx: Sequence[int]) -> "RecordBatch":
        out = ctypes.c_void_p()
        _lib.call("b200_batch_select", self._h, _i32(idx), ctypes.c_int32(len(idx)), ctypes.byref(out))
        if self._names is not None:
            return RecordBatch._wrap(out, names=[self._names[i] for i in idx], dtypes=[self._dtypes[i] for i in idx])
        return RecordBatch._wrap(out)

    def rename(self, names: Sequence[str]) -> "RecordBatch":
        old = self.column_names()
        out = self.select_columns(list(range(len(old))))
        for i, n in enumerate(names):
            if n != old[i]:
                _lib.call("b200_batch_rename", out._h, ctypes.c_int32(i), n.encode())
        out._names, out._dtypes = list(names), self.dtypes()
        return out

    def union(self, other: "RecordBatch") -> "RecordBatch":
        out = ctypes.c_void_p()
        _lib.call("b200_batch_union", self._h, other._h, ctypes.byref(out))
        if self._names is not None and other._names is not None:
            return RecordBatch._wrap(out, names=self._names + other._names, dtypes=self._dtypes + other._dtypes)
        return RecordBatch._wrap(out)

    def slice(self, start: int, end: int) -> "RecordBatch":
        if not isinstance(start, int) or not isinstance(end, int):
            raise TypeError("slice bounds must be ints")
        if start < 0 or end < 0 or end < start:
            raise ValueError(f"invalid slice bounds {start}..{end}")
        out = ctypes.c_void_p()
        _lib.call("b200_batch_slice", self._h, ctypes.c_int64(start), ctypes.c_int64(min(end, len(self))), ctypes.byref(out))
        return RecordBatch._wrap(out, like=self)

    def head(self, num: int) -> "RecordBatch":
        return self.slice(0, min(num, len(self)))

    @classmethod
    def concat(cls, to_merge: Sequence["RecordBatch"]) -> "RecordBatch":
        """RecordBatch::concat (error behaviour pinned by tests/recordbatch/test_concat.py:29-73)."""
        if len(to_merge) == 0:
            raise ValueError("Need at least 1 RecordBatch to concat")
        for t in to_merge:
            if not isinstance(t, RecordBatch):
                raise TypeError(f"Expected a RecordBatch for concat, got {type(t).__name__}")
        names = to_merge[0].column_names()
        for t in to_merge[1:]:
            if t.column_names() != names or t.dtypes() != to_merge[0].dtypes():
                raise ValueError("RecordBatch::concat requires all schemas to match")
        arr = (ctypes.c_void_p * len(to_merge))(*[t._h.value for t in to_merge])
        out = ctypes.c_void_p()
        _lib.call("b200_concat", arr, ctypes.c_int32(len(to_merge)), ctypes.byref(out))
        return RecordBatch._wrap(out, like=to_merge[0])

    # ---- expressions --------------------------------------------------------------------------------
    def _resolve_keys(self, exprs: Sequence[Expression]) -> tuple["RecordBatch", list[int]]:
        """Batch + column indices holding the values of `exprs` (plain columns are used in place)."""
        names = self.column_names()
        if all(e.kind == "col" and e.value in names for e in exprs):
            return self, [names.index(e.value) for e in exprs]
        keys = self.eval_expression_list([e.alias(f"__key{i}") for i, e in enumerate(exprs)])
        both = self.union(keys)
        base = len(names)
        return both, [base + i for i in range(len(exprs))]

    # Serialised plan parameters are cached per (expression objects, input schema): an operator instance serialises
    # its plan node once, not once per batch.  Values keep the expressions alive so the
    # ids in the key stay unique.
    _plan_cache: dict = {}

    def _cached_plan(self, kind: str, groups: Sequence[Sequence[Expression] | None], build):
        key = (kind, tuple(None if g is None else tuple(map(id, g)) for g in groups), tuple(self.column_names()), tuple(self.dtypes()))
        hit = RecordBatch._plan_cache.get(key)
        if hit is None:
            if len(RecordBatch._plan_cache) > 512:
                RecordBatch._plan_cache.clear()
            hit = (build(), [list(g) for g in groups if g is not None])
            RecordBatch._plan_cache[key] = hit
        return hit[0]

    def eval_expression_list(self, exprs: Sequence[Expression]) -> "RecordBatch":
        """RecordBatch::eval_expression_list (src/daft-recordbatch/src/lib.rs:1603-1631)."""
        assert all(isinstance(e, Expression) for e in exprs)
        seen = set()
        for e in exprs:
            if e.name() in seen:
                raise ValueError(f"duplicate output column name {e.name()!r}")
            seen.add(e.name())
        # Bare column references (possibly aliased) are zero-copy, as in the reference where Expr::Column evaluates to the
        # same Series (an Arc clone, src/daft-recordbatch/src/lib.rs:1290-1300): only computed expressions reach the kernel.
        names = self.column_names()

        def passthrough(e):
            inner = e
            while inner.kind == "alias":
                inner = inner.args[0]
            return names.index(inner.value) if inner.kind == "col" and inner.value in names else None

        src = [passthrough(e) for e in exprs]
        computed = [e for e, p in zip(exprs, src) if p is None]
        if not computed:
            return self.select_columns(src).rename([e.name() for e in exprs])
        plan, onames, odt = self._cached_plan("project", [computed], lambda: encode_project(computed, names, self.dtypes()))
        out = ctypes.c_void_p()
        _lib.call("b200_eval_expression_list", self._h, plan, ctypes.c_size_t(len(plan)), ctypes.byref(out))
        res = RecordBatch._wrap(out, names=onames, dtypes=odt)
        if len(computed) == len(exprs):
            return res
        both = self.union(res)
        base, k, order = len(names), 0, []
        for p in src:
            if p is None:
                order.append(base + k)
                k += 1
            else:
                order.append(p)
        return both.select_columns(order).rename([e.name() for e in exprs])

    def filter(self, exprs: Sequence[Expression]) -> "RecordBatch":
        """RecordBatch::filter (lib.rs:520-570): conjuncts AND-ed; a null predicate drops the row."""
        assert all(isinstance(e, Expression) for e in exprs)
        if len(exprs) == 0:
            return self
        plan = self._cached_plan("filter", [exprs], lambda: encode_filter(exprs, self.column_names(), self.dtypes()))
        out = ctypes.c_void_p()
        _lib.call("b200_filter", self._h, plan, ctypes.c_size_t(len(plan)), ctypes.byref(out))
        return RecordBatch._wrap(out, like=self)

    def agg(self, to_agg: Sequence[Expression], group_by: Sequence[Expression] | None = None,
            predicate: Sequence[Expression] | None = None) -> "RecordBatch":
        """RecordBatch::agg (src/daft-recordbatch/src/ops/agg.rs:20-88).  `predicate` fuses the Filter that
        feeds the aggregate into the same kernel (rows failing it are skipped; nothing is materialised)."""
        if any(_inner(a).op == "count_distinct" for a in to_agg):
            return self._agg_with_count_distinct(to_agg, group_by or [], predicate)
        if any(self._is_string_minmax(a) for a in to_agg):
            return self._agg_with_string_minmax(to_agg, group_by or [], predicate)
        if any(_inner(a).op == "percentile" for a in to_agg):
            return self._agg_with_percentile(to_agg, group_by or [], predicate)
        staged = stage_moment_aggs(to_agg)
        if staged is not None:  # stddev / var: first-stage sums + counts, then the final projection (agg.rs:190-258)
            first, final = staged
            keys = [col(g.name()) for g in (group_by or [])]
            return self.agg(first, group_by, predicate).eval_expression_list(keys + final)
        plan = self._cached_plan("agg", [to_agg, group_by or [], predicate],
                                 lambda: encode_agg(to_agg, group_by or [], self.column_names(), self.dtypes(), predicate))
        out = ctypes.c_void_p()
        _lib.call("b200_agg", self._h, plan, ctypes.c_size_t(len(plan)), ctypes.byref(out))
        onames = [g.name() for g in (group_by or [])] + [a.name() for a in to_agg]
        odt = (ctypes.c_int32 * len(onames))()
        _lib.call("b200_batch_dtypes", out, odt, ctypes.c_int32(len(onames)))
        return RecordBatch._wrap(out, names=onames, dtypes=list(odt))

    def _is_string_minmax(self, a: Expression) -> bool:
        from .expressions import bind
        inner = _inner(a)
        if inner.kind != "agg" or inner.op not in ("min", "max"):
            return False
        return dt.is_string(bind(inner.args[0], self.column_names(), self.dtypes()).dtype)

    def _agg_with_string_minmax(self, to_agg, group_by, predicate) -> "RecordBatch":
        """min / max over Utf8 / Binary (grouped_cmp_utf8, src/daft-core/src/array/ops/compare_agg.rs:99-134: bytewise
        comparison, nulls ignored, an all-null group gives null).  Composed from the sort and dedup kernels: rows sorted by
        (keys, value) with nulls last, then the first row of every key group.  The numeric aggregates run in the fused
        kernel as usual; both results list one row per group and are aligned by sorting on the keys."""
        base = self.filter(predicate) if predicate else self
        keys = [g.alias(f"__k{i}") for i, g in enumerate(group_by)]
        key_cols = [col(f"__k{i}") for i in range(len(group_by))]
        plain = [a for a in to_agg if not self._is_string_minmax(a)]
        anchor = col(base.column_names()[0]).count("all").alias("__rows_sm")  # keeps the group list complete (its own name: the rewrites nest)
        main = base.agg(plain + [anchor], keys)
        if key_cols:
            main = main.sort(key_cols)
        pieces = main
        for i, a in enumerate(to_agg):
            if not self._is_string_minmax(a):
                continue
            inner = _inner(a)
            vals = base.eval_expression_list(keys + [inner.args[0].alias("__x")])
            if key_cols:
                srt = vals.sort(key_cols + [col("__x")], descending=[False] * len(key_cols) + [inner.op == "max"],
                                nulls_first=[False] * (len(key_cols) + 1))
                best = srt.dedup(key_cols).sort(key_cols)
            else:
                best = vals.sort([col("__x")], descending=[inner.op == "max"], nulls_first=[False]).head(1)
                if len(best) == 0:  # an ungrouped aggregate over no rows still yields one (null) row
                    best = vals.agg([col("__x").any_value().alias("__x")], [])
            pieces = pieces.union(best.eval_expression_list([col("__x").alias(f"__sm{i}")]))
        final = [col(f"__k{j}").alias(g.name()) for j, g in enumerate(group_by)]
        final += [col(f"__sm{i}").alias(a.name()) if self._is_string_minmax(a) else col(a.name()) for i, a in enumerate(to_agg)]
        return pieces.eval_expression_list(final)

    def _agg_with_percentile(self, to_agg, group_by, predicate) -> "RecordBatch":
        """percentile / median per group (exact_percentile, src/daft-core/src/utils/stats.rs:31-59), composed from the sort-based
        window kernels: the non-null values of every group are ranked by value (row_number over (keys; value)); with n values
        rank = p (n - 1), and the answer interpolates between the rows at floor(rank) and ceil(rank) -- found by a filter on the
        row number, folded with min / max per group -- as (hi - lo) * (rank - floor(rank)) + lo.  The reference fuses that last step
        (mul_add): results agree to one ulp.  NaNs sort last, as +NaN does under f64::total_cmp.  Groups without a valid value give
        null (left join back onto the group list, whose order the other aggregates define)."""
        from .expressions import lit
        base = self.filter(predicate) if predicate else self
        # a constant leading key keeps the ungrouped case on the same path (one partition)
        keys = [lit(1).alias("__k0")] + [g.alias(f"__k{i + 1}") for i, g in enumerate(group_by)]
        key_cols = [col(f"__k{i}") for i in range(len(keys))]
        plain = [a for a in to_agg if _inner(a).op != "percentile"]
        keyed = base.union(base.eval_expression_list(keys)) if base.num_columns() else base.eval_expression_list(keys)
        main = keyed.agg(plain + [col("__k0").count("all").alias("__rows_pc")], key_cols)
        if not group_by and len(main) == 0:
            # an ungrouped aggregate over no rows still yields one row: the other aggregates' empty-input values, null percentiles
            return base.agg([a if _inner(a).op != "percentile" else _inner(a).args[0].cast(dt.F64).min().alias(a.name()) for a in to_agg], [])
        group_list = main.eval_expression_list(key_cols)
        pieces = main
        for i, a in enumerate(to_agg):
            inner = _inner(a)
            if inner.op != "percentile":
                continue
            p = float(inner.value)
            vals = keyed.eval_expression_list(key_cols + [inner.args[0].cast(dt.F64).alias("__x")]).filter([col("__x").not_null()])
            w = vals.window(key_cols, [col("__x")], functions=[("row_number", "__rn"), ("agg", "__cnt", col("__x").count())])
            rank = lit(p) * (col("__cnt").cast(dt.F64) - 1.0)
            pos = col("__rn").cast(dt.F64) - 1.0
            w = w.eval_expression_list(key_cols + [col("__x"), rank.alias("__rank"), pos.alias("__pos")])
            kept = w.filter([(col("__pos") == col("__rank").floor()) | (col("__pos") == col("__rank").ceil())])
            g = kept.agg([col("__x").min().alias("__lo"), col("__x").max().alias("__hi"), col("__rank").max().alias("__r")], key_cols)
            val = ((col("__hi") - col("__lo")) * (col("__r") - col("__r").floor()) + col("__lo")).alias(f"__pc{i}")
            g = g.eval_expression_list(key_cols + [val])
            j = group_list.hash_join(g, key_cols, key_cols, how="left", null_equals_null=[True] * len(key_cols))
            pieces = pieces.union(j.eval_expression_list([col(f"__pc{i}")]))
        final = [col(f"__k{j + 1}").alias(g.name()) for j, g in enumerate(group_by)]
        final += [col(f"__pc{i}").alias(a.name()) if _inner(a).op == "percentile" else col(a.name()) for i, a in enumerate(to_agg)]
        return pieces.eval_expression_list(final)

    def _agg_with_count_distinct(self, to_agg, group_by, predicate) -> "RecordBatch":
        """count_distinct(x) per group = dedup on (keys, x), then count the non-null x per key.  Dedup keeps first
        occurrences in row order, so every partial result lists the groups in the same first-occurrence order and the
        columns are simply placed side by side (the reference builds per-group lists and one probe table per group,
        src/daft-functions-list/src/series.rs:37-40,234-247)."""
        base = self.filter(predicate) if predicate else self
        keys = [g.alias(f"__k{i}") for i, g in enumerate(group_by)]
        key_cols = [col(f"__k{i}") for i in range(len(group_by))]
        plain = [a for a in to_agg if _inner(a).op != "count_distinct"]
        nested = any(_inner(a).op == "percentile" or self._is_string_minmax(a) for a in plain)
        # a count(all) keeps the group list complete even when every other aggregate is a count_distinct
        parts = base.agg(plain + [col(base.column_names()[0]).count("all").alias("__rows_cd")] if base.num_columns() else plain, group_by)
        out_cols, pieces = {}, parts
        for i, a in enumerate(to_agg):
            if _inner(a).op != "count_distinct":
                continue
            x = _inner(a).args[0]
            pairs = base.eval_expression_list(keys + [x.alias("__x")]).dedup(key_cols + [col("__x")])
            cnt = pairs.agg([col("__x").count("valid").alias(f"__cd{i}")], key_cols)
            if nested and group_by:
                # the other aggregates went through a rewrite of their own (string min / max sorts its groups by key): the two group
                # lists are the same set in different orders -- align this column on the keys instead of placing it side by side
                glist = parts.eval_expression_list([col(g.name()).alias(f"__k{j}") for j, g in enumerate(group_by)])
                cnt = glist.hash_join(cnt, key_cols, key_cols, how="left", null_equals_null=[True] * len(key_cols))
            pieces = pieces.union(cnt.eval_expression_list([col(f"__cd{i}")]))
            out_cols[i] = f"__cd{i}"
        final = [col(g.name()) for g in group_by]
        final += [col(out_cols[i]).alias(a.name()) if i in out_cols else col(a.name()) for i, a in enumerate(to_agg)]
        return pieces.eval_expression_list(final)

    def dedup(self, exprs: Sequence[Expression]) -> "RecordBatch":
        if len(exprs) == 0:
            raise ValueError("Attempting to dedup RecordBatch on no columns")
        both, idx = self._resolve_keys(exprs)
        out = ctypes.c_void_p()
        _lib.call("b200_dedup", both._h, _i32(idx), ctypes.c_int32(len(idx)), ctypes.byref(out))
        return RecordBatch._wrap(out).select_columns(list(range(self.num_columns())))

    # ---- hashing / partitioning ---------------------------------------------------------------------
    def hash_rows(self, exprs: Sequence[Expression] | None = None, seed: "RecordBatch | None" = None) -> "RecordBatch":
        """RecordBatch::hash_rows (src/daft-recordbatch/src/ops/hash.rs:14-26) -> one UInt64 column."""
        exprs = exprs if exprs is not None else [col(n) for n in self.column_names()]
        both, idx = self._resolve_keys(exprs)
        out = ctypes.c_void_p()
        _lib.call("b200_hash_rows", both._h, _i32(idx), ctypes.c_int32(len(idx)), seed._h if seed is not None else None,
                  ctypes.byref(out))
        return RecordBatch._wrap(out)

    def partition_by_hash(self, exprs: Sequence[Expression], num_partitions: int) -> list["RecordBatch"]:
        """RecordBatch::partition_by_hash (src/daft-recordbatch/src/ops/partition.rs:51-70)."""
        if not isinstance(num_partitions, int):
            raise TypeError(f"Expected a num_partitions to be int, got {type(num_partitions)}")
        if num_partitions <= 0:
            raise ValueError("Can not partition a Table by 0 partitions" if num_partitions == 0 else "num_partitions must be positive")
        both, idx = self._resolve_keys(exprs)
        outs = (ctypes.c_void_p * num_partitions)()
        _lib.call("b200_partition_by_hash", both._h, _i32(idx), ctypes.c_int32(len(idx)), ctypes.c_int32(num_partitions), outs)
        keep = list(range(self.num_columns()))
        return [RecordBatch._wrap(ctypes.c_void_p(h)).select_columns(keep) for h in outs]

    def partition_by_random(self, num_partitions: int, seed: int) -> list["RecordBatch"]:
        """RecordBatch::partition_by_random (src/daft-recordbatch/src/ops/partition.rs:72-86): deterministic per (seed, input)."""
        if not isinstance(num_partitions, int):
            raise TypeError(f"Expected a num_partitions to be int, got {type(num_partitions)}")
        if not isinstance(seed, int):
            raise TypeError(f"Expected a seed to be int, got {type(seed)}")
        if num_partitions <= 0:
            raise ValueError("Can not partition a Table by 0 partitions" if num_partitions == 0 else "num_partitions must be positive")
        outs = (ctypes.c_void_p * num_partitions)()
        _lib.call("b200_partition_by_random", self._h, ctypes.c_int32(num_partitions), ctypes.c_uint64(seed & (2**64 - 1)), outs)
        return [RecordBatch._wrap(ctypes.c_void_p(h), like=self) for h in outs]

    def quantiles(self, num: int) -> "RecordBatch":
        """RecordBatch::quantiles (src/daft-recordbatch/src/lib.rs:491-513): rows floor(i / num * len), i = 1..num-1 -- the range
        boundaries of a distributed sort when `self` is the sorted sample."""
        n = len(self)
        if n == 0:
            return self
        pts = [min(int((i / num) * n), n - 1) for i in range(1, num)] if num > 0 else []
        return self.take(pa.array(pts, type=pa.int64()))

    def search_sorted(self, keys: "RecordBatch", descending: Sequence[bool] | bool = False) -> "RecordBatch":
        """RecordBatch::search_sorted (src/daft-recordbatch/src/ops/search_sorted.rs:13-36): self is the sorted table."""
        if self.dtypes() != keys.dtypes():
            raise ValueError(f"Schema Mismatch in search_sorted: data: {self.schema()} vs keys: {keys.schema()}")
        desc = [descending] * self.num_columns() if isinstance(descending, bool) else list(descending)
        if len(desc) != self.num_columns():
            raise ValueError("Mismatch in number of arguments for `descending` in search sorted")
        out = ctypes.c_void_p()
        _lib.call("b200_search_sorted", self._h, keys._h, _i32(list(range(self.num_columns()))), ctypes.c_int32(self.num_columns()), _u8(desc),
                  ctypes.byref(out))
        return RecordBatch._wrap(out)

    def partition_by_range(self, partition_keys: Sequence[Expression], boundaries: "RecordBatch", descending: Sequence[bool]) -> list["RecordBatch"]:
        """RecordBatch::partition_by_range (src/daft-recordbatch/src/ops/partition.rs:88-101)."""
        if not isinstance(boundaries, RecordBatch):
            raise TypeError(f"Expected a RecordBatch for `boundaries` in partition_by_range but got {type(boundaries)}")
        if len(boundaries) == 0:
            return [self]
        from .expressions import bind
        names, dtypes = self.column_names(), self.dtypes()
        key_schema = [(e.name(), bind(e, names, dtypes).dtype) for e in partition_keys]
        if key_schema != boundaries.schema():  # names and dtypes first, as RecordBatch::search_sorted does (search_sorted.rs:14-26)
            raise ValueError(f"Schema Mismatch in search_sorted: data: {boundaries.schema()} vs keys: {key_schema}")
        if len(descending) != boundaries.num_columns():
            raise ValueError(f"Mismatch in number of arguments for `descending` in search sorted: num_columns: "
                             f"{boundaries.num_columns()} vs : descending.len() {len(descending)}")
        both, idx = self._resolve_keys(partition_keys)
        nparts = len(boundaries) + 1
        outs = (ctypes.c_void_p * nparts)()
        _lib.call("b200_partition_by_range", both._h, _i32(idx), ctypes.c_int32(len(idx)), boundaries._h, _u8(list(descending)), outs)
        keep = list(range(self.num_columns()))
        return [RecordBatch._wrap(ctypes.c_void_p(h)).select_columns(keep) for h in outs]

    # ---- take / sort ------------------------------------------------------------------------------------
    def take(self, indices: "RecordBatch | Sequence[int] | pa.Array") -> "RecordBatch":
        if not isinstance(indices, RecordBatch):
            arr = indices if isinstance(indices, pa.Array) else pa.array(indices, type=pa.int64())
            indices = RecordBatch.from_arrow_table(pa.table({"idx": arr}))
        out = ctypes.c_void_p()
        _lib.call("b200_take", self._h, indices._h, ctypes.byref(out))
        return RecordBatch._wrap(out, like=self)

    @staticmethod
    def _sort_flags(sort_keys, descending, nulls_first):
        n = len(sort_keys)
        if descending is None:
            descending = [False] * n
        elif isinstance(descending, bool):
            descending = [descending] * n
        elif isinstance(descending, list):
            if len(descending) != n:
                raise ValueError(f"Expected length of `descending` to be the same length as `sort_keys` since a list was passed in,"
                                 f"got {len(descending)} instead of {n}")
        else:
            raise TypeError(f"Expected a bool, list[bool] or None for `descending` but got {type(descending)}")
        if nulls_first is None:
            nulls_first = list(descending)
        elif isinstance(nulls_first, bool):
            nulls_first = [nulls_first] * n
        elif isinstance(nulls_first, list):
            if len(nulls_first) != n:
                raise ValueError(f"Expected length of `nulls_first` to be the same length as `sort_keys` since a list was passed in,"
                                 f"got {len(nulls_first)} instead of {n}")
        else:
            raise TypeError(f"Expected a bool, list[bool] or None for `nulls_first` but got {type(nulls_first)}")
        return [bool(x) for x in descending], [bool(x) for x in nulls_first]

    def argsort(self, sort_keys: Sequence[Expression], descending=None, nulls_first=None) -> "RecordBatch":
        """RecordBatch::argsort (src/daft-recordbatch/src/ops/sort.rs:7-45) -> UInt64 indices."""
        desc, nf = self._sort_flags(sort_keys, descending, nulls_first)
        both, idx = self._resolve_keys(sort_keys)
        out = ctypes.c_void_p()
        _lib.call("b200_argsort", both._h, _i32(idx), ctypes.c_int32(len(idx)), _u8(desc), _u8(nf), ctypes.byref(out))
        return RecordBatch._wrap(out)

    def sort(self, sort_keys: Sequence[Expression], descending=None, nulls_first=None) -> "RecordBatch":
        desc, nf = self._sort_flags(sort_keys, descending, nulls_first)
        both, idx = self._resolve_keys(sort_keys)
        out = ctypes.c_void_p()
        _lib.call("b200_sort", both._h, _i32(idx), ctypes.c_int32(len(idx)), _u8(desc), _u8(nf), ctypes.byref(out))
        res = RecordBatch._wrap(out, like=both)
        return res if both is self else res.select_columns(list(range(self.num_columns())))

    def top_n(self, sort_keys: Sequence[Expression], limit: int, descending=None, nulls_first=None, offset: int = 0) -> "RecordBatch":
        desc, nf = self._sort_flags(sort_keys, descending, nulls_first)
        both, idx = self._resolve_keys(sort_keys)
        out = ctypes.c_void_p()
        _lib.call("b200_top_n", both._h, _i32(idx), ctypes.c_int32(len(idx)), _u8(desc), _u8(nf), ctypes.c_int64(limit),
                  ctypes.c_int64(offset), ctypes.byref(out))
        return RecordBatch._wrap(out).select_columns(list(range(self.num_columns())))

    # ---- join ---------------------------------------------------------------------------------------------
    def hash_join(self, right: "RecordBatch", left_on: Sequence[Expression], right_on: Sequence[Expression], how: str = "inner",
                  null_equals_null: Sequence[bool] | None = None, build_on_left: bool | None = None) -> "RecordBatch":
        """RecordBatch::hash_join (src/daft-recordbatch/src/ops/joins/mod.rs) with the operator-level output
        schema of src/daft-local-execution/src/join/inner_join.rs:50-78: join keys that share a name (taken
        from the left), left non-key columns, right non-key columns (clashing names get a `right.` prefix)."""
        from .join import hash_join  # late import: join.py builds on this class
        return hash_join(self, right, left_on, right_on, how, null_equals_null, build_on_left)

    def sort_merge_join(self, right: "RecordBatch", left_on: Sequence[Expression], right_on: Sequence[Expression],
                        is_sorted: bool = False) -> "RecordBatch":
        """RecordBatch::sort_merge_join (src/daft-recordbatch/src/ops/joins/mod.rs:154-229)."""
        from .join import sort_merge_join
        return sort_merge_join(self, right, left_on, right_on, is_sorted)

    def asof_join(self, right: "RecordBatch", left_by: Sequence[Expression], right_by: Sequence[Expression], left_on: Expression,
                  right_on: Expression) -> "RecordBatch":
        """MicroPartition::asof_join, strategy backward (src/daft-micropartition/src/ops/join.rs:127-187)."""
        from .join import asof_join
        return asof_join(self, right, left_by, right_by, left_on, right_on)

    def cross_join(self, right: "RecordBatch", outer_loop_side: str = "left") -> "RecordBatch":
        """RecordBatch::cross_join (src/daft-recordbatch/src/ops/joins/mod.rs:231-284)."""
        from .join import cross_join
        return cross_join(self, right, outer_loop_side)

    def _window_ranks(self, partition_cols: Sequence[int], order_cols: Sequence[int]) -> "RecordBatch":
        """b200_window_ranks over this batch, which is sorted by (partition columns, order columns)."""
        out = ctypes.c_void_p()
        _lib.call("b200_window_ranks", self._h, _i32(partition_cols), ctypes.c_int32(len(partition_cols)), _i32(order_cols),
                  ctypes.c_int32(len(order_cols)), ctypes.byref(out))
        return RecordBatch._wrap(out, names=["__row_number", "__rank", "__dense_rank", "__partition_begin", "__partition_id"],
                                 dtypes=[dt.U64, dt.U64, dt.U64, dt.I64, dt.I64])

    def _window_offset(self, partition_ids: "RecordBatch", offset: int) -> "RecordBatch":
        """b200_window_offset: this batch's columns read `offset` rows away inside the partition, plus `__src` (the row read, -1 = none)."""
        out = ctypes.c_void_p()
        _lib.call("b200_window_offset", self._h, partition_ids._h, ctypes.c_int64(offset), ctypes.byref(out))
        return RecordBatch._wrap(out, names=self.column_names() + ["__src"], dtypes=self.dtypes() + [dt.I64])

    def _window_frame(self, partition_begin: "RecordBatch", agg: int, start, end, min_periods: int, out_dtype: int, kind: int = 1,
                      order: "RecordBatch | None" = None, descending: bool = False) -> "RecordBatch":
        """b200_window_frame: aggregate `agg` (include/daft_b200.h) of this batch's one column -- already in the accumulate type -- over the
        frame [start, end] (None = unbounded) of every row; kind 1: row offsets, 2: offsets on the value of `order` (one Int64 column);
        `self` may have no columns for count(*)."""
        out = ctypes.c_void_p()
        _lib.call("b200_window_frame", self._h if self.num_columns() else None, partition_begin._h, order._h if order is not None else None,
                  ctypes.c_int32(agg), ctypes.c_int32(kind if start is not None else 0), ctypes.c_int64(start or 0),
                  ctypes.c_int32(kind if end is not None else 0), ctypes.c_int64(end or 0), ctypes.c_int32(bool(descending)), ctypes.c_int64(min_periods),
                  ctypes.byref(out))
        return RecordBatch._wrap(out, names=[self.column_names()[0] if self.num_columns() else "count"], dtypes=[out_dtype])

    def window(self, partition_by: Sequence[Expression], order_by: Sequence[Expression] = (), descending=None, nulls_first=None,
               functions: Sequence[tuple] = ()) -> "RecordBatch":
        """row_number / rank / dense_rank / lead / lag / partition aggregates / aggregates over ROWS frames (daft_b200/window.py)."""
        from .window import window
        return window(self, partition_by, order_by, descending, nulls_first, functions)

    def pivot(self, group_by: Sequence[Expression], pivot_col: Expression, values_col: Expression, names: Sequence[str]) -> "RecordBatch":
        """RecordBatch::pivot (src/daft-recordbatch/src/ops/pivot.rs:70-150)."""
        from .window import pivot
        return pivot(self, group_by, pivot_col, values_col, names)

    def agg_pivot(self, group_by: Sequence[Expression], pivot_col: Expression, value_col: Expression, agg: str, names: Sequence[str]) -> "RecordBatch":
        """PivotSink: aggregate per (group, pivot value), then pivot (src/daft-local-execution/src/sinks/pivot.rs)."""
        from .window import agg_pivot
        return agg_pivot(self, group_by, pivot_col, value_col, agg, names)

