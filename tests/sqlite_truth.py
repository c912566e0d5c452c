"""Query-level ground truth the way the reference obtains it (tests/integration/conftest.py:31-53): the generated tables are
loaded into stdlib sqlite3 and the TPC-H SQL of tests/assets/tpch-sqlite-queries/{1,5,6}.sql is run on them.  The SQL text
below is that of those files (the public TPC-H query definitions); dates are stored as ISO strings, which SQLite's date()
compares lexicographically, exactly as the reference's loader does."""
from __future__ import annotations

import datetime
import sqlite3

import pyarrow as pa

Q = {
    1: """select l_returnflag, l_linestatus, sum(l_quantity) as sum_qty, sum(l_extendedprice) as sum_base_price,
              sum(l_extendedprice * (1 - l_discount)) as sum_disc_price, sum(l_extendedprice * (1 - l_discount) * (1 + l_tax)) as sum_charge,
              avg(l_quantity) as avg_qty, avg(l_extendedprice) as avg_price, avg(l_discount) as avg_disc, count(*) as count_order
       from lineitem where l_shipdate <= date('1998-12-01', '-90 days')
       group by l_returnflag, l_linestatus order by l_returnflag, l_linestatus""",
    5: """select n_name, sum(l_extendedprice * (1 - l_discount)) as revenue
       from customer, orders, lineitem, supplier, nation, region
       where c_custkey = o_custkey and l_orderkey = o_orderkey and l_suppkey = s_suppkey and c_nationkey = s_nationkey
         and s_nationkey = n_nationkey and n_regionkey = r_regionkey and r_name = 'ASIA'
         and o_orderdate >= date('1994-01-01') and o_orderdate < date('1994-01-01', '+1 year')
       group by n_name order by revenue desc""",
    6: """select sum(l_extendedprice * l_discount) as revenue from lineitem
       where l_shipdate >= date('1994-01-01') and l_shipdate < date('1994-01-01', '+1 year')
         and l_discount between 0.05 and 0.07 and l_quantity < 24""",
}


def load(tables: dict[str, pa.Table]) -> sqlite3.Connection:
    conn = sqlite3.connect(":memory:")
    for name, tbl in tables.items():
        cols = [c.lower() for c in tbl.column_names]
        conn.execute(f"create table {name} ({', '.join(cols)})")
        data = []
        for c in tbl.columns:
            vals = c.to_pylist()
            if pa.types.is_date(c.type):
                vals = [v.isoformat() if isinstance(v, datetime.date) else v for v in vals]
            data.append(vals)
        conn.executemany(f"insert into {name} values ({', '.join('?' * len(cols))})", list(zip(*data)))
    # the join keys of Q5: without indexes SQLite's nested loops over the generated tables take minutes
    for name, tbl in tables.items():
        for c in tbl.column_names:
            if c.lower().endswith("key"):
                conn.execute(f"create index idx_{name}_{c.lower()} on {name} ({c.lower()})")
    conn.commit()
    return conn


def run(conn: sqlite3.Connection, q: int) -> tuple[list[str], list[tuple]]:
    cur = conn.execute(Q[q])
    return [d[0] for d in cur.description], cur.fetchall()


def assert_matches(tbl: pa.Table, names: list[str], rows: list[tuple], rel: float = 1e-9):
    """ordering asserted (every query ends in ORDER BY), floats within `rel`, everything else exact; column names compared
    case-insensitively (the reference's check renames SQLite's columns to the DataFrame's, conftest.py:45)."""
    import math
    assert [n.lower() for n in tbl.column_names] == [n.lower() for n in names], (tbl.column_names, names)
    got = list(zip(*[c.to_pylist() for c in tbl.columns])) if tbl.num_rows else []
    assert len(got) == len(rows), (len(got), len(rows))
    for i, (g, w) in enumerate(zip(got, rows)):
        for j, (a, b) in enumerate(zip(g, w)):
            if isinstance(b, float) or isinstance(a, float):
                assert a is not None and math.isclose(float(a), float(b), rel_tol=rel, abs_tol=1e-9), (i, names[j], a, b)
            else:
                assert a == b, (i, names[j], a, b)
