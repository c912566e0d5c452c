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
    3: """select l_orderkey, sum(l_extendedprice * (1 - l_discount)) as revenue, o_orderdate, o_shippriority
       from customer, orders, lineitem
       where c_mktsegment = 'BUILDING' and c_custkey = o_custkey and l_orderkey = o_orderkey and o_orderdate < date('1995-03-15')
         and l_shipdate > date('1995-03-15')
       group by l_orderkey, o_orderdate, o_shippriority order by revenue desc, o_orderdate limit 10""",
    4: """select o_orderpriority, count(*) as order_count from orders
       where o_orderdate >= date('1993-07-01') and o_orderdate < date('1993-07-01', '+3 months')
         and exists (select * from lineitem where l_orderkey = o_orderkey and l_commitdate < l_receiptdate)
       group by o_orderpriority order by o_orderpriority""",
    12: """select l_shipmode,
              sum(case when o_orderpriority = '1-URGENT' or o_orderpriority = '2-HIGH' then 1 else 0 end) as high_line_count,
              sum(case when o_orderpriority <> '1-URGENT' and o_orderpriority <> '2-HIGH' then 1 else 0 end) as low_line_count
       from orders, lineitem
       where o_orderkey = l_orderkey and l_shipmode in ('MAIL', 'SHIP') and l_commitdate < l_receiptdate and l_shipdate < l_commitdate
         and l_receiptdate >= date('1994-01-01') and l_receiptdate < date('1994-01-01', '+1 year')
       group by l_shipmode order by l_shipmode""",
    14: """select 100.00 * sum(case when p_type like 'PROMO%' then l_extendedprice * (1 - l_discount) else 0 end)
                  / sum(l_extendedprice * (1 - l_discount)) as promo_revenue
       from lineitem, part
       where l_partkey = p_partkey and l_shipdate >= date('1995-09-01') and l_shipdate < date('1995-09-01', '+1 month')""",
    18: """select c_custkey, o_orderkey, o_orderdate, o_totalprice, sum(l_quantity)
       from customer, orders, lineitem
       where o_orderkey in (select l_orderkey from lineitem group by l_orderkey having sum(l_quantity) > 300)
         and c_custkey = o_custkey and o_orderkey = l_orderkey
       group by c_custkey, o_orderkey, o_orderdate, o_totalprice order by o_totalprice desc, o_orderdate limit 100""",
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


def assert_matches(tbl: pa.Table, names: list[str], rows: list[tuple], rel: float = 1e-9, check_names: bool = True):
    """ordering asserted (every query ends in ORDER BY), floats within `rel`, everything else exact; column names compared
    case-insensitively (the reference's check renames SQLite's columns to the DataFrame's, conftest.py:45)."""
    import math
    if check_names:
        assert [n.lower() for n in tbl.column_names] == [n.lower() for n in names], (tbl.column_names, names)
    assert tbl.num_columns == len(names)
    cols = []
    for c in tbl.columns:
        vals = c.to_pylist()
        if pa.types.is_date(c.type):  # SQLite holds dates as ISO strings
            vals = [v.isoformat() if v is not None else None for v in vals]
        cols.append(vals)
    got = list(zip(*cols)) if tbl.num_rows else []
    assert len(got) == len(rows), (len(got), len(rows))
    for i, (g, w) in enumerate(zip(got, rows)):
        for j, (a, b) in enumerate(zip(g, w)):
            if isinstance(b, float) or isinstance(a, float):
                assert a is not None and math.isclose(float(a), float(b), rel_tol=rel, abs_tol=1e-9), (i, names[j], a, b)
            else:
                assert a == b, (i, names[j], a, b)
