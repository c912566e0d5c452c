"""ClickBench (`benchmarking/clickbench/queries.sql`, BASELINE config 5) on the RecordBatch operator surface.

The reference runs these 43 statements through `daft.sql` over the `hits` table (`tests/benchmarks/test_clickbench.py:28-66`: EventDate
cast to Date, EventTime to Timestamp); each lowers to Filter -> Project -> (Hash)Aggregate -> TopN / Sort, which is what the functions
below compose by hand.  All 43 are built (Q28's `REGEXP_REPLACE(..., '\\1')` runs in the device's regex VM, daft_b200/regex.py).

`generate_hits` is a synthetic stand-in for the 100 M-row public dataset (no network here): the columns the queries read, with the
skew that makes them interesting (Zipf-like users / phrases / URLs, mostly-empty SearchPhrase and MobilePhoneModel, a dominant
CounterID 62 in July 2013).  It is a HOST generator (numpy): the queries have been verified on the CPU only -- oracle twin against SQLite,
generated code through the host emulation (tests/test_clickbench_cpu.py) -- and have not been run on the device this round.

Thresholds that only make sense at 100 M rows (HAVING COUNT(*) > 100000, OFFSET 1000 ...) are parameters with the benchmark's values as
defaults."""
from __future__ import annotations

import datetime

import numpy as np
import pyarrow as pa

from .expressions import col, lit

EPOCH = datetime.date(1970, 1, 1)


def generate_hits(n: int, seed: int = 7) -> pa.Table:
    rng = np.random.default_rng(seed)

    def zipf(k: int, a: float = 1.3):
        return np.minimum(rng.zipf(a, n), k) - 1

    users = rng.integers(1, 2**35, 5000)   # AVG(UserID) sums them in Int64, as the reference does: keep the sum exact
    phrases = np.array([""] + [f"phrase {i} {'x' * (i % 9)}" for i in range(1, 400)], dtype=object)
    sp_idx = np.where(rng.random(n) < 0.8, 0, 1 + zipf(399))
    hosts = np.array(["google.com", "yandex.ru", "mail.google.com", "example.org", "news.site.net", "shop.example.com"], dtype=object)
    url_idx = zipf(3000, 1.2)
    urls = np.array([""] + [f"http://{hosts[i % len(hosts)]}/path/{i}?q={i % 17}" for i in range(1, 3001)], dtype=object)
    titles = np.array([""] + [f"{'Google' if i % 5 == 0 else 'Page'} title {i}" for i in range(1, 800)], dtype=object)
    models = np.array(["", "iPad", "iPhone", "Nokia 3310", "GT-I9300", "Lumia"], dtype=object)
    day0 = (datetime.date(2013, 7, 1) - EPOCH).days
    days = day0 + rng.integers(0, 31, n)
    secs = rng.integers(0, 86400, n)
    counter = np.where(rng.random(n) < 0.3, 62, rng.integers(1, 500, n)).astype(np.int32)
    url_col = urls[url_idx]
    return pa.table({
        "WatchID": pa.array(rng.integers(1, 2**62, n)),
        "Title": pa.array(titles[zipf(799, 1.25)], type=pa.large_string()),
        "EventTime": pa.array((days.astype(np.int64) * 86400 + secs) * 1_000_000, type=pa.timestamp("us")),
        "EventDate": pa.array(days.astype(np.int32), type=pa.date32()),
        "CounterID": pa.array(counter),
        "ClientIP": pa.array(rng.integers(-2**31, 2**31, 3000).astype(np.int32)[zipf(2999, 1.15)]),
        "RegionID": pa.array(zipf(300, 1.4).astype(np.int32)),
        "UserID": pa.array(users[zipf(4999, 1.2)]),
        "URL": pa.array(url_col, type=pa.large_string()),
        "Referer": pa.array(urls[zipf(3000, 1.3)], type=pa.large_string()),
        "IsRefresh": pa.array((rng.random(n) < 0.1).astype(np.int16)),
        "ResolutionWidth": pa.array(rng.choice(np.array([0, 1024, 1280, 1366, 1920], dtype=np.int16), n)),
        "MobilePhone": pa.array(rng.integers(0, 6, n).astype(np.int16)),
        "MobilePhoneModel": pa.array(models[np.where(rng.random(n) < 0.85, 0, rng.integers(1, 6, n))], type=pa.large_string()),
        "SearchPhrase": pa.array(phrases[sp_idx], type=pa.large_string()),
        "SearchEngineID": pa.array(np.where(sp_idx == 0, 0, rng.integers(1, 5, n)).astype(np.int16)),
        "AdvEngineID": pa.array(np.where(rng.random(n) < 0.95, 0, rng.integers(1, 20, n)).astype(np.int16)),
        "TraficSourceID": pa.array(rng.choice(np.array([-1, 0, 1, 2, 3, 6], dtype=np.int16), n)),
        "IsLink": pa.array((rng.random(n) < 0.2).astype(np.int16)),
        "IsDownload": pa.array((rng.random(n) < 0.05).astype(np.int16)),
        "DontCountHits": pa.array((rng.random(n) < 0.1).astype(np.int16)),
        "WindowClientWidth": pa.array(rng.choice(np.array([0, 1000, 1200, 1600], dtype=np.int16), n)),
        "WindowClientHeight": pa.array(rng.choice(np.array([0, 600, 800, 900], dtype=np.int16), n)),
        "URLHash": pa.array((url_idx.astype(np.int64) * 2654435761) % (2**61)),
        "RefererHash": pa.array(rng.integers(0, 50, n).astype(np.int64) * 72057594037927 + 3),
    })


def _count(name="c"):
    return col("WatchID").count("all").alias(name)


def _top(rb, key: str, k: int, descending: bool = True, offset: int = 0):
    return rb.top_n([col(key)], k, descending=[descending], offset=offset)


JULY = [col("EventDate") >= datetime.date(2013, 7, 1), col("EventDate") <= datetime.date(2013, 7, 31)]


def q0(h): return h.agg([_count("count")], [])
def q1(h): return h.agg([_count("count")], [], predicate=[col("AdvEngineID") != 0])
def q2(h): return h.agg([col("AdvEngineID").sum().alias("s"), _count("c"), col("ResolutionWidth").mean().alias("a")], [])
def q3(h): return h.agg([col("UserID").mean().alias("a")], [])
def q4(h): return h.agg([col("UserID").count_distinct().alias("u")], [])
def q5(h): return h.agg([col("SearchPhrase").count_distinct().alias("u")], [])
def q6(h): return h.agg([col("EventDate").min().alias("m1"), col("EventDate").max().alias("m2")], [])


def q7(h):
    return h.agg([_count("c")], [col("AdvEngineID")], predicate=[col("AdvEngineID") != 0]).sort([col("c")], descending=True)


def q8(h): return _top(h.agg([col("UserID").count_distinct().alias("u")], [col("RegionID")]), "u", 10)


def q9(h):
    g = h.agg([col("AdvEngineID").sum().alias("s"), _count("c"), col("ResolutionWidth").mean().alias("a"), col("UserID").count_distinct().alias("u")],
              [col("RegionID")])
    return _top(g, "c", 10)


def q10(h):
    g = h.filter([col("MobilePhoneModel") != ""]).agg([col("UserID").count_distinct().alias("u")], [col("MobilePhoneModel")])
    return _top(g, "u", 10)


def q11(h):
    g = h.filter([col("MobilePhoneModel") != ""]).agg([col("UserID").count_distinct().alias("u")], [col("MobilePhone"), col("MobilePhoneModel")])
    return _top(g, "u", 10)


def q12(h): return _top(h.agg([_count("c")], [col("SearchPhrase")], predicate=[col("SearchPhrase") != ""]), "c", 10)


def q13(h):
    g = h.filter([col("SearchPhrase") != ""]).agg([col("UserID").count_distinct().alias("u")], [col("SearchPhrase")])
    return _top(g, "u", 10)


def q14(h): return _top(h.agg([_count("c")], [col("SearchEngineID"), col("SearchPhrase")], predicate=[col("SearchPhrase") != ""]), "c", 10)
def q15(h): return _top(h.agg([_count("c")], [col("UserID")]), "c", 10)
def q16(h): return _top(h.agg([_count("c")], [col("UserID"), col("SearchPhrase")]), "c", 10)
def q17(h): return h.agg([_count("c")], [col("UserID"), col("SearchPhrase")]).head(10)
def q18(h):
    g = h.eval_expression_list([col("UserID"), col("EventTime").minute().alias("m"), col("SearchPhrase"), col("WatchID")])
    return _top(g.agg([_count("c")], [col("UserID"), col("m"), col("SearchPhrase")]), "c", 10)


def q19(h, user_id: int = 435090932899640449): return h.filter([col("UserID") == user_id]).eval_expression_list([col("UserID")])
def q20(h): return h.agg([_count("count")], [], predicate=[col("URL").like("%google%")])


def q21(h):
    g = h.filter([col("URL").like("%google%"), col("SearchPhrase") != ""]).agg([col("URL").min().alias("m"), _count("c")], [col("SearchPhrase")])
    return _top(g, "c", 10)


def q22(h):
    f = h.filter([col("Title").like("%Google%"), ~col("URL").like("%.google.%"), col("SearchPhrase") != ""])
    g = f.agg([col("URL").min().alias("mu"), col("Title").min().alias("mt"), _count("c"), col("UserID").count_distinct().alias("u")], [col("SearchPhrase")])
    return _top(g, "c", 10)


def q23(h): return h.filter([col("URL").like("%google%")]).top_n([col("EventTime")], 10, descending=[False])


def q24(h):
    return h.filter([col("SearchPhrase") != ""]).top_n([col("EventTime")], 10, descending=[False]).eval_expression_list([col("SearchPhrase")])


def q25(h):
    return h.filter([col("SearchPhrase") != ""]).eval_expression_list([col("SearchPhrase")]).top_n([col("SearchPhrase")], 10, descending=[False])


def q26(h):
    f = h.filter([col("SearchPhrase") != ""]).top_n([col("EventTime"), col("SearchPhrase")], 10, descending=[False, False])
    return f.eval_expression_list([col("SearchPhrase")])


def q27(h, having: int = 100000):
    g = h.agg([col("URL").length().mean().alias("l"), _count("c")], [col("CounterID")], predicate=[col("URL") != ""])
    return g.filter([col("c") > having]).top_n([col("l")], 25, descending=[True])


def q28(h, having: int = 100000):
    k = col("Referer").regexp_replace(r"^https?://(?:www\.)?([^/]+)/.*$", r"\1").alias("k")
    f = h.filter([col("Referer") != ""]).eval_expression_list([k, col("Referer"), col("WatchID")])
    g = f.agg([col("Referer").length().mean().alias("l"), _count("c"), col("Referer").min().alias("m")], [col("k")])
    return g.filter([col("c") > having]).top_n([col("l")], 25, descending=[True])


def q29(h): return h.agg([(col("ResolutionWidth") + i).sum().alias(f"s{i}") if i else col("ResolutionWidth").sum().alias("s0") for i in range(90)], [])


def _traffic(h, keys, pred=None):
    g = h.agg([_count("c"), col("IsRefresh").sum().alias("r"), col("ResolutionWidth").mean().alias("a")], keys, predicate=pred)
    return _top(g, "c", 10)


def q30(h): return _traffic(h, [col("SearchEngineID"), col("ClientIP")], [col("SearchPhrase") != ""])
def q31(h): return _traffic(h, [col("WatchID"), col("ClientIP")], [col("SearchPhrase") != ""])
def q32(h): return _traffic(h, [col("WatchID"), col("ClientIP")])
def q33(h): return _top(h.agg([_count("c")], [col("URL")]), "c", 10)


def q34(h):
    g = h.agg([_count("c")], [lit(1).alias("one"), col("URL")])
    return _top(g, "c", 10)


def q35(h):
    keys = [col("ClientIP"), (col("ClientIP") - 1).alias("ip1"), (col("ClientIP") - 2).alias("ip2"), (col("ClientIP") - 3).alias("ip3")]
    return _top(h.agg([_count("c")], keys), "c", 10)


def _pageviews(h, key, preds, offset=0, k=10):
    g = h.agg([_count("PageViews")], key, predicate=[col("CounterID") == 62] + JULY + preds)
    return _top(g, "PageViews", k, offset=offset)


def q36(h): return _pageviews(h, [col("URL")], [col("DontCountHits") == 0, col("IsRefresh") == 0, col("URL") != ""])
def q37(h): return _pageviews(h, [col("Title")], [col("DontCountHits") == 0, col("IsRefresh") == 0, col("Title") != ""])


def q38(h, offset: int = 1000):
    return _pageviews(h, [col("URL")], [col("IsRefresh") == 0, col("IsLink") != 0, col("IsDownload") == 0], offset=offset)


def q39(h, offset: int = 1000):
    src = ((col("SearchEngineID") == 0) & (col("AdvEngineID") == 0)).if_else(col("Referer"), lit("")).alias("Src")
    f = h.filter([col("CounterID") == 62] + JULY + [col("IsRefresh") == 0])
    f = f.eval_expression_list([col("TraficSourceID"), col("SearchEngineID"), col("AdvEngineID"), src, col("URL").alias("Dst"), col("WatchID")])
    g = f.agg([_count("PageViews")], [col("TraficSourceID"), col("SearchEngineID"), col("AdvEngineID"), col("Src"), col("Dst")])
    return _top(g, "PageViews", 10, offset=offset)


def q40(h, offset: int = 100, referer_hash: int = 3594120000172545465):
    return _pageviews(h, [col("URLHash"), col("EventDate")], [col("IsRefresh") == 0, col("TraficSourceID").is_in([-1, 6]), col("RefererHash") == referer_hash],
                      offset=offset)


def q41(h, offset: int = 10000, url_hash: int = 2868770270353813622):
    return _pageviews(h, [col("WindowClientWidth"), col("WindowClientHeight")], [col("IsRefresh") == 0, col("DontCountHits") == 0, col("URLHash") == url_hash],
                      offset=offset)


def q42(h, offset: int = 1000):
    f = h.filter([col("CounterID") == 62, col("EventDate") >= datetime.date(2013, 7, 14), col("EventDate") <= datetime.date(2013, 7, 15), col("IsRefresh") == 0,
                  col("DontCountHits") == 0])
    g = f.eval_expression_list([col("EventTime").truncate("minute").alias("M"), col("WatchID")]).agg([_count("PageViews")], [col("M")])
    return g.top_n([col("M")], 10, descending=[False], offset=offset)


QUERIES = {i: f for i, f in ((int(n[1:]), f) for n, f in list(globals().items()) if n[0] == "q" and n[1:].isdigit())}
NOT_BUILT: dict = {}
