"""Reads the dump of scripts/trace_desc.py and prints, per kernel and role, where the cycles go (median CTA)."""
import sys

import numpy as np

d = np.load(sys.argv[1])


def events(tr, cta, role):
    v = tr[cta, role]
    v = v[v != 0]
    return (v >> 8), (v & 0xFF)


def spans(t, tag, a, b):
    """sum over consecutive (a -> b) tag pairs of the elapsed cycles"""
    tot, n = 0, 0
    ia = np.where(tag == a)[0]
    for i in ia:
        j = i + 1
        while j < len(tag) and tag[j] != b:
            j += 1
        if j < len(tag):
            tot += t[j] - t[i]
            n += 1
    return tot, n


for name in ("fwd", "bwd"):
    tr = d[name]
    print("== %s: traced launch %.1f us" % (name, float(d[name + "_us"])))
    rows = []
    for cta in range(tr.shape[0]):
        t0, tag0 = events(tr, cta, 0)
        if len(t0) == 0:
            continue
        total = t0[-1] - t0[0]
        r = {"cta": cta, "mma_total": total}
        if name == "bwd":
            r["mma_wait_dempty"] = spans(t0, tag0, 2, 3)[0]
            r["mma_wait_b"] = spans(t0, tag0, 4, 5)[0]
            r["mma_wait_a"] = spans(t0, tag0, 5, 7)[0]
            r["mma_issue"] = spans(t0, tag0, 7, 6)[0]
            r["n_stages"] = spans(t0, tag0, 4, 5)[1]
        else:
            r["mma_wait_afull"] = spans(t0, tag0, 1, 8)[0]
            r["mma_wait_tempty"] = spans(t0, tag0, 2, 3)[0]
            r["mma_wait_b"] = spans(t0, tag0, 4, 5)[0]
            r["mma_issue"] = spans(t0, tag0, 5, 6)[0]
            r["n_stages"] = spans(t0, tag0, 4, 5)[1]
        t1, tag1 = events(tr, cta, 1)
        r["prod_wait_free"] = spans(t1, tag1, 12, 13)[0]
        if name == "fwd":
            r["prod_wait_aempty"] = spans(t1, tag1, 10, 11)[0]
        t2, tag2 = events(tr, cta, 2)
        if name == "bwd":
            r["exp_wait_free"] = spans(t2, tag2, 20, 21)[0]
            r["exp_lut"] = spans(t2, tag2, 21, 22)[0]
            r["exp_st"] = spans(t2, tag2, 22, 23)[0]
            r["exp_between"] = spans(t2, tag2, 23, 20)[0]
            t3, tag3 = events(tr, cta, 3)
            r["epi_wait_dfull"] = spans(t3, tag3, 30, 31)[0]
            r["epi_work"] = spans(t3, tag3, 31, 32)[0]
            r["epi_items"] = spans(t3, tag3, 31, 32)[1]
        else:
            for role, nm in ((2, "epi0"), (3, "epi7")):
                t3, tag3 = events(tr, cta, role)
                r[nm + "_wait_tfull"] = spans(t3, tag3, 30, 31)[0]
                r[nm + "_work"] = spans(t3, tag3, 31, 32)[0]
                r[nm + "_items"] = spans(t3, tag3, 31, 32)[1]
        rows.append(r)
    keys = [k for k in rows[0] if k != "cta"]
    print("   CTAs traced: %d" % len(rows))
    for k in keys:
        v = np.array([r[k] for r in rows], dtype=np.float64)
        print("   %-18s median %10.0f   min %10.0f   max %10.0f" % (k, np.median(v), v.min(), v.max()))
    # per-stage detail of the median CTA: issue-to-issue interval histogram of the MMA thread
    cta = rows[len(rows) // 2]["cta"]
    t0, tag0 = events(tr, cta, 0)
    iss = t0[tag0 == 6]
    if len(iss) > 2:
        dd = np.diff(iss)
        print("   CTA %d: commit-to-commit interval of the MMA thread: median %d, p10 %d, p90 %d, max %d cycles (%d stages)"
              % (cta, np.median(dd), np.percentile(dd, 10), np.percentile(dd, 90), dd.max(), len(dd)))
