#!/usr/bin/env python3
"""Transcribe the reference's known-answer regression outputs into JSON fixtures.

Run HERE (dev container, /root/reference present):
    python tests/golden/make_golden.py
It reads /root/reference/test/expected/{vector_type,halfvec,bit,sparsevec}.out (psql
echo of each statement followed by its result), keeps the statements on the
distance hot path and writes tests/golden/distance_kat.json
(and tests/golden/sparsevec_kat.json).
It also transcribes the tiny index-level orderings of
test/expected/{ivfflat_*,hnsw_*}.out into tests/golden/index_orderings.json.
Nothing here is executed on the GPU box; the JSON files are committed.
"""
import json
import os
import re
import sys

REF = os.environ.get("PGV_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

FUNCS = {"l2_distance", "inner_product", "cosine_distance", "l1_distance", "vector_norm", "l2_norm",
         "l2_normalize", "binary_quantize", "hamming_distance", "jaccard_distance"}
OPS = {"<->": "l2_distance", "<#>": "negative_inner_product", "<=>": "cosine_distance", "<+>": "l1_distance",
       "<~>": "hamming_distance", "<%>": "jaccard_distance"}

LIT = r"'([^']*)'(?:::(\w+)(?:\((\d+)\))?)?"


def parse_statement(stmt, default_type):
    """-> dict(fn, args[list of str], types[list]) or None"""
    m = re.fullmatch(r"SELECT (\w+)\((.*)\)(::real)?;", stmt)
    if m and m.group(1) in FUNCS:
        fn, inner = m.group(1), m.group(2)
        lits = re.findall(LIT, inner)
        rest = re.sub(LIT, "", inner).replace(",", "").strip()
        if rest or not lits:
            return None
        return dict(fn=fn, args=[l[0] for l in lits], casts=[(l[1] + (f"({l[2]})" if l[2] else "")) for l in lits],
                    real=bool(m.group(3)))
    m = re.fullmatch(r"SELECT " + LIT + r" (<->|<#>|<=>|<\+>|<~>|<%>) " + LIT + ";", stmt)
    if m:
        g = m.groups()
        return dict(fn=OPS[g[3]], args=[g[0], g[4]], casts=[(g[1] or "") + (f"({g[2]})" if g[2] else ""),
                                                            (g[5] or "") + (f"({g[6]})" if g[6] else "")], real=False)
    return None


def parse_out(path, default_type):
    lines = open(path).read().split("\n")
    cases = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("SELECT ") and ln.endswith(";"):
            st = parse_statement(ln, default_type)
            nxt = lines[i + 1] if i + 1 < len(lines) else ""
            if st is not None:
                if nxt.startswith("ERROR:"):
                    st["error"] = nxt[len("ERROR:"):].strip()
                    st["expected"] = None
                elif i + 3 < len(lines) and set(lines[i + 2].strip()) <= {"-"} and lines[i + 2].strip():
                    st["expected"] = lines[i + 3].strip()
                    st["error"] = None
                else:
                    st = None
                if st is not None:
                    st["type"] = default_type
                    st["source"] = f"test/expected/{os.path.basename(path)}:{i + 1}"
                    cases.append(st)
        i += 1
    return cases


def parse_orderings(path):
    """CREATE TABLE/INSERT/CREATE INDEX/SELECT ... ORDER BY blocks with small literal tables."""
    lines = open(path).read().split("\n")
    blocks, cur = [], None
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"CREATE TABLE t \(val (\w+)\((\d+)\)\);", ln)
        if m:
            cur = dict(type=m.group(1), dim=int(m.group(2)), rows=[], index=None, queries=[], source=f"test/expected/{os.path.basename(path)}:{i + 1}")
        elif cur is not None and ln.startswith("INSERT INTO t (val) VALUES"):
            vals = re.findall(r"\((B?'[^']*'|NULL)\)", ln)
            cur["rows"].append(dict(values=[v.strip("B").strip("'") if v != "NULL" else None for v in vals],
                                    after_index=cur["index"] is not None))
        elif cur is not None and ln.startswith("CREATE INDEX ON t USING"):
            m2 = re.match(r"CREATE INDEX ON t USING (\w+) \(val (\w+)\)(?: WITH \((.*)\))?;", ln)
            if m2 and not (i + 1 < len(lines) and lines[i + 1].startswith("ERROR")):
                cur["index"] = dict(am=m2.group(1), opclass=m2.group(2), options=m2.group(3) or "")
        elif cur is not None and re.match(r"SELECT \* FROM t ORDER BY val (\S+) B?'[^']*';", ln):
            m3 = re.match(r"SELECT \* FROM t ORDER BY val (\S+) B?'([^']*)';", ln)
            res = []
            j = i + 3
            while j < len(lines) and not lines[j].startswith("("):
                res.append(lines[j].strip())
                j += 1
            cur["queries"].append(dict(op=m3.group(1), query=m3.group(2), expected=res))
        elif cur is not None and (ln.startswith("DROP TABLE t") or ln.startswith("TRUNCATE t")):
            if cur["index"] and cur["queries"]:
                blocks.append(cur)
            cur = None
        i += 1
    return blocks


def main():
    exp = os.path.join(REF, "test", "expected")
    if not os.path.isdir(exp):
        sys.exit(f"{exp} not found: run this in the dev container")
    cases = []
    cases += parse_out(os.path.join(exp, "vector_type.out"), "vector")
    cases += parse_out(os.path.join(exp, "halfvec.out"), "halfvec")
    cases += parse_out(os.path.join(exp, "bit.out"), "bit")
    sparse = parse_out(os.path.join(exp, "sparsevec.out"), "sparsevec")
    with open(os.path.join(OUT, "distance_kat.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py", reference="pgvector @ e48241b (v0.8.6+)",
                       cases=cases), f, indent=1)
    # sparsevec goes to its own file: the literals are '{index:value,...}/dim' with 1-based indices
    with open(os.path.join(OUT, "sparsevec_kat.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py", reference="pgvector @ e48241b (v0.8.6+)",
                       cases=sparse), f, indent=1)
    blocks = []
    for name in ("ivfflat_vector", "ivfflat_halfvec", "ivfflat_bit", "hnsw_vector", "hnsw_halfvec", "hnsw_bit"):
        p = os.path.join(exp, name + ".out")
        if os.path.exists(p):
            blocks += parse_orderings(p)
    with open(os.path.join(OUT, "index_orderings.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py", blocks=blocks), f, indent=1)
    sp_blocks = parse_orderings(os.path.join(exp, "hnsw_sparsevec.out"))
    with open(os.path.join(OUT, "sparsevec_orderings.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py", blocks=sp_blocks), f, indent=1)
    print(f"{len(sp_blocks)} sparsevec ordering blocks")
    print(f"{len(cases)} known-answer cases, {len(sparse)} sparsevec cases, {len(blocks)} index ordering blocks")


if __name__ == "__main__":
    main()
