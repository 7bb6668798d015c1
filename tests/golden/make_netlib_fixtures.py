"""Regenerates tests/golden/netlib_*.json from the reference's benchmark inputs.

Run in the build container only (needs /root/reference):
    python tests/golden/make_netlib_fixtures.py
The three files under /root/reference/tests/benchmark_problems/ are the netlib LPs AFIRO, ADLITTLE
and BLEND in the one-pair-per-line layout the reference's reader accepts (src/parse_mps.rs:290-295).
The fixture keeps FILE ORDER (rows as listed in ROWS, columns in order of first appearance), which
is the deterministic order SURVEY.md 8(d) prescribes (the reference itself iterates HashMaps,
src/parse_mps.rs:29,41).  Expected objectives: tests/problems/mod.rs:661,667,673.
"""
import json
import os

REF = "/root/reference/tests/benchmark_problems"
OUT = os.path.dirname(os.path.abspath(__file__))
EXPECTED = {"afiro": -4.6475314286E+02, "adlittle": 2.2549496316E+05, "blend": -3.0812149846E+01}


def parse(path):
    section = None
    rows, row_index, cols, col_index = [], {}, [], {}
    obj_row = None
    for raw in open(path):
        if not raw.strip():
            continue
        tok = raw.split()
        if not raw[0].isspace():
            section = tok[0]
            continue
        if section == "ROWS":
            kind, name = tok
            if kind == "N":
                obj_row = name
            else:
                row_index[name] = len(rows)
                rows.append({"name": name, "op": {"L": 0, "E": 1, "G": 2}[kind], "rhs": 0.0})
        elif section == "COLUMNS":
            cname, rname, val = tok
            if cname not in col_index:
                col_index[cname] = len(cols)
                cols.append({"name": cname, "obj": 0.0, "rows": [], "vals": []})
            col = cols[col_index[cname]]
            if rname == obj_row:
                col["obj"] = float(val)
            else:
                col["rows"].append(row_index[rname])
                col["vals"].append(float(val))
        elif section == "RHS":
            rname, val = tok[-2], tok[-1]
            rows[row_index[rname]]["rhs"] = float(val)
        elif section in ("BOUNDS", "RANGES"):
            raise SystemExit("unexpected section " + section)
    return rows, cols


for name, expected in EXPECTED.items():
    rows, cols = parse(os.path.join(REF, name, name + ".mps"))
    with open(os.path.join(OUT, f"netlib_{name}.json"), "w") as f:
        json.dump({"name": name, "expected_obj": expected, "source": f"tests/benchmark_problems/{name}/{name}.mps",
                   "rows": rows, "cols": cols}, f, separators=(",", ":"))
    print(name, len(rows), "rows", len(cols), "cols", sum(len(c["rows"]) for c in cols), "nnz")
