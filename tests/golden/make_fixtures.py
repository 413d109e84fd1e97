"""Extracts the reference's own DATA fixtures (not code) into small .npy files.

Run once in the authoring container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
Sources:
  * 64 x 24 "protein" table  -- examples/01_compare_cosine.rs:10-75 (same table in
    examples/02_proteins_lookup.rs:30-95 and benches/index_compute_bench.rs:21-86)
  * QUORA_EMBEDDS 15 x 384   -- src/tests/test_data.rs:6-5797
"""
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def proteins():
    txt = (REF / "examples/01_compare_cosine.rs").read_text()
    body = txt.split('r#"', 1)[1].split('"#', 1)[0]
    rows = []
    for line in body.strip().splitlines():
        _, vals = line.split(";")
        rows.append([float(v) for v in vals.split(",")])
    a = np.array(rows, dtype=np.float64)
    assert a.shape == (64, 24), a.shape
    np.save(OUT / "proteins_64x24.npy", a)


def quora():
    txt = (REF / "src/tests/test_data.rs").read_text()
    start = txt.index("QUORA_EMBEDDS")
    end = txt.index("];", start)
    seg = txt[txt.index("=", start):end]
    rows = re.findall(r"\[([^\[\]]+)\]", seg)
    a = np.array([[float(v) for v in r.replace("\n", " ").split(",") if v.strip()] for r in rows], dtype=np.float64)
    assert a.shape == (15, 384), a.shape
    np.save(OUT / "quora_15x384.npy", a)


if __name__ == "__main__":
    proteins()
    quora()
    print("fixtures written to", OUT)
