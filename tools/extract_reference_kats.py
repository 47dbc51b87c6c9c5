#!/usr/bin/env python3
"""Extract the reference's known-answer constants into tests/golden/ (run in the build container).

/root/reference does not exist on the GPU box, so everything the tests need from it is frozen here:
  * tests/golden/reference_vectors/*.zz              <- reference tests/*.zz (binary test DATA, verbatim)
  * tests/golden/reference_vectors/fuzz_corpus_inflate/* <- reference fuzz/corpus/inflate/* (verbatim)
  * tests/golden/reference_tables.json               <- numeric constants of src/tables.rs that the
        reference's own unit tests use as expected values (FIXED_LITLEN_TABLE, FIXED_DIST_TABLE,
        decompress.rs:1218-1233) or that define the ultra-fast format (HUFFMAN_LENGTHS, LENGTH_TO_*),
        plus the 54-byte HEADER of src/compress/ultrafast.rs:82-86.
No reference source code is copied; only data constants and test vectors.
"""
import json
import re
import shutil
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def arr(src: str, name: str):
    m = re.search(r"const " + name + r"[^=]*=\s*\[(.*?)\];", src, re.S)
    return [int(x) for x in re.findall(r"\d+", m.group(1))]


def main():
    tables = (REF / "src/tables.rs").read_text()
    uf = (REF / "src/compress/ultrafast.rs").read_text()
    doc = {
        "source": "image-rs/fdeflate 0.4.0-dev src/tables.rs, src/compress/ultrafast.rs:82-86",
        "HUFFMAN_LENGTHS": arr(tables, "HUFFMAN_LENGTHS"),
        "LENGTH_TO_SYMBOL": arr(tables, "LENGTH_TO_SYMBOL"),
        "LENGTH_TO_LEN_EXTRA": arr(tables, "LENGTH_TO_LEN_EXTRA"),
        "FIXED_LITLEN_TABLE": arr(tables, "FIXED_LITLEN_TABLE"),
        "FIXED_DIST_TABLE": arr(tables, "FIXED_DIST_TABLE"),
        "ULTRAFAST_HEADER": arr(uf, "HEADER"),
    }
    assert len(doc["HUFFMAN_LENGTHS"]) == 286 and len(doc["FIXED_LITLEN_TABLE"]) == 512
    assert len(doc["FIXED_DIST_TABLE"]) == 32 and len(doc["ULTRAFAST_HEADER"]) == 54
    (OUT / "reference_tables.json").write_text(json.dumps(doc, separators=(",", ":")) + "\n")
    rv = OUT / "reference_vectors"
    (rv / "fuzz_corpus_inflate").mkdir(parents=True, exist_ok=True)
    for f in (REF / "tests").glob("*.zz"):
        shutil.copy(f, rv / f.name)
    for f in (REF / "fuzz/corpus/inflate").iterdir():
        shutil.copy(f, rv / "fuzz_corpus_inflate" / f.name)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
