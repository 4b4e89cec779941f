#!/usr/bin/env python3
"""Extract the reference's known-answer constants into tests/golden/reference_kats.json.

Run in the authoring container only (it reads /root/reference, which does not exist on the
GPU box):

    python tests/golden/extract_kats.py

It parses every `static` / `const` item with an integer-array body from the reference's hot-path
sources and the literal test vectors that live inside test bodies (division KAT, Ristretto hex
encodings, Elligator KAT, scalar bit / NAF strings).  Only numbers are extracted -- no code.

Sources (reference file:line ranges are recorded per item in the JSON):
  src/backend/u64/constants.rs   moduli, Montgomery constants, curve constants, basepoint, table
  src/backend/u64/field.rs       field KATs   (tests module, :930-1556)
  src/backend/u64/scalar.rs      scalar KATs  (tests module, :677-1053)
  src/edwards.rs                 point KATs   (tests module, :1136-1636)
  src/ristretto.rs               Ristretto encodings of [0..15]B, Elligator KAT (:526-721)
"""
import json
import os
import re
import sys

REF = os.environ.get("ZEROCAF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

ITEM_RE = re.compile(
    r"(?:pub(?:\([a-z]+\))?\s+)?(?:static|const)\s+([A-Za-z0-9_]+)\s*:\s*([^=]+?)\s*=\s*(.*?);\s*$",
    re.S | re.M,
)
INT_RE = re.compile(r"(?<![A-Za-z_0-9])(\d[\d_]*)(?:u8|u64|u128|i8)?(?![A-Za-z_0-9\.])")


def strip_comments(text):
    text = re.sub(r"//[^\n]*", "", text)
    return text


def ints_of(body):
    # drop the array-length part of a type such as `[u64; 5]` if it sneaks in
    return [int(x.replace("_", "")) for x in INT_RE.findall(body)]


def items_of(path):
    with open(path) as f:
        raw = f.read()
    text = strip_comments(raw)
    out = {}
    # walk item by item: find "static|const NAME: TYPE =" then balance brackets up to ';'
    for m in re.finditer(
        r"(?:static|const)\s+([A-Za-z0-9_]+)\s*:\s*((?:\[[^\]]*\]|[^=;\[])+?)\s*=", text
    ):
        name, ty = m.group(1), " ".join(m.group(2).split())
        if ty.startswith("fn") or "(" in ty and "[" not in ty and "Point" not in ty:
            continue
        i = m.end()
        depth = 0
        j = i
        while j < len(text):
            ch = text[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == ";" and depth == 0:
                break
            j += 1
        body = text[i:j]
        # identifiers like u64 / FieldElement contain digits we must not read as values
        body_clean = re.sub(r"[A-Za-z_][A-Za-z_0-9]*", " ", body)
        vals = [int(x) for x in re.findall(r"\d+", body_clean)]
        if not vals:
            continue
        line = raw.count("\n", 0, raw.find(name + ":")) + 1
        out[name] = {"type": ty, "values": vals, "line": line}
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s" % REF)
    kats = {}
    for key, rel in [
        ("constants", "src/backend/u64/constants.rs"),
        ("field", "src/backend/u64/field.rs"),
        ("scalar", "src/backend/u64/scalar.rs"),
        ("edwards", "src/edwards.rs"),
    ]:
        kats[key] = {"source": rel, "items": items_of(os.path.join(REF, rel))}

    # ---- literals that live inside test bodies -------------------------------------------
    with open(os.path.join(REF, "src/backend/u64/field.rs")) as f:
        field_src = f.read()
    m = re.search(r"fn division\(\).*?let expected = FieldElement\(\[(.*?)\]\)", field_src, re.S)
    kats["field"]["inline"] = {
        "division": {
            "a": 86649,
            "b": 86650,
            "neg_a_over_b": [int(x) for x in re.findall(r"\d+", m.group(1))],
            "source": "src/backend/u64/field.rs:1242-1260",
        }
    }
    m = re.search(r"from_canonical_bytes\(\[(.*?)\]\)", field_src, re.S)
    dalek_bytes = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1))]
    m = re.search(r"fn from_ristretto255scalar.*?let res = FieldElement\(\[(.*?)\]\)", field_src, re.S)
    kats["field"]["inline"]["from_bytes_vector"] = {
        "bytes": dalek_bytes,
        "limbs": [int(x) for x in re.findall(r"\d+", m.group(1))],
        "source": "src/backend/u64/field.rs:1378-1399",
    }

    with open(os.path.join(REF, "src/backend/u64/scalar.rs")) as f:
        scalar_src = f.read()
    inline = {}
    m = re.search(r"let minus_one = \[(.*?)\];", scalar_src, re.S)
    inline["into_bits_minus_one"] = [int(x) for x in re.findall(r"\d+", m.group(1))]
    for w in (2, 3, 4, 5, 6):
        m = re.search(r"let naf%d_scalar = \[(.*?)\];" % w, scalar_src, re.S)
        inline["wnaf%d_1122334455" % w] = [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    inline["source"] = "src/backend/u64/scalar.rs:979-1052"
    kats["scalar"]["inline"] = inline

    with open(os.path.join(REF, "src/edwards.rs")) as f:
        edw_src = f.read()
    inline = {}
    m = re.search(r"fn point_compression\(\).*?from_slice\(&\[(.*?)\]\).*?from_slice\(&\[(.*?)\]\)", edw_src, re.S)
    inline["P1_compress"] = [int(x) for x in re.findall(r"\d+", m.group(1))]
    inline["P2_compress"] = [int(x) for x in re.findall(r"\d+", m.group(2))]
    inline["source"] = "src/edwards.rs:1549-1562"
    kats["edwards"]["inline"] = inline

    with open(os.path.join(REF, "src/ristretto.rs")) as f:
        ris_src = f.read()
    hexes = re.findall(r'"([0-9a-f]{64})"', ris_src)
    m = re.search(r"fn elligator_vs_ristretto_sage.*?EdwardsPoint \{(.*?)\}\);", ris_src, re.S)
    body_clean = re.sub(r"[A-Za-z_][A-Za-z_0-9]*", " ", strip_comments(m.group(1)))
    m2 = re.search(r"fn validity_check.*?from_bytes\(&\[(.*?)\]\)", ris_src, re.S)
    kats["ristretto"] = {
        "source": "src/ristretto.rs",
        "small_multiples_hex": hexes[:16],  # compress([k]B), k = 0..15   (:546-566)
        "elligator_input_hex": hexes[16],  # :705
        "elligator_expected_point": [int(x) for x in re.findall(r"\d+", body_clean)],  # :682-702
        "order_8L_point_y_bytes": [int(x) for x in re.findall(r"\d+", m2.group(1))],  # :654-657
    }
    with open(os.path.join(REF, "src/constants.rs")) as f:
        csrc = f.read()
    kats["ristretto"]["compressed_basepoints"] = {
        k: v["values"] for k, v in items_of(os.path.join(REF, "src/constants.rs")).items()
    }

    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1, sort_keys=True)
    n = sum(len(v.get("items", {})) for v in kats.values())
    print("wrote %s: %d named items" % (OUT, n))


if __name__ == "__main__":
    main()
