"""Extracts the reference's known-answer vectors for hot path (a) into tests/golden/*.json.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_kats.py
Sources:
  * RSA-2048 PKCS#1 v1.5 signature KATs: /root/reference/src/chip.rs:683-803
    (test_rsa_signature_circuit1/2 must verify, test_bad_rsa_signature_circuit2 must fail)
  * unreduced-product KATs for BigIntChip::mul / square: /root/reference/src/big_integer/chip.rs:2797-3100
    (test_mul_case1,3,4,5,6,7).  The Rust BigUint expressions `let a_big = ...;` are rewritten to
    Python integer expressions and evaluated, so the vectors are the reference's own literals.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def rsa_kats():
    src = open(os.path.join(REF, "src/chip.rs")).read()
    out = []
    for name, should_err in (("test_rsa_signature_circuit1", False), ("test_rsa_signature_circuit2", False),
                             ("test_bad_rsa_signature_circuit2", True)):
        i = src.find(name + ",")
        if i < 0:
            continue
        blk = src[i:i + 12000]
        blk = blk[:blk.index("impl_rsa_signature_test_circuit!(") if "impl_rsa_signature_test_circuit!(" in blk[10:] else len(blk)]
        nums = re.findall(r'BigUint::from_str\("(\d+)"\)', blk)
        if len(nums) < 3:
            continue
        line = src[:i].count("\n") + 1
        out.append({"name": name, "ref": f"src/chip.rs:{line}", "should_be_error": should_err,
                    "n": nums[0], "sig": nums[1], "hash": nums[2]})
    return out


def _rust_biguint_expr(expr: str) -> int:
    e = expr
    e = re.sub(r'BigUint::from_str\("(\d+)"\)\s*\.unwrap\(\)', r"\1", e)
    e = re.sub(r"BigUint::from\((\d+)(?:usize|u128|u64)\)", r"\1", e)
    e = re.sub(r"&?out_base\.pow\((\d+)(?:u32)?\)", r"(B**\1)", e)
    e = re.sub(r"&out_base", "B", e)
    e = re.sub(r"(\d+)(?:usize|u128|u64)", r"\1", e)
    e = e.replace("zero_big.clone()", "0").replace("zero_big", "0")
    assert re.fullmatch(r"[\d\s+*()B]+", e), e
    return eval(" ".join(e.split()), {"B": 1 << 64})


def mul_kats():
    path = "src/big_integer/chip.rs"
    src = open(os.path.join(REF, path)).read()
    out = []
    for name in ("test_mul_case1", "test_mul_case3", "test_mul_case4", "test_mul_case5", "test_mul_case6", "test_mul_case7"):
        i = src.find(name + ",")
        assert i >= 0, name
        j = src.find("impl_bigint_test_circuit!(", i)
        blk = src[i:j if j > 0 else len(src)]
        line = src[:i].count("\n") + 1
        desc = re.search(r'\|\|\s*"([^"]*)"', blk).group(1).strip()
        vals = {}
        for var in ("a_big", "b_big", "ans_big"):
            m = re.search(r"let %s\s*=\s*(.*?);" % var, blk, re.S)
            if m:
                vals[var] = _rust_biguint_expr(m.group(1))
        if name == "test_mul_case1":       # one * one == one.to_muled(): written with assign_constant_fresh(1)
            vals = {"a_big": 1, "b_big": 1, "ans_big": 1}
        if "b_big" not in vals:            # square(a)
            vals["b_big"] = vals["a_big"]
        out.append({"name": name, "ref": f"{path}:{line}", "desc": desc, "bits_len": 2048,
                    "a": str(vals["a_big"]), "b": str(vals["b_big"]), "ans": str(vals["ans_big"])})
    return out


if __name__ == "__main__":
    k = rsa_kats()
    json.dump(k, open(os.path.join(HERE, "rsa_kats.json"), "w"), indent=1)
    print("wrote", len(k), "rsa kats:", [x["name"] for x in k])
    m = mul_kats()
    json.dump(m, open(os.path.join(HERE, "bigint_mul_kats.json"), "w"), indent=1)
    print("wrote", len(m), "mul kats:", [x["name"] for x in m])
