#!/usr/bin/env python
"""Restate the known-answer vectors of the reference's experimental/seal tests as data.

Run in the authoring container (reads /root/reference/test/experimental/seal/*.cpp,
nothing at test time):  python tests/golden/extract_seal_kats.py
Writes tests/golden/seal_kats.json.  Only numeric initialiser lists and the scalar
parameters next to them are taken; arithmetic expressions in the DyadicMultiply
expectations (e.g. "(1 * 8 + 4 * 2) % 10") are evaluated.
"""
import json
import os
import re

REF = os.environ.get("HEXL_REFERENCE_TREE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def strip_comments(t):
    return re.sub(r"//[^\n]*", "", t)


def split_top_level(body):
    """split a brace-initialiser body at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in body:
        if ch in "({":
            depth += 1
        elif ch in ")}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [c.strip() for c in out if c.strip()]


def eval_int(expr, env=None):
    """integer arithmetic only; `op1[k]` may refer to an already-parsed vector"""
    assert re.fullmatch(r"[0-9+\-*%() \n]|(?:[0-9+\-*%() \n]|op1\[\d+\])+", expr), expr
    return int(eval(expr, {"__builtins__": {}}, env or {}))


def vec(body, env=None):
    return [eval_int(x, env) for x in split_top_level(body)]


def find_vec(text, name, env=None):
    m = re.search(r"std::vector<uint64_t>\s+" + name + r"\s*\{(.*?)\};", text, flags=re.S)
    return vec(m.group(1), env)


def scalar(text, name):
    return int(re.search(r"size_t\s+" + name + r"\s*=\s*(\d+)", text).group(1))


def tests(text):
    parts = re.split(r"\nTEST\((\w+),\s*(\w+)\)\s*\{", text)
    return {parts[i + 1]: parts[i + 2] for i in range(1, len(parts) - 2, 3)}


out = {"_comment": "Restated from the reference's tests by tests/golden/extract_seal_kats.py"}

# ---- KeySwitch: test/experimental/seal/test-key-switch.cpp:16-186
t = strip_comments(open(os.path.join(REF, "test/experimental/seal/test-key-switch.cpp")).read())
body = tests(t)["small"]
m = re.search(r"key_vector\s*\{(.*?)\}\s*;\s*\n\s*size_t coeff_count", body, flags=re.S)
keys = [vec(k.strip()[1:-1]) for k in split_top_level(m.group(1))]
out["key_switch"] = {
    "_source": "test/experimental/seal/test-key-switch.cpp:16-186 (KeySwitch, small)",
    "coeff_count": scalar(body, "coeff_count"),
    "decomp_modulus_size": scalar(body, "decomp_modulus_size"),
    "key_modulus_size": scalar(body, "key_modulus_size"),
    "rns_modulus_size": scalar(body, "rns_modulus_size"),
    "key_component_count": scalar(body, "key_component_count"),
    "moduli": find_vec(body, "moduli"),
    "modswitch_factors": find_vec(body, "modswitch_factors"),
    "k_switch_keys": keys,
    "input": find_vec(body, "input"),
    "t_target_iter_ptr": find_vec(body, "t_target_iter_ptr"),
    "expected_output": find_vec(body, "expected_output"),
}

# ---- DyadicMultiply: test/experimental/seal/test-dyadic-multiply.cpp:16-155
t = strip_comments(open(os.path.join(REF, "test/experimental/seal/test-dyadic-multiply.cpp")).read())
cases = []
for name, body in tests(t).items():
    op1 = find_vec(body, "op1")
    c = {"name": name, "coeff_count": scalar(body, "coeff_count"), "moduli": find_vec(body, "moduli"),
         "op1": op1, "exp_out": find_vec(body, "exp_out", {"op1": op1})}
    c["op2"] = find_vec(body, "op2") if re.search(r"std::vector<uint64_t>\s+op2\s*\{", body) else None
    call = re.search(r"DyadicMultiply\((\w+)\.data\(\),\s*(\w+)\.data\(\),\s*(\w+)\.data\(\)", body)
    c["call"] = {"result": call.group(1), "operand1": call.group(2), "operand2": call.group(3)}
    cases.append(c)
out["dyadic_multiply"] = {"_source": "test/experimental/seal/test-dyadic-multiply.cpp:16-155", "cases": cases}

json.dump(out, open(os.path.join(HERE, "seal_kats.json"), "w"), indent=1)
print({k: (len(v["cases"]) if "cases" in v else "1 case") for k, v in out.items() if k != "_comment"})
