"""bench.py's roofline numbers are algorithmic FLOPs / bytes derived from the ARGUMENTS of each C-ABI call, picked by position
(`bench._work`).  This test ties those positions to the parameter NAMES in include/unetca_b200.h, so a changed prototype
cannot silently mis-account a kernel: every entry gets distinct prime-valued arguments and the result must equal the formula
written with names."""
import os
import re

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DECL = re.compile(r"^(?:const char\*|int|void)\s+(unetca_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)
_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101]


def _params():
    text = open(os.path.join(ROOT, "include", "unetca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for name, args in _DECL.findall(text):
        names = []
        for a in args.split(","):
            a = a.strip()
            if a and a != "void":
                names.append(re.findall(r"(\w+)\s*$", a)[0])
        out[name] = names
    return out


def _call(name, params, e=2, cin=3):
    names = params[name]
    assert len(names) <= len(_PRIMES)
    v = dict(zip(names, _PRIMES))
    args = [v[n] for n in names]
    return v, bench._work(name, args, e, cin)


CONV = ["unetca_conv3x3_fwd", "unetca_conv3x3_wgrad", "unetca_conv3x3_fwd_paired", "unetca_conv3x3_fwd_split",
        "unetca_conv3x3_dgrad_bnstats", "unetca_conv3x3_fwd_cat", "unetca_conv3x3_wgrad_cat"]


def test_conv_entries_use_the_named_shape_arguments():
    params = _params()
    for name in CONV:
        v, (cls, fl, by) = _call(name, params)
        assert cls == "tensor" and by == 0, name
        assert fl == 2.0 * v["B"] * v["H"] * v["W"] * 9 * v["C"] * v["O"], (name, params[name])


def test_fixed_width_conv_entries():
    params = _params()
    v, (cls, fl, _) = _call("unetca_conv3x3_fwd_kw", params)
    assert cls == "tensor" and fl == 2.0 * v["B"] * v["H"] * v["W"] * 9 * v["C"] * 64
    v, (cls, fl, _) = _call("unetca_conv3x3_fwd_rp64", params)
    assert cls == "tensor" and fl == 2.0 * v["B"] * v["H"] * v["W"] * 9 * 64 * 64


def test_conv_transpose_and_first_conv_entries():
    params = _params()
    for name in ("unetca_convT2x2_fwd", "unetca_convT2x2_dgrad", "unetca_convT2x2_wgrad"):
        v, (cls, fl, _) = _call(name, params)
        hh, ww = "h", "wd"                                     # input extents (w is the filter)
        assert cls == "tensor" and fl == 2.0 * v["B"] * v[hh] * v[ww] * v["Cin"] * 4 * v["Cout"], (name, params[name])
    v, (cls, fl, _) = _call("unetca_first_pairs_fwd", params, cin=3)
    assert cls == "tensor" and fl == 2.0 * v["B"] * v["H"] * v["W"] * v["O"] * 9 * 3
    v, (cls, fl, _) = _call("unetca_first_pairs_wgrad", params)
    assert cls == "tensor" and fl == 2.0 * v["B"] * v["H"] * v["W"] * 9 * v["Cin"] * v["O"]


def test_hbm_entries_use_the_named_shape_arguments():
    params = _params()
    e = 2
    v, (cls, _, by) = _call("unetca_bn_bwd_apply", params, e)
    assert cls == "hbm" and by == 3 * v["B"] * v["pix_per_img"] * v["C"] * e
    v, (cls, _, by) = _call("unetca_se_squeeze", params, e)
    assert cls == "hbm" and by == v["B"] * v["pix_per_img"] * v["C"] * e
    v, (cls, _, by) = _call("unetca_bn_bwd_apply_pool", params, e)
    n = v["B"] * v["H"] * v["W"] * v["C"]
    assert cls == "hbm" and by == 3 * n * e + (n // 4) * (e + 1)
    v, (cls, _, by) = _call("unetca_se_bn_bwd_reduce_pool", params, e)
    n = v["B"] * v["H"] * v["W"] * v["C"]
    assert cls == "hbm" and by == 2 * n * e + (n // 4) * (e + 1)
