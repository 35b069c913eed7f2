"""Summarise an ncu launch list of one bench step into profiles/ (per-kernel time share + DRAM bytes).

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --launch-skip 1400 --csv --log-file gpurun_out/ncu_step.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
  python tools/ncu_step_summary.py gpurun_out/ncu_step.csv profiles/r1u_ncu_step_kernels [--traffic profiles/ncu_traffic.json]

The last step is everything from the last launch of the step's first kernel (im2col_pairs_kernel: the input transform of
the first convolution) to the end of the list.  Times under ncu are cold-cache and serialised: only the SHARES are
comparable with bench.py's live CUDA-event numbers.
"""
import argparse
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_hash():
    """Identity of the kernel sources the capture was taken on (bench.py compares it with the build it is timing)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "insar-unet-ca_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"unetca::\(anonymous namespace\)::|unetca::<unnamed>::|unetca::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.strip()


def klass(name):
    if name.startswith("tc_"):
        return "tensor"
    if name.startswith(("at::", "void at::", "adam_", "pack_", "wgrad_reduce", "sum_parts", "first_pairs_fold", "img_parts_sum")):
        return "other"
    if "finalize" in name or name.startswith(("se_fc", "confusion")):
        return "other"              # tiny per-channel kernels
    return "hbm"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("out_prefix")
    ap.add_argument("--first-kernel", default="im2col_pairs_kernel")
    ap.add_argument("--traffic", default=None, help="also rewrite this ncu_traffic.json (read by bench.py)")
    ap.add_argument("--command", default="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e")
    a = ap.parse_args()
    rows = {}
    order = []
    with open(a.csv, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    col = {n: i for i, n in enumerate(header)}
    for r in rd:
        if len(r) < len(header):
            continue
        kid = int(r[col["ID"]])
        if kid not in rows:
            rows[kid] = {"name": short(r[col["Kernel Name"]])}
            order.append(kid)
        val = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            val *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
            rows[kid]["ms"] = val
        else:
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
            rows[kid]["rd" if "read" in m else "wr"] = val
    starts = [k for k in order if rows[k]["name"].startswith(a.first_kernel)]
    if not starts:
        sys.exit(f"no launch of {a.first_kernel} in {a.csv}")
    step = [rows[k] for k in order if k >= starts[-1]]
    agg = {}
    for r in step:
        d = agg.setdefault(r["name"], {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        d["launches"] += 1
        d["ms"] += r.get("ms", 0.0)
        d["dram_read_bytes"] += r.get("rd", 0.0)
        d["dram_write_bytes"] += r.get("wr", 0.0)
    total = sum(d["ms"] for d in agg.values())
    for n, d in agg.items():
        d["share"] = d["ms"] / total
        d["class"] = klass(n)
    ordered = dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))
    json.dump(ordered, open(a.out_prefix + ".json", "w"), indent=1)
    with open(a.out_prefix + ".txt", "w") as f:
        f.write(f"ncu launch list of `{a.command}` (last step), per kernel; cold-cache, serialised replays\n")
        f.write(f"total {total:.3f} ms under ncu, {len(step)} launches\n")
        for n, d in ordered.items():
            f.write(f"{n[:80]:80s} n={d['launches']:4d} {d['ms']:8.3f} ms {100 * d['share']:5.1f}%  dram rd {d['dram_read_bytes'] / 1e9:7.2f} GB "
                    f"wr {d['dram_write_bytes'] / 1e9:7.2f} GB  [{d['class']}]\n")
        for c in ("tensor", "hbm", "other"):
            ms = sum(d["ms"] for d in agg.values() if d["class"] == c)
            f.write(f"class {c:6s}: {ms:8.3f} ms {100 * ms / total:5.1f}%\n")
    if a.traffic:
        def cls(c):
            ds = [d for d in agg.values() if d["class"] == c]
            n = sum(d["launches"] for d in ds)
            return {"launches_per_step": n, "dram_bytes_per_launch": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in ds) / max(n, 1),
                    "ms_per_step_under_ncu": sum(d["ms"] for d in ds)}

        def one(pred, label):
            ds = [d for n, d in agg.items() if pred(n)]
            n = sum(d["launches"] for d in ds)
            return {"kernel": label, "launches_per_step": n,
                    "dram_bytes_per_launch": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in ds) / max(n, 1),
                    "ms_per_step_under_ncu": sum(d["ms"] for d in ds)}
        out = {"csrc_sha": csrc_hash(),
               "source": f"{a.out_prefix}.json (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         f"--clock-control none {a.command}, last step)",
               "tensor": cls("tensor"), "hbm": cls("hbm"),
               "dominant": one(lambda n: n.startswith("tc_conv3x3_hpix_kernel"), "tc_conv3x3_hpix_kernel"),
               "dominant_hbm": one(lambda n: n.startswith("bn_bwd_apply_stream_kernel") or re.match(r"bn_bwd_kernel<.*,\s*1>$", n) is not None,
                                   "bn_bwd_apply_stream_kernel / bn_bwd_kernel<bf16,*,APPLY> (unetca_bn_bwd_apply)")}
        json.dump(out, open(a.traffic, "w"), indent=1)
    print(open(a.out_prefix + ".txt").read())


if __name__ == "__main__":
    main()
