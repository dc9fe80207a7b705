"""Turns `ncu --page raw --csv` exports of the two hot kernels into profiles/traffic.json, the record bench.py reads for
`roofline.traffic`.  Each entry is stamped with the revision (sha1) of the kernel sources it was captured at; bench.py
refuses a record whose revision is not the current one, so the number cannot silently go stale.

usage: python tools/traffic_from_ncu.py <accum.csv> <ntt.csv> [--vectors 256] [--log-n 17]
  accum.csv : ncu -i accum.ncu-rep --page raw --csv   (k_accum_entries launches of a bench.py step)
  ntt.csv   : the same for k_ntt_pass"""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_revision, source lists)


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units, data = r[0], r[1], r[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for d in data:
        def val(name):
            v = float(d[idx[name]].replace(",", ""))
            u = units[idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        out.append({"kernel": d[idx["Kernel Name"]], "grid": d[idx["launch__grid_size"]], "ms": float(d[idx["gpu__time_duration.sum"]].replace(",", "")) *
                    {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "second": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1),
                    "rd": val("dram__bytes_read.sum"), "wr": val("dram__bytes_write.sum")})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("accum")
    ap.add_argument("ntt")
    ap.add_argument("--vectors", type=int, default=256, help="scalar vectors per captured k_accum_entries launch")
    ap.add_argument("--log-n", type=int, default=17)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    acc = [r for r in rows(a.accum) if "k_accum_entries" in r["kernel"]]
    ntt = [r for r in rows(a.ntt) if "k_ntt_pass" in r["kernel"]]
    big = max(acc, key=lambda r: r["ms"])       # the launch over the dense columns
    rec = {
        "k_accum_entries": {
            "kernel_revision": bench.kernel_revision(bench.MSM_SOURCES), "dram_bytes_per_launch": big["rd"] + big["wr"],
            "dram_read": big["rd"], "dram_write": big["wr"], "launch_ms_under_ncu": big["ms"], "grid": big["grid"],
            "algorithmic_bytes_per_launch": a.vectors * (1 << a.log_n) * 96,
            "source": f"ncu --set full, the k_accum_entries launch over {a.vectors} x 2^{a.log_n} full-size scalars of one bench.py step. {a.note}".strip()},
        "k_ntt_pass": {
            "kernel_revision": bench.kernel_revision(bench.NTT_SOURCES),
            "dram_bytes_per_launch": sum(r["rd"] + r["wr"] for r in ntt) / max(1, len(ntt)),
            "launches_captured": len(ntt), "dram_read_total": sum(r["rd"] for r in ntt), "dram_write_total": sum(r["wr"] for r in ntt),
            "ms_total_under_ncu": sum(r["ms"] for r in ntt),
            "algorithmic_bytes_per_launch": None,
            "source": f"ncu --set full over {len(ntt)} consecutive k_ntt_pass launches of one bench.py step (mean per launch). {a.note}".strip()},
    }
    json.dump(rec, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
