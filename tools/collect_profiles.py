#!/usr/bin/env python3
"""Copy the evidence of a round from gpurun_out/ into profiles/ (tracked) and print the table for profiles/README.md:
    python tools/collect_profiles.py r02"""
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def lines(path):
    out = []
    try:
        for ln in open(path):
            if ln.startswith("{"):
                out.append(json.loads(ln))
    except OSError:
        pass
    return out


def row(name, d):
    e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    p = d.get("parity_check") or {}
    c = d.get("clocks") or {}
    return (f"| `{name}` | {d['n_gpus']} | {d.get('scaling', '')} | {d['config'].get('frames_per_gpu', '')} x "
            f"{d['config'].get('workload', '').split(' synthetic ')[-1].split(' BGR24')[0]} | {d['value']:,.0f} | "
            f"{r.get('frac', 0):.3f} | {(e.get('value') and format(e['value'], ',.0f')) or '-'} | "
            f"{p.get('bit_equal') if p.get('bit_equal') is not None else p.get('within_1e-4')}/{p.get('shard_boundaries_equal', '-')} | {c.get('sm_mhz')} {c.get('reasons')} |")


def main():
    copies = []
    table = ["| file | GPUs | scaling | frames per GPU x size | frames/s (device-timed) | roofline frac | e2e frames/s | parity / shard boundaries | SM MHz, throttle reasons |",
             "|---|---|---|---|---|---|---|---|---|"]
    for src in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "multi_n*", "*.json*")) +
                      glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_final", "bench*.json")) +
                      glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_final", "sweep*.jsonl"))):
        ds = lines(src)
        if not ds:
            continue
        sub = os.path.basename(os.path.dirname(src))
        dst = f"{sub}_{os.path.basename(src)}" if sub.startswith(tag) else f"{tag}_{sub}_{os.path.basename(src)}"
        with open(os.path.join(ROOT, "profiles", dst), "w") as f:
            for d in ds:
                f.write(json.dumps(d) + "\n")
        copies.append(dst)
        for d in ds:
            table.append(row(dst, d))
    print("\n".join(table))
    print("\ncopied:", len(copies), "files", file=sys.stderr)


if __name__ == "__main__":
    main()
