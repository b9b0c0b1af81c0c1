"""Summarises the per-launch device timelines written by `bench.py --trace PREFIX` (one JSON per rank).

    python tools/timeline_summary.py PREFIX [> profiles/xyz_timeline.txt]

For the LAST traced step of every rank: step span, time inside kernels (union of the launch intervals), idle time
between launches, in-kernel waiting for neighbour ranks, and a per-kernel table (count, total us, mean us, mean wait).
Times are %globaltimer differences on one GPU (never compared across GPUs)."""
import glob
import json
import sys
from collections import defaultdict


def union_len(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot


def main():
    prefix = sys.argv[1]
    files = sorted(glob.glob(prefix + ".rank*.json"), key=lambda p: int(p.split(".rank")[1].split(".")[0]))
    for path in files:
        d = json.load(open(path))
        st = d["steps"][-1]
        L = [x for x in st["launches"] if x["end_ns"] >= x["start_ns"] > 0]
        t0, t1 = min(x["start_ns"] for x in L), max(x["end_ns"] for x in L)
        busy = union_len([(x["start_ns"], x["end_ns"]) for x in L])
        wait = sum(x["wait_ns"] for x in L)
        print(f"rank {d['rank']}/{d['nranks']} rows {d['band']} step {st['step']}: span {1e-3 * (t1 - t0):.1f} us, "
              f"{len(L)} launches, in kernels {1e-3 * busy:.1f} us ({100 * busy / (t1 - t0):.0f} %), idle {1e-3 * (t1 - t0 - busy):.1f} us, "
              f"waiting for neighbours inside kernels {1e-3 * wait:.1f} us")
        agg = defaultdict(lambda: [0, 0, 0])
        for x in L:
            a = agg[x["kernel"]]
            a[0] += 1
            a[1] += x["end_ns"] - x["start_ns"]
            a[2] += x["wait_ns"]
        for k, (n, t, w) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"    {k:28s} x{n:4d}  total {1e-3 * t:8.1f} us  mean {1e-3 * t / n:6.2f} us  mean wait {1e-3 * w / n:6.2f} us")
        cp = d.get("cap_phases_ns") or []
        if cp:
            n = len(cp)
            m = [sum(r[k] for r in cp) / n * 1e-3 for k in range(len(cp[0]))]
            print(f"    k_cap phases of CTA 0 (mean of {n} launches): sweep {m[0]:.2f} us, grid barrier {m[1]:.2f} us, "
                  f"first polar item {m[2]:.2f} us, all items {m[3]:.2f} us")
            if len(m) >= 9:
                print(f"      inside the first item: loads + analysis {m[4]:.2f}, cluster sum {m[5]:.2f}, synthesis {m[6]:.2f}, "
                      f"cluster sum {m[7]:.2f}, stores {m[8]:.2f} us")
        # chain of one fast predict_correct in the middle of the step
        mid = L[len(L) // 2: len(L) // 2 + 12]
        print("    sample (us since step start: kernel start-end wait):")
        for x in mid:
            print(f"      {x['kernel']:28s} {1e-3 * (x['start_ns'] - t0):8.1f} - {1e-3 * (x['end_ns'] - t0):8.1f}  wait {1e-3 * x['wait_ns']:.1f}")


if __name__ == "__main__":
    main()
