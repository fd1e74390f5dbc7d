"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into per-kernel shares of ONE step
(the last complete step: from one rng_advance_kernel launch to the next / the end).
    python tests/gpu_checks/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            t = float(r[ix["Metric Value"]])
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]] if "Metric Unit" in ix else "ns"
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        launches.append((r[ix["Kernel Name"]], t * scale))
    marks = [i for i, (n, _) in enumerate(launches) if "rng_advance" in n]
    if len(marks) >= 2:
        step = launches[marks[-2]:marks[-1]]
        # the last mark starts the final step; use it when it is the longer (complete) one
        tail = launches[marks[-1]:]
        if len(tail) >= len(step):
            step = tail
    elif marks:
        step = launches[marks[-1]:]
    else:
        step = launches
    agg = collections.OrderedDict()
    for name, us in step:
        key = re.sub(r"\(.*", "", name).replace("void ", "").replace("univl::", "")
        key = re.sub(r"<.*", "", key) if not key.startswith("gemm") else key
        a = agg.setdefault(key, [0.0, 0])
        a[0] += us
        a[1] += 1
    total = sum(v[0] for v in agg.values())
    print("one training step: %d launches, %.3f ms summed device time (serialised, cold-cache per launch: shares are "
          "meaningful, the sum is not the step time)" % (len(step), total / 1e3))
    for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-62s %9.3f ms %5.1f%%  n=%d" % (k[:62], us / 1e3, 100 * us / total, n))


if __name__ == "__main__":
    main(sys.argv[1])
