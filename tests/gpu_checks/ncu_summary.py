"""One-line-per-kernel summary of an `ncu --set full` report (run where ncu is installed; no GPU needed).
    python tests/gpu_checks/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
        ("dram__bytes_read.sum", "DRAM rd MB", None), ("dram__bytes_write.sum", "DRAM wr MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM GB", None),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print("# " + path)
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0]
            parts = []
            for key, label, scale in WANT:
                if key not in ix:
                    continue
                v, u = r[ix[key]], units[ix[key]]
                try:
                    f = float(v)
                except ValueError:
                    continue
                if scale is None:   # bytes with a unit column
                    mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}.get(u, 1e-6)
                    f = f * mult
                    if label.endswith("GB"):
                        f /= 1e3
                else:
                    if key.startswith("gpu__time"):
                        f = f * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
                    else:
                        f *= scale
                parts.append("%s %.4g" % (label, f))
            print("%-44s %s" % (name[:44], " | ".join(parts)))


if __name__ == "__main__":
    main(sys.argv[1:])
