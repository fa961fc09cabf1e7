"""Compact text summary of an .ncu-rep (one block per captured kernel): duration, tensor/XU/issue utilisation, DRAM
bytes and throughput, L2->SM bytes, occupancy limits, top warp-stall reasons.

    ncu_summary.py report.ncu-rep > profiles/x.txt
    ncu_summary.py report.ncu-rep --traffic profiles/r01_traffic.json   # also writes avg DRAM bytes per launch
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(d, name):
    if name not in idx or d[idx[name]] in ("", "no data"):
        return None
    return float(d[idx[name]].replace(",", "")) * SCALE.get(units[idx[name]], 1)


print(f"# ncu --set full summary of {rep}")
tot = []
for d in data:
    print(f"\n## {d[idx['Kernel Name']][:110]}  grid={d[idx['Grid Size']]} block={d[idx['Block Size']]}")
    for w in WANT:
        if w in idx and d[idx[w]] not in ("", "no data"):
            print(f"  {w}: {d[idx[w]]} {units[idx[w]]}")
    st = []
    for h in hdr:
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[idx[h]] not in ("", "no data"):
            st.append((float(d[idx[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    st.sort(reverse=True)
    print("  top stalls (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))
    r, w = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    if r is not None and w is not None:
        tot.append(r + w)
if "--traffic" in sys.argv and tot:
    out = sys.argv[sys.argv.index("--traffic") + 1]
    json.dump(dict(kernel=data[0][idx["Kernel Name"]][:60], launches=len(tot), dram_bytes_per_launch=sum(tot) / len(tot),
                   source=rep, note="dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches"),
              open(out, "w"), indent=1)
