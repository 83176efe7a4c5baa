"""Per-launch DRAM traffic and local-memory instruction counts of k_message_tc from an `ncu --set full` report, written to
profiles/r02_ncu_traffic.json together with the hash of the kernel sources (bench.py reports `roofline.traffic` only for the
build the capture was taken from):  python tools/ncu_traffic.py gpurun_out/r2_msg.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import kernel_source_sha  # noqa: E402

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v, u = float(r[col[name]].replace(",", "")), units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return v * scale


out = []
for r in data:
    if "k_message_tc" not in r[col["Kernel Name"]]:
        continue
    out.append(dict(dram_bytes_read=val(r, "dram__bytes_read.sum"), dram_bytes_write=val(r, "dram__bytes_write.sum"),
                    duration_us=val(r, "gpu__time_duration.sum") / (1e3 if units[col["gpu__time_duration.sum"]].startswith("n") else 1.0),
                    local_load_inst=val(r, "smsp__inst_executed_op_local_ld.sum") if "smsp__inst_executed_op_local_ld.sum" in col else None,
                    local_store_inst=val(r, "smsp__inst_executed_op_local_st.sum") if "smsp__inst_executed_op_local_st.sum" in col else None,
                    inst_executed=val(r, "smsp__inst_executed.sum")))
assert out, "no k_message_tc launch in the report"
avg = {k: (sum(o[k] for o in out) / len(out) if out[0][k] is not None else None) for k in out[0]}
avg.update(src_sha=kernel_source_sha(), launches_captured=len(out), report=os.path.basename(rep),
           note="ncu --set full --clock-control none; bench.py --steps 2 --warmup 3 --no-cpu, message launches of layers >= 2")
with open(os.path.join(REPO, "profiles", "r02_ncu_traffic.json"), "w") as f:
    json.dump({"k_message_tc": avg}, f, indent=1)
print(json.dumps(avg, indent=1))
