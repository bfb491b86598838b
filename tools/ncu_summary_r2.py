#!/usr/bin/env python
"""profiles/r2_ncu_summary.json from the raw page of the round-2 `ncu --set full` capture at the config-3 size
(gpurun_out/r2d_hot_raw.csv, written by tools/r2_gpu_pass.sh): measured DRAM bytes per launch of the dominant kernel of every
kernel class bench.py reports, for the `roofline.traffic` field.
usage: python tools/ncu_summary_r2.py gpurun_out/r2d_hot_raw.csv > profiles/r2_ncu_summary.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}
first = {}
for r in rows[2:]:
    n = r[col["Kernel Name"]]
    for key in ("ChunkFactorBody", "ChunkFwdBody", "ChunkBwdBody", "BandMatvecBody", "SchurBlockBody", "LinStereoTileBody", "StereoPoseBody",
                "StereoLmBody", "NodeAsmBody", "PairAsmBody"):
        if key in n and key not in first:
            first[key] = r


def gb(r, k):
    v = float(r[col[k]].replace(",", ""))
    u = units[col[k]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


def kern(key):
    r = first[key]
    return {"kernel": key, "dram_read_bytes": gb(r, "dram__bytes_read.sum"), "dram_write_bytes": gb(r, "dram__bytes_write.sum"),
            "time_ms": float(r[col["gpu__time_duration.sum"]]), "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]]}


Ns, P = 11112, 148
levels_sep = 8                      # ceil(log2(P - 1)) cyclic-reduction levels over the separators
out = {"_source": "ncu --set full --clock-control none, one config-3 solve (tools/ncu_target.py 100000), first launch of each kernel; "
                  "tools/r2_gpu_pass.sh r2d full; table in profiles/r2_ncu_hot_c3.txt",
       "_kernels": {k: kern(k) for k in first}}
f = kern("ChunkFactorBody")
out["bcr_factor"] = {"Ns": Ns, "band_chunks": P, "per": "try", "dominant_kernel": "ChunkFactorBody",
                     "dram_bytes_per_launch": f["dram_read_bytes"] + f["dram_write_bytes"],
                     "note": "one launch per factorization; the separator kernels (8 levels over 147 supernodes) are not included"}
fw, bw = kern("ChunkFwdBody"), kern("ChunkBwdBody")
out["bcr_solve"] = {"Ns": Ns, "band_chunks": P, "launches_per_unit": 2 + 1 + 2 * levels_sep + 1, "dominant_kernel": "ChunkFwdBody + ChunkBwdBody",
                    "dram_bytes_per_launch": fw["dram_read_bytes"] + fw["dram_write_bytes"] + bw["dram_read_bytes"] + bw["dram_write_bytes"],
                    "note": "bytes of one band solve (forward + backward sweep over the chunks); unit = one solve = 20 launches of the class"}
m = kern("BandMatvecBody")
out["matvec"] = {"Ns": Ns, "band_chunks": P, "launches_per_unit": 1, "dominant_kernel": "BandMatvecBody",
                 "dram_bytes_per_launch": m["dram_read_bytes"] + m["dram_write_bytes"]}
s = kern("SchurBlockBody")
out["schur"] = {"Ns": Ns, "band_chunks": P, "per": "try", "dominant_kernel": "SchurBlockBody",
                "dram_bytes_per_launch": s["dram_read_bytes"] + s["dram_write_bytes"],
                "note": "the kernel only; the device-to-device copy of the base system (2.3 GB per try) is a memcpy, not a kernel"}
print(json.dumps(out, indent=1))
