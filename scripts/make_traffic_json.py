#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu --set full summary (scripts/ncu_summary.py output): the DRAM bytes of the aggregation
launch group of ONE frame, stamped with the hash of the CUDA sources the capture was taken from (bench.py quotes the figure
only while the hash matches).

  python scripts/make_traffic_json.py profiles/r02_ncu_all_kernels.txt
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_hash  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    txt = open(sys.argv[1]).read()
    per = {}
    for block in txt.split("\n== ")[1:]:
        name = re.match(r"(?:void )?(\w+)", block).group(1)
        rd = re.search(r"dram read\s+([\d.]+) (\w+)", block)
        wr = re.search(r"dram write\s+([\d.]+) (\w+)", block)
        per[name] = {"read": float(rd.group(1)) * UNIT[rd.group(2)], "write": float(wr.group(1)) * UNIT[wr.group(2)]}
    agg = sum(per[k]["read"] + per[k]["write"] for k in ("k_sgm_sweeps", "k_sgm_final"))
    out = {"what": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the aggregation launch group for ONE frame of "
                   "BASELINE.json configs[1] (1280x960, D=192), crop-only aggregation (the product path), from one ncu --set full capture on B200",
           "sources": {k: f"{sys.argv[1]} ({per[k]['read'] / 1e9:.3f} GB read + {per[k]['write'] / 1e9:.3f} GB write)" for k in ("k_sgm_sweeps", "k_sgm_final")},
           "aggregation_dram_bytes_per_frame": int(agg), "source_hash": source_hash()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
