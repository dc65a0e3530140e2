#!/usr/bin/env python3
"""Record which kernel sources the committed ncu --set full capture belongs to, so that bench.py reports `roofline.traffic`
only while csrc/ still hashes to them (a stale constant would be worse than none).  Run after a capture:
    python scripts/stamp_traffic.py c2_plate_qeph_1m shell_forces 1582.98 "profiles/r02_qeph_forces_ncu.md (prof_r2m_qeph)"
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha
wl, kernel, bpe, cap = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
p = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
d[wl] = {"kernel": kernel, "dram_bytes_per_element": bpe, "capture": cap, "src_sha": csrc_sha()}
json.dump(d, open(p, "w"), indent=1)
print(d[wl])
