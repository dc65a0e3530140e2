#!/usr/bin/env python3
"""C2-sized plate (1 M QEPH shells) cut into many super-groups (alternating properties every `run` groups of 128): what a deck
with many parts costs per cycle against the single super-group of bench.py.  Usage (gpurun): python scripts/many_sg_bench.py [run ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openradioss_b200 import meshgen
from openradioss_b200.engine import Engine

runs = [int(a) for a in sys.argv[1:]] or [8192, 64, 16]
out = {}
for run in runs:
    m = meshgen.shell_plate(1000, 1000, 1000.0, 1000.0, pulse_tau=0.05)
    pa, pb = meshgen.default_prop_shell(thick=2.0), meshgen.default_prop_shell(thick=2.0)
    pb.h1 = pa.h1 * 1.25
    for k, sg in enumerate(m.shell_groups):
        sg.prop = pa if (k // run) % 2 == 0 else pb
    nsg = (len(m.shell_groups) + run - 1) // run
    g = Engine(m)
    g.run_cycles(20); g.synchronize()
    g.run_cycles(100); g.synchronize()
    out[str(nsg)] = g.last_run_ms() / 100
    del g
print(json.dumps({"ms_per_cycle_by_super_groups": out}))
