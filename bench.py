#!/usr/bin/env python3
"""bench.py -- element-cycles/s of the explicit element cycle on N B200s (BASELINE.json metric).

A "step" is one explicit cycle (internal forces -> dt argmin -> /PARITH/ON assembly -> nodal
update) over the whole synthetic mesh.  `value` is measured with the model resident in HBM
(orgpu_run_cycles, CUDA events on the library's stream, max over ranks); `e2e` drives the same
cycles through the host-buffer C-ABI call orgpu_step_host (pinned host X/V in, X/V out, every
step) -- the usage pattern of the reference's own -gpu path, which re-uploads the nodal arrays
each cycle (shell_internal_forces.F90:106).  `--impl reference` times the CPU oracle restatement
(the reference Engine itself is Fortran and cannot be built in this image) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic HBM bytes per element-cycle (SURVEY.md 8d / BASELINE.md 3), split per kernel:
#   brick forces : IXS 32 + X,V gather 48 + state r/w 208 + SMSTR write 168 + corner rows write 256
#   node kernel  : corner rows read 256 + nodal update 152     (1 node / element)
B_ALG = {
    "brick": dict(total=1120, forces=712, node=408),
    "shell": dict(total=1872, forces=1408, node=464),   # forces: 16+72+48+416+560+40+256 ; node: 256+208
    # 3-node shell (C3FORC3, NPT=5, LAW36; not a BASELINE config, same counting rules as SURVEY 8d with half a node per
    # element): IXTG 12 + X,V,VR 36 + SMSTR 24 + element r/w 21 x 16 = 336 + IP 560 + VARTMP 40 + rows 3 x 64 = 192;
    # node side per node: rows 6 x 64 read + update 208
    "sh3n": dict(total=1200 + 192 + 296, forces=1200, node=592),
}


# C2 / C3 start from a smooth velocity field (amplitude mm/ms, wavelength mm) that drives the plate into yield within the first
# cycles: the timed cycles then run the plastic return of SIGEPS36C (sigeps36c.F:503-593) at most integration points instead of
# timing an idle elastic plate; `config.plastic_fraction` reports the share of integration points that yielded in the last cycle
VWAVE = (60.0, 100.0)
# ... and the timed window starts PREROLL cycles into that history (untimed, part of setting the model up): over its first ~150
# cycles 60-90 % of the plate's integration points yield at once, a state no crash deck stays in; from cycle 200 on it yields
# locally (10-30 % of the points per cycle, every point has yielded before) -- the state the metric is quoted on.  The per-regime
# kernel times are in profiles/r02_qeph_forces_ncu.md.
PREROLL = {"c2_plate_qeph_1m": 200, "c2_plate_qeph_1m_rates": 200, "c3_plate_qeph_4m": 200, "tri_plate_1m_yielding": 200, "bt_plate_1m_yielding": 200}

STRONG = ("c3_plate_qeph_4m", "c4_tube", "c4_tube_small", "c1_taylor_bar")    # total model fixed, cut into `world` domains


def workload(name, world=1):
    """Global model of the run.  Weak-scaling workloads (c2, c5): per-GPU work is fixed, the mesh grows along
    the decomposition axis with the GPU count.  Strong-scaling ones (c1, c3, c4): the BASELINE model, cut into
    `world` strips / slabs.  Returns (model, family of the dominant kernel, decomposition axis)."""
    from openradioss_b200 import meshgen
    if name == "c3_plate_qeph_4m":          # C3: 2000 x 2000 QEPH shells, LAW36, x strips
        return meshgen.shell_plate(2000, 2000, 2000.0, 2000.0, pulse_tau=0.05, vwave=VWAVE), "shell", 0
    if name == "c4_tube":                   # C4: 2.0 M QEPH shells (LAW36) + 501 k bricks (LAW2), imposed-velocity crush, z slabs
        return meshgen.crush_tube(708, 706, 1), "mixed", 2
    if name == "c4_tube_small":
        return meshgen.crush_tube(100, 100, 1), "mixed", 2
    if name == "c5_brick_slab_2m":          # C5: 200 x 200 x 50 bricks (2 M) per GPU, z slabs, LAW2
        return meshgen.hex_block(200, 200, 50 * world, 200.0, 200.0, 50.0 * world, vrand=1.0, vseed=12345), "brick", 2
    if name == "c1_taylor_bar":
        return meshgen.taylor_bar(1), "brick", 2
    if name == "brick_small":
        return meshgen.hex_block(40, 40, 40 * world, 40.0, 40.0, 40.0 * world, vrand=1.0), "brick", 2
    if name == "c2_plate_qeph_1m":          # C2 / C3: 1000 x 1000 QEPH shells per GPU (x strips), LAW36, NPT=5
        return meshgen.shell_plate(1000 * world, 1000, 1000.0 * world, 1000.0, pulse_tau=0.05, vwave=VWAVE), "shell", 0
    if name == "c2_plate_qeph_1m_elastic":  # the same plate starting from rest: the pressure pulse leaves it elastic over the timed cycles
        return meshgen.shell_plate(1000 * world, 1000, 1000.0 * world, 1000.0, pulse_tau=0.05), "shell", 0
    if name == "tri_plate_1m":              # extra: 707 x 707 cells x 2 = 999 698 3-node shells per GPU, LAW36, NPT=5
        return meshgen.tri_plate(707 * world, 707, 1000.0 * world, 1000.0), "sh3n", 0
    if name == "c2_plate_qeph_1m_rates":    # C2 with a rate-dependent /MAT/PLAS_TAB: three yield curves (strain rates 0, 0.1, 10 /ms: +0 / +8 / +20 %), strain-rate filter
        x = np.array([0.0, 0.01, 0.02, 0.05, 0.10, 0.15, 0.20, 0.30]); y = np.array([250.0, 290.0, 315.0, 355.0, 395.0, 420.0, 435.0, 450.0])
        return meshgen.shell_plate(1000 * world, 1000, 1000.0 * world, 1000.0, pulse_tau=0.05, vwave=VWAVE,
                                   curves=[(x, y), (x, 1.08 * y), (x, 1.2 * y)], rates=[0.0, 0.1, 10.0]), "shell", 0
    if name == "tri_plate_1m_yielding":     # the same plate driven into yield like C2
        return meshgen.tri_plate(707 * world, 707, 1000.0 * world, 1000.0, vwave=VWAVE), "sh3n", 0
    if name == "bt_plate_1m_yielding":      # 1 M Belytschko-Tsay shells (Ishell 1), LAW36, NPT=5, yielding like C2 (the family the reference's own GPU path covers)
        return meshgen.shell_plate(1000 * world, 1000, 1000.0 * world, 1000.0, pulse_tau=0.05, vwave=VWAVE, prop=meshgen.default_prop_shell(ihbe=1)), "shell", 0
    if name == "plate_small":
        return meshgen.shell_plate(200 * world, 200, 1000.0 * world, 1000.0), "shell", 0
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.rows, self.stop = dev, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self): self.t.start(); return self

    def __exit__(self, *a): self.stop = True; self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def csrc_sha():
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "openradioss_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_extra(name, world, rank, local, dist, steps, warmup):
    """Device-resident throughput of one more BASELINE configuration inside the same run (short: `steps` timed cycles): returns
    the entry for the line's `other_configs`.  Same engine, same timing rules as the headline (CUDA events on the library's
    stream, max over ranks)."""
    import torch
    from openradioss_b200.engine import Engine
    from openradioss_b200 import domdec
    gm, fam, axis = workload(name, world)
    if world > 1:
        dom = domdec.decompose_strips(gm, world, rank, axis=axis); m = dom.model
    else:
        m = gm
    ne = m.numels + m.numelc + m.numeltg
    g = Engine(m, device=local)
    if world > 1:
        g.comm_init(dist, dom)
    net = torch.tensor([ne], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(net)
    g.run_cycles(PREROLL.get(name, 0) + warmup); g.synchronize()
    if world > 1:
        dist.barrier()
    g.run_cycles(steps); g.synchronize()
    t = torch.tensor([g.last_run_ms()], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out = {"elements": int(net.item()), "family": fam, "n_gpus": world, "steps": steps, "ms_per_step": ms / steps,
           "value": int(net.item()) * steps / (ms * 1e-3), "unit": "element-cycles/s",
           "scaling": "strong" if name in STRONG else "weak"}
    if world == 1:
        g.set_profile(True); g.run_cycles(min(steps, 32)); g.synchronize()
        prof = {k: g.profile(i) for i, k in enumerate(("brick_forces", "shell_forces", "node"))}
        out["kernel_ms"] = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
        peak, _ = peaks()
        if fam in ("brick", "shell"):
            k = "brick_forces" if fam == "brick" else "shell_forces"
            nel = m.numels if fam == "brick" else m.numelc
            if prof[k][1]:
                out["roofline_frac"] = B_ALG[fam]["forces"] * nel / (prof[k][0] / prof[k][1] * 1e-3) / 1e9 / peak
    del g
    return out


def run_reference(args, rank):
    """CPU arm: the oracle restatement on all host cores, same workload / metric."""
    if rank != 0:
        return
    from oracle.orc import Oracle
    m, fam, _ = workload(args.workload)
    ne = m.numels + m.numelc + m.numeltg
    cores = os.cpu_count() or 1
    o = Oracle(m, threads=cores)
    preroll = PREROLL.get(args.workload, 0)
    o.run_cycles(preroll + max(1, args.warmup))                        # the same point of the plate's history as the GPU arm's timed window
    t0 = time.perf_counter(); o.run_cycles(args.steps); dt = time.perf_counter() - t0
    val = ne * args.steps / dt
    line = {"impl": "reference", "metric": "element-cycles/sec", "value": val, "unit": "element-cycles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "elements": ne, "nodes": m.numnod, "family": fam, "preroll_cycles": preroll},
            "cpu_baseline": {"value": val, "unit": "element-cycles/s", "cores": cores, "kind": "port",
                             "sample": f"{args.workload}: {ne} elements x {args.steps} cycles after {preroll + max(1, args.warmup)} untimed ones, OpenMP over groups of 128"},
            "e2e": {"value": val, "unit": "element-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="orgpu", choices=["orgpu", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ORGPU_WORKLOAD", "c2_plate_qeph_1m"))
    ap.add_argument("--cpu-cycles", type=int, default=20, help="cycles of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other BASELINE configurations (other_configs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40          # bounded: one CPU step of 2 M bricks is ~0.4 s on 8 cores
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from openradioss_b200.engine import Engine
    gm, fam, axis = workload(args.workload, world)
    if world > 1:                                   # this rank's domain of the global model
        from openradioss_b200 import domdec
        dom = domdec.decompose_strips(gm, world, rank, axis=axis)
        m = dom.model
        del gm
    else:
        m = gm
    ne = m.numels + m.numelc + m.numeltg
    g = Engine(m, device=local)
    if world > 1:
        g.comm_init(dist, dom)
    net = torch.tensor([ne], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(net)
    ne_total = int(net.item())

    def barrier():
        if world > 1:
            dist.barrier()
        g.synchronize(); torch.cuda.synchronize()

    # ---- /PARITH/ON evidence inside the multi-GPU run: a small plate + brick model stepped on `world` domains (same exchange
    # kernels, same graph) must end bit-identical to its single-domain run; Adler-32 as /DEBUG/CHKSM does (spmd_flush_accel.F)
    pon_check = None
    if world > 1:
        import zlib
        from openradioss_b200 import meshgen as mg, domdec as dd
        sm = mg.shell_on_block(24, 10, 3)
        sd = dd.decompose_strips(sm, world, rank, axis=0)
        sg = Engine(sd.model, device=local); sg.comm_init(dist, sd)
        sg.run_cycles(30); sg.synchronize()
        out = sg.download_nodes(("X", "V")); st = sg.time()
        parts = [None] * world
        dist.all_gather_object(parts, (sd.node_gid, out["X"], out["V"], st["neltst"], st["dt2"]))
        if rank == 0:
            ref = Engine(sm, device=local); ref.run_cycles(30); ref.synchronize()
            xr = ref.download_nodes(("X", "V")); tr = ref.time()
            X = np.zeros_like(xr["X"]); V = np.zeros_like(xr["V"])
            same = True
            for gid, x, v, nel, dt2 in parts:
                same = same and np.array_equal(x, xr["X"][gid]) and np.array_equal(v, xr["V"][gid]) and nel == tr["neltst"] and dt2 == tr["dt2"]
                X[gid] = x; V[gid] = v
            ck = lambda a, b: zlib.adler32(np.ascontiguousarray(b).tobytes(), zlib.adler32(np.ascontiguousarray(a).tobytes()))
            pon_check = {"model": f"shell_on_block 24x10x3 ({sm.numelc} shells + {sm.numels} bricks), {world} strips, 30 cycles",
                         "adler32_1_domain": ck(xr["X"], xr["V"]), f"adler32_{world}_domains": ck(X, V), "bitwise_identical": bool(same)}
            del ref
        del sg
        dist.barrier()

    # ---- device-resident throughput
    preroll = PREROLL.get(args.workload, 0)
    g.run_cycles(preroll + args.warmup); barrier()
    l0 = g.launch_count()
    with ClockSampler(local) as cs:
        barrier()
        g.run_cycles(args.steps)
        barrier()
    ms = g.last_run_ms()
    launches = g.launch_count() - l0
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = ne_total * args.steps / (ms * 1e-3)
    clocks = cs.summary()

    # ---- per-kernel durations (CUDA events around every launch on the library's stream) over the SAME cycles as the timed
    # region: a second engine built from the same initial state, the same warm-up, then `steps` profiled cycles (un-graphed, one
    # event pair per launch).  Across domains the exchange keeps the ranks in step, so the profiled pass follows the timed one.
    if world == 1:
        gp = Engine(m, device=local)
        gp.run_cycles(preroll + args.warmup); gp.synchronize()
        gp.set_profile(True); gp.run_cycles(args.steps); gp.synchronize()
        prof = {k: gp.profile(i) for i, k in enumerate(("brick_forces", "shell_forces", "node"))}
        del gp
    else:
        g.set_profile(True)
        g.run_cycles(min(args.steps, 64)); g.synchronize()
        prof = {k: g.profile(i) for i, k in enumerate(("brick_forces", "shell_forces", "node"))}
        g.set_profile(False)
    peak, peak_src = peaks()
    dom = "brick_forces" if fam == "brick" else "shell_forces"
    dms, dn = prof[dom]
    b = B_ALG["brick" if fam == "brick" else "sh3n" if fam == "sh3n" else "shell"]
    ne_dom = m.numels if fam == "brick" else m.numeltg if fam == "sh3n" else m.numelc   # elements the dominant kernel processes per launch
    if fam == "mixed":                                      # whole-cycle bytes: weighted sum of both families
        b = dict(b, total=(B_ALG["shell"]["total"] * m.numelc + B_ALG["brick"]["total"] * m.numels) / max(ne, 1))
    ach = (b["forces"] * ne_dom) / (dms / dn * 1e-3) / 1e9 if dn else 0.0
    nms, nn = prof["node"]
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None, "peak_source": peak_src, "bytes_per_element": b["forces"],
            "avg_launch_ms": dms / dn if dn else None, "launches_timed": dn,
            "window": "the timed cycles themselves (second engine, same initial state and warm-up, one event pair per launch)" if world == 1 else "cycles after the timed ones",
            "node_kernel": {"achieved": (b["node"] * m.numnod) / (nms / nn * 1e-3) / 1e9 if nn else None,
                            "avg_launch_ms": nms / nn if nn else None, "bytes_per_node": b["node"]},
            "whole_cycle": {"achieved": b["total"] * ne * world / (ms * 1e-3 / args.steps) / 1e9 / world,
                            "frac": b["total"] * ne / (ms * 1e-3 / args.steps) / 1e9 / peak, "bytes_per_element": b["total"]}}
    # DRAM bytes per launch from the committed `ncu --set full` capture of this workload -- only while the kernel sources are
    # the ones that were captured (sha of csrc/*.cu*), otherwise null: a stale constant would be worse than none
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            rec = json.load(open(tr)).get(args.workload, {})
            if rec.get("src_sha") == csrc_sha() and rec.get("kernel") == dom:
                roof["traffic"] = rec["dram_bytes_per_element"] * ne_dom
                roof["traffic_source"] = rec.get("capture")
        except Exception:
            pass

    # ---- the same plate from rest (the pressure pulse leaves it elastic over the window): the dominant kernel without the
    # plastic return, for comparison with the yielding headline above (same mesh, same kernel, fresh state)
    elastic = None
    if world == 1 and args.workload == "c2_plate_qeph_1m":
        import copy
        m0 = copy.copy(m); m0.V = np.zeros_like(m.V); m0.VR = np.zeros_like(m.VR)
        g0 = Engine(m0, device=local)
        g0.run_cycles(args.warmup); g0.synchronize()
        g0.set_profile(True); g0.run_cycles(min(args.steps, 200)); g0.synchronize()
        ems, en = g0.profile(1); g0.set_profile(False)
        if en:
            ea = (b["forces"] * ne_dom) / (ems / en * 1e-3) / 1e9
            elastic = {"avg_launch_ms": ems / en, "achieved": ea, "frac": ea / peak, "note": "same plate starting from rest: no integration point yields"}
        del g0
    roof["elastic_state"] = elastic

    # ---- how plastic the timed cycles were: integration points whose plastic strain grew during one more cycle
    plastic = None
    if world == 1 and (m.numelc or m.numeltg):
        st = g.shell_state if m.numelc else g.sh3n_state
        p0 = st("pla"); g.run_cycles(1); g.synchronize(); p1 = st("pla")
        plastic = {"last_cycle": float((p1 > p0).mean()), "ever": float((p1 > 0).mean()), "pla_max": float(p1.max())}

    # ---- end to end through the host-buffer C-ABI calls (pinned host arrays, copies inside the timed region).  Headline:
    # orgpu_forces_host, the cycle of the reference's own -gpu ABI in one call -- the host owns the nodal arrays, hands X, V (and VR
    # for models with rotational dofs) over every step and takes the assembled internal forces (8 doubles per node) and the time
    # step back; uploads, kernels and downloads pipelined over chunks of the node range.  Beside it: orgpu_step_host_rot (X, V, VR
    # in, the whole cycle incl. the nodal update on the device, X, V, VR out), which cannot overlap anything.
    n = m.numnod
    rot = bool(m.control.iroddl)
    narr = 3 if rot else 2
    pin = lambda w=3: torch.empty((n, w), dtype=torch.float64).pin_memory()
    hin = [pin() for _ in range(narr)]; hout = [pin() for _ in range(narr)]
    nd = g.download_nodes(("X", "V", "VR"))
    for t_, k in zip(hin, ("X", "V", "VR")):
        t_.numpy()[:] = nd[k]
    e2e_steps = max(3, min(args.steps, 50))

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_host_step():
        nonlocal hin, hout
        a = [t_.numpy() for t_ in hin]; b = [t_.numpy() for t_ in hout]
        g.step_host(a[0], a[1], a[2] if rot else None, 1, b[0], b[1], b[2] if rot else None)
        hin, hout = hout, hin
    sh_dt = timed(step_host_step)
    e2e_alt = {"value": ne_total * e2e_steps / sh_dt, "ms_per_step": 1e3 * sh_dt / e2e_steps, "h2d_bytes_per_step": 24 * narr * n, "d2h_bytes_per_step": 24 * narr * n,
               "call": "orgpu_step_host_rot (X,V,VR pinned host -> 1 cycle incl. nodal update -> X,V,VR host)" if rot else "orgpu_step_host (X,V pinned host -> 1 cycle -> X,V host)"}
    fh_ok = world == 1 and not m.control.nodadt
    if fh_ok:
        F8 = pin(8)
        dt1 = g.time()["dt2"]
        a = [t_.numpy() for t_ in hin]

        def forces_host_step():
            g.forces_host(a[0], a[1], a[2] if rot else None, dt1, F8.numpy())
        fh_dt = timed(forces_host_step)
        e2e = {"value": ne_total * e2e_steps / fh_dt, "unit": "element-cycles/s", "h2d_bytes_per_step": 24 * narr * n, "d2h_bytes_per_step": 64 * n + 64,
               "steps": e2e_steps, "ms_per_step": 1e3 * fh_dt / e2e_steps,
               "call": "orgpu_forces_host (X,V,VR pinned host -> internal forces of one cycle -> F(8,NUMNOD) + dt host; pipelined over node chunks)",
               "step_host": e2e_alt}
    else:
        e2e = dict(e2e_alt, unit="element-cycles/s", steps=e2e_steps)
    # the link itself: one pinned 24 n-byte array each way, timed alone (what bounds the calls above)
    dbuf = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dbuf.copy_(hin[0], non_blocking=True); torch.cuda.synchronize()
    ev[0].record(); dbuf.copy_(hin[0], non_blocking=True); ev[1].record(); hout[0].copy_(dbuf, non_blocking=True); ev[2].record(); torch.cuda.synchronize()
    h2d_gbs = 24e-9 * n / (ev[0].elapsed_time(ev[1]) * 1e-3); d2h_gbs = 24e-9 * n / (ev[1].elapsed_time(ev[2]) * 1e-3)
    e2e["pcie"] = {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                   "floor_ms_per_step": 1e3 * max(e2e["h2d_bytes_per_step"] / h2d_gbs, e2e["d2h_bytes_per_step"] / d2h_gbs) * 1e-9,
                   "note": "single copies alone at the measured link rate; floor = the longer direction, both at once"}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.orc import Oracle
        cores = os.cpu_count() or 1
        o = Oracle(m, threads=cores)
        o.run_cycles(max(2, preroll))                                  # the same point of the plate's history as the GPU's timed window
        t0 = time.perf_counter(); o.run_cycles(args.cpu_cycles); dt = time.perf_counter() - t0
        cpu = {"value": ne * args.cpu_cycles / dt, "unit": "element-cycles/s", "cores": cores, "kind": "port",
               "sample": f"{args.workload}: {ne} elements x {args.cpu_cycles} cycles (oracle restatement, OpenMP over groups of 128), after {max(2, preroll)} untimed cycles"}
        o.close()

    # ---- the other BASELINE configurations, briefly, in the same run: C1 (Taylor bar), C5 (brick slab, 2 M per GPU), C4 (crush tube:
    # shells + bricks) on one GPU; C3 (4 M shells, strong scaling) and C5 / C4 across the domains of a multi-GPU run
    extras = None
    if not args.no_extras and os.environ.get("ORGPU_BENCH_EXTRAS", "1") != "0" and args.workload == "c2_plate_qeph_1m":
        extras = {}
        del g
        torch.cuda.empty_cache()
        names = ("c1_taylor_bar", "c5_brick_slab_2m", "c4_tube") if world == 1 else ("c3_plate_qeph_4m", "c5_brick_slab_2m", "c4_tube")
        for nm in names:
            try:
                extras[nm] = measure_extra(nm, world, rank, local, dist if world > 1 else None, 100, 10)
            except Exception as ex:                                  # an extra must never cost the headline
                extras[nm] = {"error": str(ex)[:200]}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": "element-cycles/sec", "value": value, "unit": "element-cycles/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "family": fam, "elements_per_gpu": ne, "nodes_per_gpu": n,
                           "plastic_fraction": plastic, "preroll_cycles": preroll,
                           "l2": "inputs larger than L2 (element state >> 126 MB)" if ne >= 500000 else "working set may fit L2",
                           "parallelism": f"domains={world}" + ("" if world == 1 else " (strips / slabs; peer-memory corner-row exchange + dt fold per cycle, one CUDA graph)")},
                "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "pon_check": pon_check, "other_configs": extras,
                "e2e": e2e,
                "gpu_launches": launches,
                "kernel_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
