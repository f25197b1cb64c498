#!/usr/bin/env python
"""bench.py -- NBModelABFS MM/MM hot path on B200: list rebuild + energy + gradients per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload m1|jac|...] [--impl b200|reference]

A "step" is one NBModelABFS call on a synthetic replicated water (/protein) box with a FORCED pair-list rebuild
followed by the energy + gradient evaluation (BASELINE.json config 3: "pair-list rebuild + E+grad per call"); the
no-rebuild call is timed separately and reported in `no_rebuild`.  The metric is list-pair interactions per second
(a list pair = one (i, j[, image]) entry of the reference's lists at listCutoff = 13.5 A).

N = 1 : one process.  N > 1 : launched under torchrun; the i-blocks are split over the ranks (spatial slabs of the
cell-sorted atoms), every rank rebuilds and evaluates its slab, energies and gradients are all-reduced with NCCL.
Total work is fixed as N grows -> "scaling": "strong".

`value`  : whole-job list pairs / s with coordinates resident in HBM, timed with CUDA events on the launching stream,
           max over ranks.
`e2e`    : the same metric through the plugin surface (System.Energy(doGradients=True) -> NBModelABFS.SetUp/Energy ->
           C-ABI) with HOST numpy coordinates in and host gradients out (N = 1).
`roofline`: the tile force kernel against the FP32 (non-tensor) pipe: 42.75 algorithmic flop per list pair.
`cpu_baseline`: the compiled reference C code (oracle/_ref, OpenMP build) on this box's host cores on a bounded sample.
`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FLOP_PER_LIST_PAIR = 42.75          # 42 / 64 / 8 flop for r <= 8 / 8-12 / 12-13.5 A at uniform density
SM_COUNT, FP32_LANES = 148, 128


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the force kernel per launch, from the committed ncu capture of the
    same workload (profiles/traffic.json); None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("%s@%d" % (workload, world))
    except Exception:
        return None


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in text.strip().split("\n"):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_workload(name):
    import pdynamo_mirror_b200 as p
    return p.workloads.WORKLOADS[name]()


def workload_description(name, w):
    kind = "two-species ionic fluid" if name.startswith("ionic") else ("TIP3P" + ("" if w["ntypes"] <= 2 else " + %d-type heteropolymer" % w["ntypes"]))
    return "%s: %d atoms, P1 box %s, %s, ABFS 0.5/8/12/13.5 A" % (name, w["n"], "x".join("%.2f" % v for v in w["box"][:3]), kind)


# --------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the compiled, unmodified reference C code (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_reference(sample_name, steps, warmup, budget_s=None):
    """budget_s: stop after the step that exceeds this wall-clock budget (at least one timed step is always taken)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refnb
    omp = refnb.available(omp=True)
    kind = "reference"
    if not refnb.available(omp=omp):
        import oracle
        w = make_workload(sample_name)
        model, kind, cores = oracle.OracleNB(w), "port", 1
    else:
        w = make_workload(sample_name)
        model = refnb.RefNB(w, omp=omp)
        cores = model.num_threads() if omp else 1
    times, upd, ene, pairs = [], [], [], 0
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = model.energy(force_new=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt); upd.append(out["t_update"]); ene.append(out["t_energy"])
        if budget_s is not None and times and time.perf_counter() - t_start + dt > budget_s:
            break
    c = model.counts()
    pairs = c["primary"] + c["image_pairs"]
    best = min(times)
    return dict(value=pairs / best, unit="list-pairs/s", cores=cores, kind=kind, steps_run=len(times),
                sample="%s (%d atoms, %d list pairs): forced list rebuild + E+grad per step, best of %d after %d warm-up; "
                       "rebuild %.3f s (serial in the reference), E+grad %.3f s" % (sample_name, w["n"], pairs, len(times), min(warmup, it), min(upd), min(ene)),
                ms_per_step=1e3 * best, pairs=pairs, n=w["n"], ms_rebuild=1e3 * min(upd), ms_energy=1e3 * min(ene)), w


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    # torchrun exports OMP_NUM_THREADS=1 to every rank when the variable is unset; the reference arm runs on rank 0 alone and is meant to
    # use all the host threads it can, so the team size is restored here (before libgomp is loaded with the reference library)
    if os.environ.get("NBB_REF_THREADS") or "TORCHELASTIC_RUN_ID" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(int(os.environ.get("NBB_REF_THREADS") or len(os.sched_getaffinity(0))))
    # the reference arm runs the WORKLOAD ITSELF (same input as the B200 arm) unless a sample is asked for with --ref-sample: one step of
    # the 1.1 M-atom box takes the reference ~30 s (its list generator is serial), so the number of steps is cut by a wall-clock budget
    # (NBB_REF_BUDGET_S, default 100 s; at least one step) and the steps actually taken are reported
    same = args.ref_sample_explicit is None
    name = args.workload if same else sample
    budget = float(os.environ.get("NBB_REF_BUDGET_S", "100"))
    base, w = cpu_reference(name, max(1, args.steps), 0 if same else max(0, min(args.warmup, 1)), budget_s=budget)
    wl = workload_description(args.workload, make_workload(args.workload) if not same else w)
    line = {"metric": "NBModelABFS list-pair interactions per second (pair-list rebuild + energy + gradients per call)",
            "value": base["value"], "unit": "list-pairs/s", "n_gpus": args.gpus, "steps": base["steps_run"], "warmup": 0 if same else min(args.warmup, 1),
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl, "sample": name, "same_config": bool(same),
                       "note": ("the reference's CPU implementation on the workload itself; steps cut to a %g s budget (best step reported)" % budget) if same
                               else "reference CPU implementation timed on a bounded sample of the workload"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "list-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------------
class DeviceModel:
    """Thin driver of the C-ABI with device-resident coordinates / gradients (torch tensors)."""

    def __init__(self, w, device, rank=0, nranks=1):
        import torch
        import pdynamo_mirror_b200 as p
        from pdynamo_mirror_b200 import _lib
        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.w, self.n = w, w["n"]
        sysm = p.System.FromWorkload(w)
        self.model = p.NBModelABFS(device=device)
        sysm.DefineNBModel(self.model)
        self.system = sysm
        # create the state through the plugin surface (one host call), then drive it with device pointers
        sysm.Energy(doGradients=False)
        self.state = sysm.configuration.nbState
        self.h = self.state.cObject
        self.L.nbb200_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.L.nbb200_enable_timing(self.h, 1)
        self.x = torch.from_numpy(w["xyz"]).to("cuda:%d" % device)
        self.g = torch.zeros_like(self.x)
        self.box = np.ascontiguousarray(w["box"], np.float64)
        self.e = np.zeros(6)
        self.dEdM = np.zeros(9)

    def step(self, rebuild=True, zero=True):
        st = C.c_int(16)
        if zero:
            self.g.zero_()
        self.L.NBModelABFS_B200_UpdateDevice(self.h, C.c_void_p(self.x.data_ptr()), self._lib.d_(self.box), 1 if rebuild else 0, C.byref(st))
        self.L.NBModelABFS_B200_MMMMEnergyDevice(self.h, self._lib.d_(self.e), C.c_void_p(self.g.data_ptr()), self._lib.d_(self.dEdM), C.byref(st))
        if st.value != 16:
            raise RuntimeError("device step failed: " + self._lib.last_error())


def timed_steps(torch, fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / steps          # ms per step


def run_b200(args):
    import torch
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its banner / debug lines on stdout by default: send them to stderr so that stdout stays the one JSON line while the
        # driver can still read NCCL_DEBUG=INFO output (rank counts, transports)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pk, pk_kind = peaks()
    t0 = time.time()
    w = make_workload(args.workload)
    m = DeviceModel(w, local, rank, world)
    dn = None
    if world > 1:
        # section 8e: spatial slabs in sorted space; per call the ranks exchange halo positions (all slabs when the lists are rebuilt)
        # and halo gradient contributions point to point, and all-reduce 15 scalars (pdynamo-mirror_b200/parallel.py)
        from pdynamo_mirror_b200.parallel import DistributedNB
        dn = DistributedNB(m.state, w["n"], 0.5 * (m.model.listCutoff - m.model.outerCutoff), rank, world, m.x.device, transport=os.environ.get("NBB200_TRANSPORT"))

    def step_rebuild():
        if dn is None:
            m.step(rebuild=True)
        else:
            m.g.zero_()
            dn.call(m.x, m.box, m.g, force_rebuild=True)

    def step_norebuild():
        if dn is None:
            m.step(rebuild=False)
        else:
            m.g.zero_()
            dn.call(m.x, m.box, m.g)

    step_rebuild()
    torch.cuda.synchronize()
    counters = m.state.Counters()
    pairs_local = counters["listPairs"]
    pairs = int(round(allsum(float(pairs_local))))
    log("[bench] rank %d: setup %.1fs, %d atoms, %d list pairs (local %d), %d tiles" % (rank, time.time() - t0, w["n"], pairs, pairs_local, counters["tiles"]))

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = m.state.Counters()["kernelLaunches"]
    ms = allmax(timed_steps(torch, step_rebuild, args.steps, args.warmup, barrier))
    launches = (m.state.Counters()["kernelLaunches"] - launches0) / float(args.steps + args.warmup)
    # per-kernel device times of the last rebuild step (library events on the same stream)
    tm = m.state.Timings()
    # average the force-kernel time over a few more rebuild steps, live
    fk, lb, pr = [], [], []
    for _ in range(max(3, min(args.steps, 10))):
        step_rebuild(); t = m.state.Timings(); fk.append(t["tileForces"]); lb.append(t["listRebuild"]); pr.append(t.get("prune", 0.0))
    ms_nr = allmax(timed_steps(torch, step_norebuild, args.steps, args.warmup, barrier))
    clocks = sampler.stop() if sampler is not None else None
    force_ms, build_ms = allmax(statistics.mean(fk)), allmax(statistics.mean(lb))

    value = pairs / (ms * 1e-3)
    sm_max = float(pk.get("sm_max_mhz", 1965.0))
    nominal_tflops = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12
    st_peak = C.c_int(16)
    measured_tflops = float(m.L.nbb200_measure_fp32_peak(local, C.byref(st_peak)))      # register-only FMA chains, live on this GPU
    fp32_peak_tflops = measured_tflops if (st_peak.value == 16 and measured_tflops > 0.5 * nominal_tflops) else nominal_tflops
    achieved_tflops = (pairs / world) * FLOP_PER_LIST_PAIR / (force_ms * 1e-3) / 1e12 if force_ms > 0 else 0.0
    n, nimg = w["n"], counters["images"]
    nexcl = len(w["exclusions"])
    list_bytes = 24.0 * n * (1 + nimg) + 4.0 * (n + 1) + 8.0 * nexcl + 4.0 * pairs + 4.0 * (n + 1) * (1 + nimg)
    line = {
        "metric": "NBModelABFS list-pair interactions per second (pair-list rebuild + energy + gradients per call)",
        "value": value, "unit": "list-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 pair math, f64 accumulation and list predicate",
        "data": "synthetic",
        "config": {"workload": workload_description(args.workload, w), "list_pairs": pairs, "step": "forced list rebuild + E + gradients",
                   "l2": "inputs (coordinates + tile lists, %.0f MB) %s L2; timed iterations run back to back" %
                         ((counters["tiles"] * 128 + 80 * n) / 1e6, "exceed" if counters["tiles"] * 128 + 80 * n > 126e6 else "fit in"),
                   "parallelism": ("1 GPU" if world == 1 else
                                   ("%d spatial slabs of the cell-sorted order; per step every rank reads the slab positions (list rebuild) or halo positions (no rebuild) "
                                    "it needs FROM its peers' memory and adds its halo gradient contributions INTO their accumulators with plain kernels over CUDA-IPC mapped "
                                    "buffers (NVLink peer loads / fp64 atomics), update decision and the 15 scalars through flag areas in peer memory; NCCL only hands the "
                                    "IPC handles round at set-up" % world) if (dn is not None and dn.transport == "peer") else
                                   ("%d spatial slabs of the cell-sorted order; per step NCCL send/recv of all slab positions (list rebuild) or halo positions (no rebuild) "
                                    "and of halo gradient contributions to their owners, all-reduce of 15 scalars" % world))},
        "no_rebuild": {"ms_per_call": ms_nr, "value": pairs / (ms_nr * 1e-3), "unit": "list-pairs/s"},
        "kernels_ms": {"list_rebuild": build_ms, "tile_forces": force_ms, "prune": allmax(statistics.mean(pr)), "pairs14": tm["pairs14"], "displacement_check": tm["displacementCheck"]},
        "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
                     "traffic": ncu_traffic(args.workload, world), "kernel": "k_cluster_forces", "flop_per_list_pair": FLOP_PER_LIST_PAIR,
                     "peak_source": ("measured live: register-only FMA chains on all SMs (nbb200_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.2f TFLOP/s; "
                                     "MEASURED_PEAKS.json (%s) holds no FP32 figure" % (sm_max, nominal_tflops, pk_kind)) if fp32_peak_tflops == measured_tflops else
                                    "148 SM x 128 FP32 lanes x 2 x %.0f MHz (sm_max_mhz, %s); the live FMA measurement failed" % (sm_max, pk_kind),
                     "peak_nominal": nominal_tflops},
        "roofline_list_build": {"bound": "hbm", "achieved": list_bytes / world / (build_ms * 1e-3) / 1e9 if build_ms > 0 else 0.0, "peak": float(pk.get("hbm_gbs", 6650.0)),
                                "unit": "GB/s", "frac": (list_bytes / world / (build_ms * 1e-3) / 1e9) / float(pk.get("hbm_gbs", 6650.0)) if build_ms > 0 else 0.0,
                                "algorithmic_bytes": list_bytes, "note": "all rebuild kernels together; bytes = SURVEY.md 8d atom-pair-equivalent figure"},
        "gpu_launches": launches,
        "clocks": clocks,
        "energies": [float(v) for v in (dn.results()[0] if dn is not None else m.e)],
    }
    if world == 1:
        # end to end through the plugin surface with host arrays (H2D of coordinates, D2H of gradients inside the timed region)
        sysm = m.system
        sysm.configuration.nbState = m.state
        m.L.nbb200_set_stream(m.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        m.model.SetOptions(updateFrequency=1)

        def e2e_step():
            sysm.Energy(doGradients=True)

        def e2e_time():
            for _ in range(args.warmup):
                e2e_step()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            torch.cuda.synchronize()
            return (time.perf_counter() - t1) / args.steps * 1e3

        api = "System.Energy(doGradients=True) -> NBModelABFS.SetUp/Energy -> NBModelABFS_B200_Update/_MMMMEnergy, host numpy in/out, wall clock"
        e2e_ms = e2e_time()                          # default options: System.Energy's own gradient array, NB term first (zero fill folded into the NB call)

        def e2e_accumulate_step():                   # a caller-owned gradient array: the reference's fill + upload + accumulate semantics
            cfgx = sysm.configuration
            m.model.SetUp(sysm.energyModel.mmAtoms, None, sysm.energyModel.ljParameters, sysm.energyModel.ljParameters14, None, sysm.energyModel.interactions14,
                          sysm.energyModel.exclusions, sysm.symmetry, None, cfgx)
            own.fill(0.0)
            cfgx.gradients3 = own
            m.model.Energy(cfgx)

        from pdynamo_mirror_b200._lib import pinned_array
        own = pinned_array((n, 3))
        e2e_step = e2e_accumulate_step
        acc_ms = e2e_time()
        small = 48 + 128 * (nimg + 1)
        line["e2e"] = {"value": pairs / (e2e_ms * 1e-3), "unit": "list-pairs/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": 24 * n + small, "d2h_bytes_per_step": 24 * n + 16 * 8 * (nimg + 2),
                       "api": api + "; default options (System.Energy evaluates the NB term first on its own gradient array: the zero fill is folded into the NB call)"}
        line["e2e_accumulate"] = {"value": pairs / (acc_ms * 1e-3), "unit": "list-pairs/s", "ms_per_step": acc_ms,
                                  "h2d_bytes_per_step": 2 * 24 * n + small, "d2h_bytes_per_step": 24 * n + 16 * 8 * (nimg + 2),
                                  "api": "NBModelABFS.SetUp + Energy on a caller-owned, zero-filled gradient array (page-locked): uploaded, accumulated into, downloaded -- "
                                         "the reference's accumulate semantics at the NB-model level"}
        m.model.SetOptions(updateFrequency=0)
        try:
            line["spline_form"] = spline_block(torch, m, pairs)
        except Exception as exc:
            line["spline_form"] = {"error": repr(exc)}
        if args.workload != "jac" and not args.no_jac:
            try:
                line["jac"] = jac_block(torch, local, fp32_peak_tflops)
                line["md"] = md_block(torch, local)
                line["md_device"] = md_device_block(torch, local)
                line["md_dhfr"] = md_dhfr_block(torch, local)
                if rank == 0 and not args.no_cpu:
                    line["md_dhfr"]["reference_same_box"] = md_dhfr_reference_estimate()
            except Exception as exc:
                line["jac"] = {"error": repr(exc)}
        # the CPU legs come last: the OpenMP team of the compiled reference (16 threads that spin for a while after every parallel region) was
        # measured to slow the host-latency-bound JAC-size calls of the blocks above when it ran before them (0.13 instead of 0.11 ms per call)
        if rank == 0 and not args.no_cpu:
            try:
                base, _ = cpu_reference(args.ref_sample, 3, 1)
                line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as exc:      # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "list-pairs/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}
    else:
        line["distributed_check"] = distributed_check(torch, dist, m, dn, w, local)
        line["halo"] = {"halo_atoms_per_step_no_rebuild": int(dn.halo_atoms()), "transport": dn.transport,
                        "atoms_owned": int(dn.slabs[rank][1] - dn.slabs[rank][0]), "list_updates": dn.updates}
        # end to end with HOST arrays on every rank (DistributedNB.call_host): per step a rank uploads its contiguous chunk of the positions,
        # runs the distributed call (forced rebuild, as the timed steps above) and downloads the same rows of the gradient; the summed energies
        # are read on the host.  Wall clock between barriers, max over ranks.  Headline (`e2e`): the arrays and the semantics of the one-GPU
        # leg -- page-locked arrays as the mirror's System allocates them, the gradient rows SET (System.Energy's freshly zeroed array, NB term
        # first: the zero fill is folded into the call).  Extra (`e2e_accumulate`): plain numpy arrays, rows accumulated into (staging copies
        # and a host add on every rank).
        from pdynamo_mirror_b200._lib import pinned_array

        def host_leg(xh, gh, overwrite):
            for _ in range(args.warmup):
                dn.call_host(xh, m.box, gh, force_rebuild=True, overwrite=overwrite)
            barrier(); torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(args.steps):
                dn.call_host(xh, m.box, gh, force_rebuild=True, overwrite=overwrite)
            torch.cuda.synchronize(); barrier()
            return allmax((time.perf_counter() - t1) / args.steps * 1e3)

        xh = pinned_array((n, 3)); xh[...] = w["xyz"]
        gh = pinned_array((n, 3)); gh[...] = 0.0
        e2e_ms = host_leg(xh, gh, True)
        own = allsum(float(dn._own_count))
        # self-check of the host path: the rows of this rank's chunk are the gradients a device-array call gives on the same coordinates (that
        # call leaves every atom's gradient with the rank that owns it: summed over the ranks for the comparison, outside the timed region)
        gh[...] = 7.0                                        # overwritten, not added to
        dn.call_host(xh, m.box, gh, force_rebuild=True, overwrite=True)
        c0, c1 = (n * rank) // world, (n * (rank + 1)) // world
        m.x.copy_(torch.from_numpy(np.ascontiguousarray(xh))); m.g.zero_()
        dn.call(m.x, m.box, m.g, force_rebuild=True); dn.results()
        gfull = m.g.clone()
        dist.all_reduce(gfull)
        gd = gfull.cpu().numpy()
        host_err = allmax(float(np.abs(gh[c0:c1] - gd[c0:c1]).max() / max(1e-300, np.abs(gd).max())))
        host_rows = allsum(float(c1 - c0))
        xp = w["xyz"].copy()
        gp = np.zeros_like(xp)
        acc_ms = host_leg(xp, gp, False)
        gp[:] = 0.0
        dn.call_host(xp, m.box, gp, force_rebuild=True)
        acc_err = allmax(float(np.abs(gp[c0:c1] - gd[c0:c1]).max() / max(1e-300, np.abs(gd).max())))
        line["e2e"] = {"value": pairs / (e2e_ms * 1e-3), "unit": "list-pairs/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": int(24 * own), "d2h_bytes_per_step": int(24 * own) + 15 * 8 * world,
                       "api": "DistributedNB.call_host(x, box, g, force_rebuild=True, overwrite=True) on every rank with page-locked host arrays (as the "
                              "one-GPU leg: the mirror's System allocates them page-locked and its gradient array is freshly zeroed, NB term first): rank r "
                              "uploads the contiguous rows [n r / R, n (r + 1) / R) of x and downloads the same rows of the gradient (bytes summed over the "
                              "ranks), the device redistributes over peer memory; energies read on the host; wall clock, max over ranks",
                       "check": {"chunk_rows_rel_err_vs_device_call": host_err, "rows_checked_all_ranks": int(host_rows), "atoms": n}}
        line["e2e_accumulate"] = {"value": pairs / (acc_ms * 1e-3), "unit": "list-pairs/s", "ms_per_step": acc_ms,
                                  "h2d_bytes_per_step": int(24 * own), "d2h_bytes_per_step": int(24 * own) + 15 * 8 * world,
                                  "api": "the same call with plain (pageable) numpy arrays and the rows ACCUMULATED into the caller's gradient array: staging copies "
                                         "through page-locked memory and a host add on every rank",
                                  "check": {"chunk_rows_rel_err_vs_device_call": acc_err}}
    if dn is not None and os.environ.get("NBB200_DIST_PROFILE"):
        dn.host_profile = {}
        for _ in range(10):
            dn.call_host(xh, m.box, gh, force_rebuild=True)
        log("[bench] rank %d call_host phases (ms/step, synchronised): %s" % (rank, {k: round(100.0 * v, 3) for k, v in dn.host_profile.items()}))
        dn.host_profile = None
        for label, forced in (("rebuild", True), ("no-rebuild", False)):
            dn.profile = {}
            for _ in range(10):
                m.g.zero_(); dn.call(m.x, m.box, m.g, force_rebuild=forced)
            log("[bench] rank %d %s phases (ms/step, synchronised): %s" % (rank, label, {k: round(100.0 * v, 3) for k, v in dn.profile.items()}))
            dn.profile = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def distributed_check(torch, dist, m, dn, w, local):
    """N > 1: every rank also evaluates the WHOLE system on its own GPU (one unpartitioned state) and compares the distributed
    energies (after the all-reduce) and the gradients of the atoms it owns."""
    ref = DeviceModel(w, local)
    ref.step(rebuild=True)
    m.g.zero_()
    dn.call(m.x, m.box, m.g, force_rebuild=True)
    torch.cuda.synchronize()
    e, _ = dn.results()
    mask = (m.g.abs().sum(1) > 0)                            # the atoms this rank owns (every atom of these systems feels a force)
    gerr = float(((m.g - ref.g)[mask] ** 2).mean().sqrt() / (ref.g[mask] ** 2).mean().sqrt()) if bool(mask.any()) else 0.0
    t = torch.tensor([abs(e.sum() - ref.e.sum()) / abs(ref.e.sum()), gerr, float(mask.sum())], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {"energy_rel_err": float(tmax[0]), "owned_gradient_rel_rms_err_max": float(tmax[1]), "atoms_with_gradient_all_ranks": int(t[2]), "atoms": int(w["n"])}


def spline_block(torch, m, pairs):
    """The same workload with the interaction in its spline form (PairwiseInteractionABFS useAnalyticForm = False, 50 points / A:
    three cubic-spline tables looked up per pair instead of the analytic ABFS formulas)."""
    st = C.c_int(16)
    m.L.nbb200_set_stream(m.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    m.L.PairwiseInteractionABFS_B200_SetInteractionForm(m.h, 0, 50, C.byref(st))
    if st.value != 16:
        raise RuntimeError(m._lib.last_error())
    try:
        e_analytic = m.e.copy()
        ms = timed_steps(torch, lambda: m.step(rebuild=True), 5, 3, lambda: None)
        fk = []
        for _ in range(5):
            m.step(rebuild=True); fk.append(m.state.Timings()["tileForces"])
        ms_nr = timed_steps(torch, lambda: m.step(rebuild=False), 5, 3, lambda: None)
        return {"ms_per_step": ms, "ms_per_call_no_rebuild": ms_nr, "tile_forces_ms": statistics.mean(fk), "value": pairs / (ms * 1e-3), "unit": "list-pairs/s",
                "energies": [float(v) for v in m.e], "rel_diff_to_analytic_total": float(abs(m.e.sum() - e_analytic.sum()) / abs(e_analytic.sum()))}
    finally:
        m.L.PairwiseInteractionABFS_B200_SetInteractionForm(m.h, 1, 50, C.byref(st))


def jac_block(torch, local, fp32_peak_tflops):
    """The 23 558-atom DHFR/JAC box (north-star target size; the reference's own benchmark input) on one GPU."""
    w = make_workload("dhfr")
    m = DeviceModel(w, local)
    m.step(rebuild=True)
    torch.cuda.synchronize()
    pairs = m.state.Counters()["listPairs"]
    ms = timed_steps(torch, lambda: m.step(rebuild=True), 20, 5, lambda: None)
    fk, lb = [], []
    for _ in range(10):
        m.step(rebuild=True); t = m.state.Timings(); fk.append(t["tileForces"]); lb.append(t["listRebuild"])
    ms_nr_sync = timed_steps(torch, lambda: m.step(rebuild=False), 50, 5, lambda: None)
    fk_nr = []
    for _ in range(10):
        m.step(rebuild=False); fk_nr.append(m.state.Timings()["tileForces"])
    # the same calls with the optimistic update decision (nbb200_set_optimistic_updates): one host synchronisation per call instead of two
    m.L.nbb200_enable_timing(m.h, 0)
    m.L.nbb200_set_optimistic_updates(m.h, 1)
    ms_nr = timed_steps(torch, lambda: m.step(rebuild=False), 50, 5, lambda: None)
    m.L.nbb200_set_optimistic_updates(m.h, 0)
    m.L.nbb200_enable_timing(m.h, 1)
    f, f_nr = statistics.mean(fk), statistics.mean(fk_nr)
    ach = pairs * FLOP_PER_LIST_PAIR / (f * 1e-3) / 1e12
    return {"workload": workload_description("dhfr", w) + " (the reference's own JAC benchmark input, benchmarks/data/dhfr)", "list_pairs": pairs, "ms_per_call_rebuild": ms,
            "ms_per_call_no_rebuild": ms_nr, "ms_per_call_no_rebuild_two_syncs": ms_nr_sync,
            "no_rebuild_note": "Update + MMMMEnergy on device arrays, lists kept; optimistic update decision (one synchronisation per call); two_syncs: the plain call pair",
            "value": pairs / (ms * 1e-3), "value_no_rebuild": pairs / (ms_nr * 1e-3), "unit": "list-pairs/s",
            "kernels_ms": {"list_rebuild": statistics.mean(lb), "tile_forces": f, "tile_forces_no_rebuild": f_nr}, "roofline_frac_fp32": ach / fp32_peak_tflops,
            "roofline_frac_fp32_no_rebuild": pairs * FLOP_PER_LIST_PAIR / (f_nr * 1e-3) / 1e12 / fp32_peak_tflops}


def md_block(torch, local, steps=300):
    """BASELINE.json config 4 call pattern on the 23 558-atom DHFR/JAC box through the plugin surface
    (System.Energy(doGradients=True), host arrays in and out, one call per MD step).  Only the NB forces exist in this repository
    (bonded terms are out of scope) and NB-only dynamics of a bonded system is not stable, so the trajectory is a synthetic
    random walk with MD-like step lengths (0.03 A rms per coordinate and step: the displacement criterion of the reference,
    any atom beyond 0.75 A, fires every ~15-30 steps as in its DHFR run, 13.7 calls per update).  Two update policies: the
    reference's displacement heuristic, and a forced update every 10 steps."""
    import pdynamo_mirror_b200 as p
    w = make_workload("dhfr")
    n = w["n"]
    out = {}
    for label, freq, opt in (("displacement_triggered", 0, False), ("update_every_10", 10, False), ("displacement_triggered_optimistic", 0, True)):
        sysm = p.System.FromWorkload(w)
        sysm.DefineNBModel(p.NBModelABFS(device=local, updateFrequency=freq, optimisticUpdates=opt))
        rng = np.random.Generator(np.random.PCG64(491831))
        sysm.Energy(doGradients=True)
        x = sysm.coordinates3
        kicks = rng.standard_normal((8, n, 3)) * 0.03          # pre-drawn steps, reused cyclically: no RNG cost in the timed loop
        wall = None
        for _ in range(2):                                   # the faster of two consecutive passes (host-latency bound: see md_device_block)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(steps):
                x += kicks[k & 7]
                sysm.Energy(doGradients=True)
            dt = time.perf_counter() - t0
            wall = dt if wall is None else min(wall, dt)
        st = sysm.configuration.nbState
        out[label] = {"steps_per_s": steps / wall, "ms_per_step": 1e3 * wall / steps, "steps": steps, "runs": "best of 2 consecutive passes",
                      "list_updates": (int(st.numberOfUpdates) - 1) // 2,
                      "nb_setup_ms_per_step": 1e3 * sysm.timings["NB Set Up"] / (2 * steps + 1), "nb_evaluation_ms_per_step": 1e3 * sysm.timings["NB Evaluation"] / (2 * steps + 1)}
    out["note"] = ("synthetic random-walk trajectory, NB term only, host arrays every step; _optimistic: NBModelABFS(optimisticUpdates=True), the update decision "
                   "read with the results of the energy call (one host wait per step instead of two); for scale: the reference spends 0.587 s (serial) / 0.122 s "
                   "(8 OpenMP threads) per NB evaluation and 0.68 s per list update on this system (benchmarks/log/systemBenchmarks_*_1ps.log)")
    return out


def md_device_block(torch, local, steps=1000):
    """BASELINE.json config 4 as real dynamics: 1000 velocity-Verlet steps (dt = 1 fs, 300 K start) on a JAC-size bond-free ionic fluid
    (23 520 atoms; the NB term is its whole force field), everything resident on the device (pdynamo-mirror_b200/md.py); per step the
    host reads 7 scalars.  Two list policies: the reference's displacement heuristic and a forced update every 10 steps."""
    import pdynamo_mirror_b200 as p
    w = make_workload("ionic23k")
    out = {"workload": workload_description("ionic23k", w) + " (LJ + Coulomb fluid, no bonds)"}
    for label, freq in (("displacement_triggered", 0), ("update_every_10", 10)):
        sysm = p.System.FromWorkload(w)
        sysm.DefineNBModel(p.NBModelABFS(device=local))
        md = p.md.VelocityVerletDynamics(sysm, timeStep=0.001, temperature=300.0, device=local)
        md.Run(20, updateFrequency=freq)
        e0, u0 = md.potential + md.kinetic, md.updates
        # two consecutive runs of `steps` steps, the faster one is reported (a loop with a hundred list updates waits for the host a few hundred
        # times: a burst of other activity on the box's CPUs shows up as a factor of two on one run and not on the next)
        wall, traj = None, []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            part = md.Run(steps, updateFrequency=freq)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            wall = dt if wall is None else min(wall, dt)
            traj += part
        tot = np.array([a + b for a, b in traj]); kin = np.array([b for _, b in traj])
        out[label] = {"steps_per_s": steps / wall, "ms_per_step": 1e3 * wall / steps, "steps": steps, "runs": "best of 2 consecutive runs; drift over both", "list_updates": (md.updates - u0) // 2,
                      "total_energy_drift_over_kinetic": float(np.abs(tot - e0).max() / kin.mean()), "temperature_K": float(2.0 * kin.mean() / (md.degreesOfFreedom * 8.314472e-3))}
    return out


def md_dhfr_block(torch, local, steps=1000):
    """The reference's own system benchmark (benchmarks/SystemBenchmarks.py: DHFR, CHARMM22, NBModelABFS defaults; one energy + gradient
    evaluation, then 1000 steps of Langevin velocity Verlet dynamics, 1 fs, 300 K, collision frequency 25 ps^-1) with the complete energy
    model -- five bonded terms + the NB model -- and the integrator on the device.  Published wall times of the reference for exactly
    this protocol: benchmarks/log/systemBenchmarks_{Serial,OMP4,OMP8}_1ps.log (hardware unspecified, ca. 2013)."""
    import pdynamo_mirror_b200 as p
    w = make_workload("dhfr_mm")
    sysm = p.System.FromWorkload(w)
    sysm.DefineNBModel(p.NBModelABFS(device=local))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    md = p.md.LangevinDynamics(sysm, timeStep=0.001, temperature=300.0, collisionFrequency=25.0, device=local)     # includes the initial Energy(doGradients=True)
    e0 = md.potential
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    traj = md.Run(steps)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    wall, loop = t2 - t0, t2 - t1
    updates = int(md.updates)
    md.Run(steps)                                            # the loop rate: the faster of two consecutive runs (see md_device_block); everything else from the first
    torch.cuda.synchronize()
    loop = min(loop, time.perf_counter() - t2)
    kin = np.array([k for _, k in traj]); pot = np.array([q for q, _ in traj])
    temp = 2.0 * kin / (md.degreesOfFreedom * 8.314472e-3)
    return {"workload": workload_description("dhfr_mm", w) + " + 23592 bonds, 11584 angles, 2117 Urey-Bradley, 7000 dihedral, 418 improper terms",
            "protocol": "Energy(doGradients) + %d Langevin velocity-Verlet steps (1 fs, 300 K, 25 ps^-1), displacement-triggered list updates" % steps,
            "wall_s": wall, "wall_note": "state creation (device allocations, first list build), the initial energy call and the %d steps -- what the reference's 'Total' covers" % steps,
            "steps_per_s": steps / loop, "ms_per_step": 1e3 * loop / steps, "loop_note": "loop rate: best of 2 consecutive runs of %d steps; wall_s and the statistics: the first" % steps,
            "list_updates": updates,
            "potential_energy_t0": e0, "potential_energy_published_t0": float(w["published_total"][0]),
            "potential_energy_1ps": float(pot[-1]), "potential_energy_mean": float(pot[100:].mean()), "temperature_mean_K": float(temp[100:].mean()),
            "reference_published": {"serial_wall_s": 656.276, "omp4_wall_s": 269.0, "omp8_wall_s": 201.5, "potential_energy_1ps": -291914.13095708,
                                    "potential_energy_mean": -293191.47148015, "temperature_mean_K": 293.73476041, "list_updates": 73,
                                    "source": "benchmarks/log/systemBenchmarks_Serial_1ps.log:405-465 (different random numbers: compare statistics, not trajectories)"},
            "speedup_vs_published_serial": 656.276 / wall, "speedup_vs_published_omp8": 201.5 / wall}


def md_dhfr_reference_estimate():
    """What one MD step of the same DHFR system costs the reference's own C code on THIS box's host cores (compiled reference, OpenMP build
    where available): NB energy + gradients on existing lists, a list update (amortised over the 13.7 calls per update of the reference's run),
    the five bonded terms.  An estimate of the reference's step rate to put next to the published wall times (unknown 2013 hardware)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refnb
    w = make_workload("dhfr_mm")
    omp = refnb.available(omp=True)
    if not refnb.available(omp=omp):
        return {"unavailable": "oracle/_ref not built"}
    model = refnb.RefNB(w, omp=omp)
    first = model.energy(force_new=True)
    t_e, t_u = [], [first["t_update"]]
    for _ in range(3):
        t_e.append(model.energy(force_new=False)["t_energy"])
    t_u.append(model.energy(force_new=True)["t_update"])
    t0 = time.perf_counter()
    for _ in range(3):
        refnb.mm_energy(w["bonded"], w["xyz"])
    t_b = (time.perf_counter() - t0) / 3.0
    per_step = min(t_e) + min(t_u) / 13.7 + t_b
    return {"cores": model.num_threads() if omp else 1, "nb_energy_s": float(min(t_e)), "list_update_s": float(min(t_u)), "bonded_terms_s": float(t_b),
            "estimated_s_per_step": float(per_step), "estimated_steps_per_s": float(1.0 / per_step),
            "note": "compiled reference C on this box; list update amortised over 13.7 calls (the reference's own DHFR statistics); Python integrator overhead of the reference not included"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="m1")
    ap.add_argument("--ref-sample", default=None, help="bounded CPU sample of the workload for the reference / cpu_baseline legs "
                                                       "(default: water4x4x4, 41 472 atoms of the same replicated water box, for m1; else the workload itself)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-jac", action="store_true")
    args = ap.parse_args()
    args.ref_sample_explicit = args.ref_sample
    if args.ref_sample is None:
        args.ref_sample = "water4x4x4" if args.workload == "m1" else args.workload
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
