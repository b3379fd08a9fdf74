#!/usr/bin/env python
r"""Benchmark of the CI-Hamiltonian hot path (BASELINE.json: "sparse_op nnz/sec build; SpMV HBM GB/s;
time-to-E0 at 1/2/4/8 B200 vs CPU").

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code, same metric

One *step* = one construction of the sparse CI operator for the workload: determinant hash index, count,
scan, fill + per-row sort -- everything ``pyci.sparse_op(ham, wfn)`` does.  The headline ``value`` is
stored non-zeros per second in the REFERENCE's storage format (``op.size``: lower triangle + diagonal for
the default symmetric call), with integrals and determinants already resident in HBM; ``e2e`` is the same
through the public ``pyci_b200.sparse_op`` call from host arrays (host->device upload of integrals and
determinants and a device->host read of the row pointer inside the timed region).  The other two parts of
the metric ride along in the same JSON line: ``spmv`` (achieved HBM GB/s of the fp64 CSR SpMV kernel,
per-launch CUDA events) and ``time_to_e0`` (construction + Davidson solve).

Workloads (BASELINE.json configs): N=1 -> config 3 (FullCI 14 orbitals 4a4b, synthetic integrals), the
largest single-GPU configuration; N>1 -> config 4 (FullCI 16 orbitals 4a4b) row-sharded over the ranks.
``--workload`` selects any of cfg1..cfg4 or synNN explicitly.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 1234
E_TOL = 1.0e-9  # residual tolerance of the solve (relative to |theta|): eigenvalue error << 1e-10 Eh


# ------------------------------------------------------------------------------------------------
# workloads


def datafile(name):
    import gzip
    import shutil
    import tempfile
    plain = os.path.join(ROOT, "tests", "data", name + ".fcidump")
    if os.path.exists(plain):
        return plain
    out = os.path.join(tempfile.gettempdir(), "pyci_b200_%s_%d.fcidump" % (name, os.getpid()))
    if not os.path.exists(out):
        with gzip.open(plain + ".gz", "rb") as src, open(out, "wb") as dst:
            shutil.copyfileobj(src, dst)
    return out


def workload_spec(name):
    """name -> dict(kind, nbasis, occ, source, label)."""
    if name == "cfg1":
        return dict(kind="fullci", occ=(2, 2), file="be_ccpvdz", label="cfg1: Be cc-pVDZ FullCI(2,2), test FCIDUMP")
    if name == "cfg2":
        return dict(kind="doci", occ=(5, 5), file="h2o_ccpvdz", label="cfg2: H2O cc-pVDZ DOCI(5,5), test FCIDUMP")
    if name == "cfg3":
        return dict(kind="fullci", occ=(4, 4), n=14, label="cfg3: FullCI 14 orbitals 4a4b, synthetic integrals seed %d" % SEED)
    if name == "cfg4":
        return dict(kind="fullci", occ=(4, 4), n=16, label="cfg4: FullCI 16 orbitals 4a4b, synthetic integrals seed %d" % SEED)
    if name == "cfg5":
        # "K,P,ndet": seniority-zero selection of P pairs in K spatial orbitals as a GenCI space over 2K
        # spin-orbitals (pyci_b200/synthetic.py); default = 50 M determinants, ~220 stored entries per row
        K, Pn, nd = (int(v) for v in os.environ.get("PYCI_B200_CFG5", "32,10,50000000").split(","))
        return dict(kind="genci", occ=(2 * Pn, 0), n=2 * K, cfg5=(K, Pn, nd),
                    label="cfg5: GenCI %d spin-orbitals %d electrons, %d selected (seniority-zero) determinants, "
                          "synthetic integrals seed %d" % (2 * K, 2 * Pn, nd, SEED))
    if name.startswith("syn"):
        n = int(name[3:])
        return dict(kind="fullci", occ=(4, 4), n=n, label="FullCI %d orbitals 4a4b, synthetic integrals seed %d" % (n, SEED))
    raise SystemExit("unknown workload %r" % name)


def _synthetic():
    """pyci_b200/synthetic.py loaded by path (pure numpy), so that the reference arm never imports the
    product package or its native libraries."""
    import importlib.util
    sp = importlib.util.spec_from_file_location("_pyci_b200_synthetic", os.path.join(ROOT, "pyci_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    return mod


def make_problem(pyci, spec):
    """(ham, wfn) through the public API of `pyci` (this repo's module or the compiled reference)."""
    if "cfg5" in spec:
        K, Pn, nd = spec["cfg5"]
        syn = _synthetic()
        _, one, two = syn.synthetic_integrals(K, SEED)
        h_so, g_so = syn.spin_orbital_integrals(one, two)
        ham = pyci.secondquant_op(0.0, h_so, g_so)
        wfn = pyci.genci_wfn(2 * K, 2 * Pn, 0, syn.seniority_zero_genci_dets(K, Pn, nd))
        return ham, wfn
    if "file" in spec:
        ham = pyci.secondquant_op(datafile(spec["file"]))
    else:
        ham = pyci.secondquant_op(*_synthetic().synthetic_integrals(spec["n"], SEED))
    wfn = getattr(pyci, spec["kind"] + "_wfn")(ham.nbasis, *spec["occ"])
    wfn.add_all_dets()
    return ham, wfn


# ------------------------------------------------------------------------------------------------
# clocks


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C++ (oracle/_ref) or, if absent, the C oracle port


def load_reference(genci=False):
    """(module, kind): oracle/_ref/pyci_ref = the reference's unmodified sources compiled here by
    oracle/Makefile ("reference"); else None and the caller uses the oracle port.  GenCI uses
    pyci_ref_gencifix (two loop bounds of sparseop.cpp:453,476 corrected; the stock GenCI kernels are defective)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    # one build per process: both register the same pybind11 types.  The GenCI-fixed build differs from the stock
    # one only in the two GenCI loop bounds of sparseop.cpp, so it also serves every DOCI / FullCI call.
    name = "pyci_ref_gencifix" if (genci or "pyci_ref_gencifix" in sys.modules) else "pyci_ref"
    if os.path.isdir(os.path.join(ref_dir, name)):
        sys.path.insert(0, ref_dir)
        try:
            import importlib
            return importlib.import_module(name), "reference"
        except ImportError:
            pass
    return None, "port"


class CpuPath:
    """Times construction of the first k rows x all columns of the workload on the host:
    sparse_op(ham, wfn, nrow=k, ncol=ndet, symmetric=False), the reference's native row-slice call
    (sparseop.cpp:49-71); per-row cost is independent of the slice, so
    nnz/s(reference format) = rows/s * (size of the whole default operator / nrow)."""

    def __init__(self, spec):
        self.spec = spec
        self.mod, self.kind = load_reference(spec["kind"] == "genci")
        if self.mod is not None:
            self.ham, self.wfn = make_problem(self.mod, spec)
            self.ndet = len(self.wfn)
        else:
            from oracle import oracle as O
            self.O = O
            if "file" in spec:
                ecore, one, two = O.read_fcidump(datafile(spec["file"]))
            else:
                ecore, one, two = O.synthetic_integrals(spec["n"], SEED)
            self.okind = {"doci": O.DOCI, "fullci": O.FULLCI, "genci": O.GENCI}[spec["kind"]]
            self.ints = O.senzero_integrals(one, two) if spec["kind"] == "doci" else (one, two)
            self.nbasis = one.shape[0]
            self.dets = O.all_dets(self.okind, self.nbasis, *spec["occ"])
            self.ndet = self.dets.shape[0]

    def build_rows(self, k):
        """seconds, stored entries of the k-row slice, and SpMV seconds on it"""
        k = min(k, self.ndet)
        x = np.random.default_rng(0).standard_normal(self.ndet)
        if self.mod is not None:
            t0 = time.perf_counter()
            op = self.mod.sparse_op(self.ham, self.wfn, nrow=k, ncol=self.ndet, symmetric=False)
            t = time.perf_counter() - t0
            nnz = int(op.size)
            y = np.empty(k)
            op(x, out=y)
            t1 = time.perf_counter()
            for _ in range(3):
                op(x, out=y)
            ts = (time.perf_counter() - t1) / 3
        else:
            O = self.O
            t0 = time.perf_counter()
            ip, ix, dv = O.sparse_op(self.okind, self.nbasis, self.spec["occ"][0], self.spec["occ"][1], self.dets,
                                     self.ints, nrow=k, ncol=self.ndet, symmetric=False)
            t = time.perf_counter() - t0
            nnz = len(ix)
            t1 = time.perf_counter()
            for _ in range(3):
                O.matvec(ip, ix, dv, x, False)
            ts = (time.perf_counter() - t1) / 3
        return t, nnz, ts

    def rows_for(self, seconds):
        probe = min(self.ndet, 64)
        t, _, _ = self.build_rows(probe)
        return int(max(probe, min(self.ndet, probe * seconds / max(t, 1e-6))))


def ref_size_per_row(spec, ndet):
    """Stored entries per row of the default (symmetric, lower-triangular) operator of a complete space."""
    from math import comb
    if "cfg5" in spec:
        return None  # selected space: estimated from the sampled rows
    n = spec.get("n")
    if n is None:
        n = {"be_ccpvdz": 14, "h2o_ccpvdz": 24}[spec["file"]]
    a, b = spec["occ"]
    if spec["kind"] == "doci":
        off = a * (n - a)
    else:
        va, vb = n - a, n - b
        off = a * va + b * vb + comb(a, 2) * comb(va, 2) + comb(b, 2) * comb(vb, 2) + a * va * b * vb
    return off / 2.0 + 1.0


def cpu_sample(spec, seconds):
    cp = CpuPath(spec)
    k = cp.rows_for(seconds)
    t, nnz, ts = cp.build_rows(k)
    per_row = ref_size_per_row(spec, cp.ndet) or ((nnz / k - 1.0) / 2.0 + 1.0)
    return {"value": (k / t) * per_row, "unit": "nnz/s", "cores": 1, "kind": cp.kind,
            "sample": "first %d of %d rows x all columns via sparse_op(nrow=k, symmetric=False), %.1f s; "
                      "rows/s scaled by %.1f stored nnz/row of the default operator; the reference hot path is "
                      "single-threaded (sparseop.cpp:196-199); host has %d cores"
                      % (k, cp.ndet, t, per_row, os.cpu_count() or 1),
            "rows_per_s": k / t, "candidate_nnz_per_s": nnz / t,
            "spmv_gbs": nnz * 16 / ts / 1e9 if ts > 0 else None}


def run_reference(args, spec, rank, world):
    if rank != 0:
        return 0
    cp = CpuPath(spec)
    k = cp.rows_for(4.0)  # ~4 s of host work per step
    per_row = ref_size_per_row(spec, cp.ndet)
    if per_row is None:
        _, nnz0, _ = cp.build_rows(k)
        per_row = (nnz0 / k - 1.0) / 2.0 + 1.0
    for _ in range(args.warmup):
        cp.build_rows(max(64, k // 8))
    t_total, spmv_t, nnz_total = 0.0, 0.0, 0
    for _ in range(args.steps):
        t, nnz, ts = cp.build_rows(k)
        t_total += t
        spmv_t += ts
        nnz_total += nnz
    value = (k * args.steps / t_total) * per_row
    line = {
        "impl": "reference", "metric": "sparse_op_build_nnz_per_s", "value": value, "unit": "nnz/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic" if "n" in spec else "reference test FCIDUMP",
        "config": {"workload": spec["label"], "ndet": cp.ndet, "rows_per_step": k},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": 1, "kind": cp.kind,
                         "sample": "each step builds the first %d of %d rows x all columns with the reference's "
                                   "sparse_op(nrow=k, symmetric=False); rows/s scaled by %.1f stored nnz/row of the "
                                   "default operator; single-threaded by construction (sparseop.cpp:196-199), host "
                                   "has %d cores" % (k, cp.ndet, per_row, os.cpu_count() or 1)},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "spmv": {"gbs": nnz_total * 16 / spmv_t / 1e9 if spmv_t > 0 else None, "bytes_per_nnz": 16,
                 "note": "row-slice CSR product through the same compiled code"},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's arm


def run_b200(args, spec, rank, world, local):
    import torch

    import pyci_b200 as pyci
    from pyci_b200 import cabi
    from pyci_b200.distributed import exchange_unique_id

    if not torch.cuda.is_available() or pyci.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the pyci_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # the library launches on torch's current stream so that torch.cuda.Event brackets its kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    pyci.set_device(local, stream.cuda_stream)
    if world > 1:
        pyci.init_comm(rank, world, exchange_unique_id(pyci.nccl_unique_id, rank, world))
    ctx = cabi.Context(local, stream.cuda_stream)
    if world > 1:
        ctx.init_comm(rank, world, exchange_unique_id(cabi.nccl_unique_id, rank, world))

    # ---- host inputs (numpy arrays / host wave function), built once, untimed
    note("building host inputs: " + spec["label"])
    ham, wfn = make_problem(pyci, spec)
    note("host inputs ready: %d determinants" % len(wfn))
    ndet = len(wfn)
    kind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI, "genci": cabi.GENCI}[spec["kind"]]
    dets = wfn.to_det_array()
    h2d_bytes = dets.nbytes + ham.one_mo.nbytes + ham.two_mo.nbytes + ham.h.nbytes + ham.v.nbytes + ham.w.nbytes

    # ---- device-resident inputs for the kernel-side number
    dham = cabi.Ham(ctx, ham.nbasis, ham.ecore, ham.one_mo, ham.two_mo, ham.h, ham.v, ham.w)
    dwfn = cabi.Wfn(ctx, kind, ham.nbasis, wfn.nocc_up, wfn.nocc_dn, dets)

    state = {"op": None}

    def step_device():
        if state["op"] is not None:
            state["op"].close()
        dwfn.reindex()
        state["op"] = cabi.Op(ctx, dham, dwfn)

    for _ in range(args.warmup):
        step_device()
    note("warm-up done, timing %d constructions" % args.steps)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    per_step = []
    for _ in range(args.steps):
        step_device()
        bt = state["op"].build_times()
        per_step.append(dwfn_index_seconds(cabi, dwfn) + bt["total"])
    e1.record(stream)
    barrier()
    launches = ctx.launches
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    op = state["op"]
    size_total = sum_over_ranks(op.size)
    stored_total = sum_over_ranks(op.stored_nnz)
    value = size_total * args.steps / (dev_ms * 1e-3)
    bt = op.build_times()
    fill_s = max_over_ranks(bt["fill_sort"])
    kernel_s = max_over_ranks(float(np.mean(per_step)))

    # ---- SpMV: per-launch CUDA events inside the library, on the same stream
    note("SpMV timing")
    reps = max(args.steps, 10)
    ms = op.time_spmv(max(args.warmup, 3), reps, 0)
    spmv_ms = max_over_ranks(float(np.mean(ms)))
    spmv_bytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
    spmv_bytes_total = sum_over_ranks(spmv_bytes)
    clocks = sampler.summary()

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9  # per GPU
    fill_bytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * (16 if kind == cabi.FULLCI else 8)
    fill_gbs = fill_bytes / max(fill_s, 1e-9) / 1e9
    fill_name = op.fill_kernel()
    op_rows = op.row_count
    spmv_name = "spmv_rows" if spmv_bytes / max(op_rows, 1) > 12 * 320 else "spmv_short_rows"
    traffic = {}
    try:  # dram bytes per launch from the committed ncu --set full captures (single-GPU workloads only)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(spec.get("key", "") if world == 1 else "", {})
    except (OSError, ValueError):
        pass

    # ---- time to E0: one more construction + the Davidson solve, device-timed
    note("time to E0")
    barrier()
    t0 = time.perf_counter()
    step_device()
    op = state["op"]
    bt2 = op.build_times()
    evals, evecs, st = op.solve(n=1, tol=E_TOL)
    torch.cuda.synchronize()
    tte_wall = max_over_ranks(time.perf_counter() - t0)
    tte_dev = max_over_ranks(dwfn_index_seconds(cabi, dwfn) + bt2["total"] + st["seconds"])
    op.close()
    state["op"] = None

    # ---- 1- and 2-RDM of the ground state through the public API (collective when row-sharded), checked by
    # the energy identity of the reference's test_compute_rdms (test_routines.py:115-133)
    note("RDMs")
    barrier()
    t0 = time.perf_counter()
    d1, d2 = pyci.compute_rdms(wfn, evecs[0])
    torch.cuda.synchronize()
    rdm_wall = max_over_ranks(time.perf_counter() - t0)
    rdm_err = abs(rdm_energy(pyci, ham, spec, wfn, d1, d2) - float(evals[0])) if rank == 0 else 0.0
    del evecs, d1, d2

    # ---- selected CI on a thinned FullCI space: one heat-bath iteration (add_hci) and the ENPT2 energy, device
    # seconds of the walk + merge; the reference's own routines timed beside them on a row sample (rank 0)
    barrier()
    note("selected-CI leg")
    sel = selected_ci_leg(cabi, ctx, rank, world, not args.no_cpu_baseline, spec["kind"] == "genci")

    # ---- end to end through the public API, host buffers in, row pointer out
    d2h_bytes = 0

    split = [0.0, 0.0]

    def step_e2e():
        nonlocal d2h_bytes
        ta = time.perf_counter()
        o = pyci.sparse_op(ham, wfn)
        tb = time.perf_counter()
        ip = o.indptr()
        tc = time.perf_counter()
        split[0] += tb - ta
        split[1] += tc - tb
        d2h_bytes = ip.nbytes
        return o.size, int(ip[-1])

    note("end-to-end leg")
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    split[0] = split[1] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sz, _ = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = sum_over_ranks(sz) * args.steps / e2e_s

    line = {
        "metric": "sparse_op_build_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic" if "n" in spec else "reference test FCIDUMP",
        "config": {"workload": spec["label"], "ndet": ndet, "nnz_reference_format": int(size_total),
                   "nnz_streamed_full_rows": int(stored_total), "parallelism": "row-shard x%d" % world,
                   "l2": "operator (%.1f GB/GPU) is larger than the 126 MB L2: no flush between iterations"
                         % (spmv_bytes / 1e9) if spmv_bytes > 4 * 126e6 else "operator fits L2: numbers are L2-resident"},
        "e2e": {"value": e2e_value, "unit": "nnz/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * e2e_s / args.steps,
                "ms_sparse_op": 1e3 * split[0] / args.steps, "ms_indptr": 1e3 * split[1] / args.steps,
                "call": "pyci_b200.sparse_op(ham, wfn); op.indptr()  (host arrays in, pageable)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # the dominant kernel of the timed step (one construction) is the fill kernel; the SpMV kernel that the
        # solve spends its time in is reported beside it
        "roofline": {"kernel": fill_name, "bound": "hbm", "achieved": fill_gbs, "peak": peak, "unit": "GB/s",
                     "frac": fill_gbs / peak, "traffic": traffic.get(fill_name, {}).get("bytes"), "peak_source": peak_src,
                     "bytes_per_launch": int(fill_bytes), "ms_per_launch": 1e3 * fill_s,
                     "share_of_step": fill_s / max(dev_ms * 1e-3 / args.steps, 1e-12),
                     "note": "bytes = CSR written once (12 B per stored non-zero + row pointer) + determinants read once; "
                             "CUDA events around the kernel launches inside the library, on the bench stream"},
        "roofline_spmv": {"kernel": spmv_name, "bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s",
                          "frac": spmv_gbs / peak, "traffic": traffic.get(spmv_name, {}).get("bytes"), "peak_source": peak_src,
                          "bytes_per_launch": int(spmv_bytes), "ms_per_launch": spmv_ms},
        "spmv": {"gbs_per_gpu": spmv_gbs, "gbs_total": spmv_bytes_total / (spmv_ms * 1e-3) / 1e9, "ms": spmv_ms,
                 "bytes_per_nnz": 12, "frac_of_peak": spmv_gbs / peak},
        "build": {"kernel_seconds_per_step": kernel_s, "index_s": dwfn_index_seconds(cabi, dwfn),
                  "count_scan_s": bt["count_scan"], "fill_sort_s": bt["fill_sort"],
                  "full_nnz_per_s": stored_total / max(kernel_s, 1e-9)},
        "time_to_e0": {"seconds_device": tte_dev, "seconds_wall": tte_wall, "E0": float(evals[0]), "matvecs": st["matvecs"],
                       "residual": st["residual"], "tol": E_TOL, "solve_seconds": st["seconds"],
                       "spmv_seconds": st["spmv_seconds"]},
        "rdm": {"seconds_wall": rdm_wall, "energy_identity_abs_error": rdm_err,
                "call": "pyci_b200.compute_rdms(wfn, c0): wfn upload + index + contraction + tensors back"},
        "selected_ci": sel,
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        note("CPU baseline (reference, host cores)")
        line["cpu_baseline"] = cpu_sample(spec, args.cpu_seconds)
    note("done")
    if rank == 0:
        emit(line)
    dwfn.close()
    dham.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


HCI_N, HCI_OCC, HCI_STRIDE, HCI_EPS, HCI_EPS_UPDATE = 16, (4, 4), 33, 2.0e-4, 2.0e-2


def selected_ci_leg(cabi, ctx, rank, world, with_cpu, genci_ref):
    """add_hci + compute_enpt2 on every 33rd determinant of FullCI(16, 4a4b) with a seeded
    coefficient vector: device seconds (max over ranks is taken by the collective itself: every rank ends with the
    merged result), and the compiled reference on the first rows of the same space, single-threaded like its build."""
    syn = _synthetic()
    _, one, two = syn.synthetic_integrals(HCI_N, SEED)
    import pyci_b200 as pyci
    full = pyci.fullci_wfn(HCI_N, *HCI_OCC)
    full.add_all_dets()
    dets = np.ascontiguousarray(full.to_det_array()[::HCI_STRIDE])
    del full
    c = np.random.default_rng(1).standard_normal(len(dets))
    c /= np.linalg.norm(c)
    ham = cabi.Ham(ctx, HCI_N, 0.0, one, two)
    out = {"workload": "FullCI(%d, %da%db), every %dth determinant, eps %g" % (HCI_N, HCI_OCC[0], HCI_OCC[1], HCI_STRIDE, HCI_EPS),
           "ndet": int(len(dets))}
    for rep in range(2):  # second pass: warm allocations
        wfn = cabi.Wfn(ctx, cabi.FULLCI, HCI_N, HCI_OCC[0], HCI_OCC[1], dets)
        pt, nt = wfn.compute_enpt2(ham, c, -10.0, HCI_EPS)
        out["enpt2_seconds_device"], out["external_determinants"], out["enpt2"] = wfn.ext_seconds(), int(nt), pt
        new = wfn.add_hci(ham, c, HCI_EPS)
        out["add_hci_seconds_device"], out["added"] = wfn.ext_seconds(), int(len(new))
        wfn.close()
    if world == 1:  # the incremental path is single-rank (row blocks move when the operator grows)
        # SparseOp::update after a selection step that adds ~50 % more determinants: incremental growth (only the new
        # determinants are enumerated) against a fresh construction of the grown operator
        wfn = cabi.Wfn(ctx, cabi.FULLCI, HCI_N, HCI_OCC[0], HCI_OCC[1], dets)
        op = cabi.Op(ctx, ham, wfn)
        grown = wfn.add_hci(ham, c, HCI_EPS_UPDATE)
        for rep in range(2):
            if rep:  # second pass on a fresh copy of the small operator: warm allocations
                op.close()
                small = cabi.Wfn(ctx, cabi.FULLCI, HCI_N, HCI_OCC[0], HCI_OCC[1], dets)
                op = cabi.Op(ctx, ham, small)
                small.close()
            op.update(ham, wfn)
            upd = op.build_times()["total"]
            fresh = cabi.Op(ctx, ham, wfn)
            full = fresh.build_times()["total"]
            same = (op.size == fresh.size) and (op.stored_nnz == fresh.stored_nnz)
            fresh.close()
        out["update"] = {"eps": HCI_EPS_UPDATE, "ndet_before": int(len(dets)), "ndet_after": int(wfn.ndet),
                         "added": int(len(grown)), "stored_nnz_after": int(op.stored_nnz),
                         "update_seconds_device": upd, "fresh_build_seconds_device": full, "same_size_as_fresh_build": bool(same)}
        op.close()
        wfn.close()
    ham.close()
    na, nv = HCI_OCC[0], HCI_N - HCI_OCC[0]
    cand = 2 * na * nv + 2 * (na * (na - 1) // 2) * (nv * (nv - 1) // 2) + (na * nv) ** 2
    out["rows_per_s"] = len(dets) / out["add_hci_seconds_device"]
    out["candidates_per_s"] = cand * len(dets) / out["add_hci_seconds_device"]
    if with_cpu and rank == 0 and world == 1:
        ref, kind = load_reference(genci_ref)
        if ref is not None:
            k = 3000
            rham = ref.secondquant_op(0.0, one, two)
            rw = ref.fullci_wfn(HCI_N, HCI_OCC[0], HCI_OCC[1], dets[:k])
            t0 = time.perf_counter()
            ref.compute_enpt2(rham, rw, c[:k], -10.0, HCI_EPS, 1)
            t1 = time.perf_counter()
            ref.add_hci(rham, rw, c[:k], HCI_EPS, 1)
            t2 = time.perf_counter()
            out["cpu_reference"] = {"kind": kind, "cores": 1, "sample": "first %d determinants" % k,
                                    "enpt2_rows_per_s": k / (t1 - t0), "add_hci_rows_per_s": k / (t2 - t1)}
    return out


def rdm_energy(pyci, ham, spec, wfn, d1, d2):
    """E = ecore + sum h gamma + 1/4 sum <pq||rs> Gamma in the spin-orbital basis (test_routines.py:130-133)."""
    if spec["kind"] == "genci":
        h2, g2, r1, r2 = ham.one_mo, ham.two_mo, d1, d2
    else:
        h2, g2 = _synthetic().spin_orbital_integrals(ham.one_mo, ham.two_mo)
        r1, r2 = pyci.spinize_rdms(d1, d2)
    e2 = np.einsum("ijkl,ijkl", g2, r2) - np.einsum("ijlk,ijkl", g2, r2)
    return ham.ecore + np.einsum("ij,ij", h2, r1) + 0.25 * e2


_JSON_OUT = None
_T0 = time.perf_counter()


def note(msg):
    """progress line on stderr (rank 0), with seconds since start"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench %7.1f s] %s\n" % (time.perf_counter() - _T0, msg))
        sys.stderr.flush()



def claim_stdout():
    """The driver reads ONE JSON line from stdout: keep the real stdout for it and send everything else that
    writes to file descriptor 1 (NCCL's version banner, library chatter) to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def dwfn_index_seconds(cabi, dwfn):
    return cabi.lib().pyci_wfn_index_seconds(dwfn.handle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    if name == "auto":
        name = "cfg3" if max(world, args.gpus) == 1 else "cfg4"
    spec = workload_spec(name)
    spec["key"] = name
    if args.impl == "reference":
        return run_reference(args, spec, rank, world)
    if world != args.gpus and rank == 0:
        print("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)" % (args.gpus, world),
              file=sys.stderr)
    return run_b200(args, spec, rank, world, local)


if __name__ == "__main__":
    sys.exit(main())
