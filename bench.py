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


def _ref_worker(job):
    """One process of the reference arm: the compiled reference's sparse_op on a k-row slice, `steps` times, started
    together with its siblings (barrier) so that the processes load the host at the same time."""
    spec, k, steps, warmup, barrier = job
    cp = CpuPath(spec)
    for _ in range(warmup):
        cp.build_rows(max(64, k // 8))
    barrier.wait()
    t_total, spmv_t, nnz_total = 0.0, 0.0, 0
    for _ in range(steps):
        t, nnz, ts = cp.build_rows(k)
        t_total += t
        spmv_t += ts
        nnz_total += nnz
    return t_total, spmv_t, nnz_total, cp.kind, cp.ndet


def run_reference(args, spec, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores.  Its row loop is serial
    (sparseop.cpp:196-199, no OpenMP in its Makefile), so "all the host threads it can use" is one per process:
    `cores` processes (one per host core, at most 32) each build a k-row slice through the reference's public
    sparse_op(ham, wfn, nrow=k, symmetric=False) at the same time; value = summed rows/s x stored nnz per row of the
    default operator.  The single-process rate is reported beside it."""
    if rank != 0:
        return 0
    import multiprocessing as mp
    cp = CpuPath(spec)
    k = cp.rows_for(4.0)  # ~4 s of host work per step and process
    per_row = ref_size_per_row(spec, cp.ndet)
    if per_row is None:
        _, nnz0, _ = cp.build_rows(k)
        per_row = (nnz0 / k - 1.0) / 2.0 + 1.0
    t1, _, _ = cp.build_rows(k)
    single = (k / t1) * per_row
    ndet, kind = cp.ndet, cp.kind
    del cp
    cores = max(1, min(os.cpu_count() or 1, 32, args.ref_procs if args.ref_procs > 0 else 1 << 30))
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        barrier = mgr.Barrier(cores)
        with ctx.Pool(cores) as pool:
            res = pool.map(_ref_worker, [(spec, k, args.steps, args.warmup, barrier)] * cores, chunksize=1)
    rows_per_s = sum(k * args.steps / r[0] for r in res)
    value = rows_per_s * per_row
    spmv_t = sum(r[1] for r in res)
    nnz_total = sum(r[2] for r in res)
    line = {
        "impl": "reference", "metric": "sparse_op_build_nnz_per_s", "value": value, "unit": "nnz/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r[0] for r in res])) / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic" if "n" in spec else "reference test FCIDUMP",
        "config": {"workload": spec["label"], "ndet": ndet, "rows_per_step": k, "processes": cores},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": cores, "kind": kind,
                         "single_core_value": single,
                         "sample": "%d processes (one per host core, <= 32; the reference's row loop is serial, "
                                   "sparseop.cpp:196-199) each build the first %d of %d rows x all columns with the "
                                   "reference's sparse_op(nrow=k, symmetric=False) per step, at the same time; summed "
                                   "rows/s scaled by %.1f stored nnz/row of the default operator; host has %d cores"
                                   % (cores, k, ndet, per_row, os.cpu_count() or 1)},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "spmv": {"gbs": nnz_total * 16 / spmv_t / 1e9 if spmv_t > 0 else None, "bytes_per_nnz": 16,
                 "note": "row-slice CSR product through the same compiled code, per process"},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's arm


class Bench:
    """Process-wide state of the B200 arm: device context, rank plumbing, measured peak."""

    def __init__(self, args, rank, world, local):
        import torch

        import pyci_b200 as pyci
        from pyci_b200 import cabi
        from pyci_b200.distributed import exchange_unique_id
        self.torch, self.pyci, self.cabi, self.args = torch, pyci, cabi, args
        self.rank, self.world, self.local = rank, world, local
        if not torch.cuda.is_available() or pyci.device_count() == 0:
            raise SystemExit("bench.py: no CUDA device; the pyci_b200 path has no CPU fallback")
        torch.cuda.set_device(local)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group(backend="nccl", rank=rank, world_size=world)
            self.dist = dist
        # the library launches on torch's current stream so that torch.cuda.Event brackets its kernels
        self.stream = torch.cuda.Stream(device=local)
        torch.cuda.set_stream(self.stream)
        pyci.set_device(local, self.stream.cuda_stream)
        if world > 1:
            pyci.init_comm(rank, world, exchange_unique_id(pyci.nccl_unique_id, rank, world))
        self.ctx = cabi.Context(local, self.stream.cuda_stream)
        if world > 1:
            self.ctx.init_comm(rank, world, exchange_unique_id(cabi.nccl_unique_id, rank, world))
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = ("measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks
                         else "fallback 6650 GB/s (B200_PROFILING.md)")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def min(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def close(self):
        self.ctx.close()
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def device_inputs(B, spec):
    """(dham, dwfn, kind, host integrals for the checker, h2d bytes of the e2e call) through the C ABI.  Complete
    spaces are unranked on the device (Wfn::add_all_dets); config 5 uploads its explicit determinant array."""
    cabi = B.cabi
    syn = _synthetic()
    kind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI, "genci": cabi.GENCI}[spec["kind"]]
    if "cfg5" in spec:
        K, Pn, nd = spec["cfg5"]
        _, one, two = syn.synthetic_integrals(K, SEED)
        h_so, g_so = syn.spin_orbital_integrals(one, two)
        dets = syn.seniority_zero_genci_dets(K, Pn, nd)
        dham = cabi.Ham(B.ctx, 2 * K, 0.0, h_so, g_so)
        dwfn = cabi.Wfn(B.ctx, kind, 2 * K, 2 * Pn, 0, dets)
        return dham, dwfn, kind, dict(n=2 * K, occ=(2 * Pn, 0), ecore=0.0, one=h_so, two=g_so, dets=dets)
    if "file" in spec:
        ham = B.pyci.secondquant_op(datafile(spec["file"]))
        n, ecore, one, two = ham.nbasis, ham.ecore, ham.one_mo, ham.two_mo
        hvw = (ham.h, ham.v, ham.w)
    else:
        ecore, one, two = syn.synthetic_integrals(spec["n"], SEED)
        n, hvw = spec["n"], (None, None, None)
    dham = cabi.Ham(B.ctx, n, ecore, one, two, *hvw)
    dwfn = cabi.Wfn(B.ctx, kind, n, spec["occ"][0], spec["occ"][1])
    return dham, dwfn, kind, dict(n=n, occ=tuple(spec["occ"]), ecore=ecore, one=one, two=two, hvw=hvw, dets=None)


def gate_rows(row0, nloc, nb, rng, target):
    """Rows the parity gate compares: first / last rows of the shard, both sides of alpha-string boundaries (every
    nb rows; the complete-space fill restages there) and of uniform CTA range boundaries (one and four ranges per SM),
    random rest."""
    if nloc <= 0:
        return np.zeros(0, dtype=np.int64)
    rows = {row0, row0 + 1, row0 + nloc - 1, row0 + nloc - 2}
    bounds = np.zeros(0, dtype=np.int64)
    if nb:
        bounds = np.arange(-(-row0 // nb) * nb, row0 + nloc, nb)
        if len(bounds) > 40:
            bounds = np.concatenate([bounds[:10], bounds[-10:], rng.choice(bounds[10:-10], 20, replace=False)])
    for parts in (148, 592):
        per = -(-nloc // parts)
        cb = row0 + per * np.arange(1, parts)
        cb = cb[cb < row0 + nloc]
        if len(cb):
            bounds = np.concatenate([bounds, rng.choice(cb, min(24, len(cb)), replace=False)])
    for b in bounds:
        rows.update((int(b) - 1, int(b), int(b) + 1))
    rows = {r for r in rows if row0 <= r < row0 + nloc}
    want = min(target, nloc)
    while len(rows) < want:  # random rest (collisions are redrawn)
        rows.update(int(r) for r in rng.integers(row0, row0 + nloc, want - len(rows)))
    return np.array(sorted(rows), dtype=np.int64)


def parity_gate(B, spec, op, dwfn, kind, host, target):
    """BASELINE.md 4.4 gate on the TIMED operator of this rank: sampled rows exported from HBM (pyci_op_export_rows)
    against the CPU oracle's row-list entry (the checker; pinned against the compiled reference by tests/test_oracle.py):
    indptr / indices bit-equal, data <= 1e-12 relative.  Returns per-rank numbers; reduced by the caller."""
    from oracle import oracle as O
    okind = {"doci": O.DOCI, "fullci": O.FULLCI, "genci": O.GENCI}[spec["kind"]]
    dets = host["dets"] if host["dets"] is not None else dwfn.download_dets()
    n, occ = host["n"], host["occ"]
    from math import comb
    nb = comb(n, occ[1]) if spec["kind"] == "fullci" else 0
    rng = np.random.default_rng(1000 + B.rank)
    rows = gate_rows(op.row_begin, op.row_count, nb, rng, target)
    if spec["kind"] == "doci":
        ints = O.senzero_integrals(host["one"], host["two"])
    else:
        ints = (host["one"], host["two"])
    t0 = time.perf_counter()
    gi, gx, gd = op.export_rows(rows)
    oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, ints, rows=rows)
    same = bool(np.array_equal(gi, oi) and np.array_equal(gx, ox))
    if same and len(od):
        rel = float(np.max(np.abs(gd - od)) / max(np.max(np.abs(od)), 1e-300))
        bit = bool(np.array_equal(gd, od))
    else:
        rel, bit = (0.0, True) if same else (float("inf"), False)
    return {"rows": int(len(rows)), "entries": int(len(ox)), "structure_equal": same, "max_rel": rel,
            "data_bit_identical": bit, "seconds": time.perf_counter() - t0}


def reduce_parity(B, g):
    out = {"rows": int(B.sum(g["rows"])), "entries": int(B.sum(g["entries"])),
           "structure_equal": bool(B.min(1.0 if g["structure_equal"] else 0.0) > 0.5),
           "max_rel": B.max(g["max_rel"]) if np.isfinite(g["max_rel"]) else float("inf"),
           "data_bit_identical": bool(B.min(1.0 if g["data_bit_identical"] else 0.0) > 0.5),
           "ranks": B.world, "checker_seconds": B.max(g["seconds"]),
           "against": "oracle_sparse_op_rows (CPU restatement pinned to the compiled reference) on sampled rows of the "
                      "timed operator: shard ends, alpha-string and CTA-range boundaries, random rest"}
    out["ok"] = bool(out["structure_equal"] and out["max_rel"] <= 1e-12)
    return out


def golden_e0(spec):
    """E0 of the oracle-built operator by ARPACK (tests/golden/e0_syn.json, made by tests/golden/make_golden_e0.py; for
    config 4, whose matrix does not fit a CPU box, by the string-driven direct-CI product of make_golden_e0_direct.py,
    which reproduces the matrix-based goldens of the smaller sizes), or the energies the reference's own tests pin
    (pyci/test/test_routines.py:44-45)."""
    if spec.get("key") == "cfg1":
        return -14.617409507, 1e-9, "pyci/test/test_routines.py:44 (atol 1e-9)"
    if spec.get("key") == "cfg2":
        return -75.634588422, 1e-9, "pyci/test/test_routines.py:45 (atol 1e-9)"
    if "n" in spec and "cfg5" not in spec:
        try:
            with open(os.path.join(ROOT, "tests", "golden", "e0_syn.json")) as f:
                g = json.load(f).get("syn%d" % spec["n"])
            if g and tuple(g["occ"]) == tuple(spec["occ"]):
                how = "string-driven direct CI" if g.get("operator", "").startswith("string-driven") else "oracle-built operator"
                return g["E0"], 1e-10, "tests/golden/e0_syn.json (%s + ARPACK, %d matvecs)" % (how, g["arpack_matvecs"])
        except (OSError, ValueError):
            pass
    return None, None, None


def measure_case(B, spec, steps, warmup, gate_target=1000, rdm=False, clocks=False):
    """One workload through the C ABI with device-resident inputs: timed constructions (CUDA events on the bench
    stream, max over ranks), the parity gate on the timed operator, SpMV per-launch timing, time to E0, optionally
    the RDMs.  Returns (result dict, live objects for the caller's extra legs)."""
    cabi, torch = B.cabi, B.torch
    note("inputs: " + spec["label"])
    dham, dwfn, kind, host = device_inputs(B, spec)
    ndet = dwfn.ndet
    state = {"op": None}

    def step_device():
        if state["op"] is not None:
            state["op"].close()
        dwfn.reindex()
        state["op"] = cabi.Op(B.ctx, dham, dwfn)

    for _ in range(warmup):
        step_device()
    note("warm-up done, timing %d constructions" % steps)
    sampler = ClockSampler(B.local) if clocks else None
    if sampler:
        sampler.start()
    B.barrier()
    B.ctx.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(B.stream)
    per_step, fills = [], []
    for _ in range(steps):
        step_device()
        bt = state["op"].build_times()
        per_step.append(dwfn_index_seconds(cabi, dwfn) + bt["total"])
        fills.append(state["op"].fill_seconds())
    e1.record(B.stream)
    B.barrier()
    launches = B.ctx.launches
    dev_ms = B.max(e0.elapsed_time(e1))
    op = state["op"]
    size_total = B.sum(op.size)
    stored_total = B.sum(op.stored_nnz)
    bt = op.build_times()
    fill_kernel_s = B.max(float(np.mean(fills)))
    kernel_s = B.max(float(np.mean(per_step)))

    # ---- parity gate on the operator that was just timed, every rank
    note("parity gate")
    parity = reduce_parity(B, parity_gate(B, spec, op, dwfn, kind, host, gate_target))

    # ---- SpMV: per-launch CUDA events inside the library, on the same stream
    note("SpMV timing")
    reps = max(steps, 10)
    ms = op.time_spmv(max(warmup, 3), reps, 0)
    spmv_ms = B.max(float(np.mean(ms)))
    spmv_bytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
    spmv_bytes_total = B.sum(spmv_bytes)
    clk = sampler.summary() if sampler else None
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9  # per GPU
    fill_bytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * (16 if kind == cabi.FULLCI else 8)
    fill_gbs = fill_bytes / max(fill_kernel_s, 1e-9) / 1e9
    fill_name, count_name = op.fill_kernel(), op.count_kernel()
    spmv_name = "spmv_rows" if spmv_bytes / max(op.row_count, 1) > 12 * 320 else "spmv_short_rows_seq"
    op_bytes = spmv_bytes

    # ---- time to E0: one more construction + the Davidson solve, device-timed
    note("time to E0")
    B.barrier()
    t0 = time.perf_counter()
    step_device()
    op = state["op"]
    bt2 = op.build_times()
    evals, evecs, st = op.solve(n=1, tol=E_TOL)
    torch.cuda.synchronize()
    tte_wall = B.max(time.perf_counter() - t0)
    tte_dev = B.max(dwfn_index_seconds(cabi, dwfn) + bt2["total"] + st["seconds"])
    op.close()
    state["op"] = None
    gold, gtol, gsrc = golden_e0(spec)
    parity["dE0"] = abs(float(evals[0]) - (gold + (host["ecore"] if "n" in spec else 0.0))) if gold is not None else None
    parity["E0_reference"] = gsrc
    parity["E0_residual"] = st["residual"]
    if gold is not None:
        parity["ok"] = bool(parity["ok"] and parity["dE0"] <= max(gtol, 1e-10))

    out = {
        "workload": spec["label"], "ndet": int(ndet), "ms_per_step": dev_ms / steps, "dev_ms": dev_ms,
        "value": size_total * steps / (dev_ms * 1e-3), "nnz_reference_format": int(size_total),
        "nnz_streamed_full_rows": int(stored_total), "gpu_launches": int(launches), "clocks": clk,
        "parity": parity,
        "roofline": {"kernel": fill_name, "bound": "hbm", "achieved": fill_gbs, "peak": B.peak, "unit": "GB/s",
                     "frac": fill_gbs / B.peak, "traffic": None, "peak_source": B.peak_src,
                     "bytes_per_launch": int(fill_bytes), "ms_per_launch": 1e3 * fill_kernel_s,
                     "share_of_step": fill_kernel_s / max(dev_ms * 1e-3 / steps, 1e-12),
                     "note": "bytes = CSR written once (12 B per stored non-zero + row pointer) + determinants read once; "
                             "duration = CUDA events around the fill kernel's launch alone, inside the library, on the "
                             "bench stream (mean of the timed steps, max over ranks)"},
        "roofline_spmv": {"kernel": spmv_name, "bound": "hbm", "achieved": spmv_gbs, "peak": B.peak, "unit": "GB/s",
                          "frac": spmv_gbs / B.peak, "traffic": None, "peak_source": B.peak_src,
                          "bytes_per_launch": int(spmv_bytes), "ms_per_launch": spmv_ms},
        "spmv": {"gbs_per_gpu": spmv_gbs, "gbs_total": spmv_bytes_total / (spmv_ms * 1e-3) / 1e9, "ms": spmv_ms,
                 "bytes_per_nnz": 12, "frac_of_peak": spmv_gbs / B.peak},
        "build": {"kernel_seconds_per_step": kernel_s, "index_s": dwfn_index_seconds(cabi, dwfn),
                  "count_scan_s": bt["count_scan"], "fill_sort_s": bt["fill_sort"], "fill_kernel_s": fill_kernel_s,
                  "step_minus_fill_ms": dev_ms / steps - 1e3 * fill_kernel_s, "count_kernel": count_name,
                  "full_nnz_per_s": stored_total / max(kernel_s, 1e-9)},
        "time_to_e0": {"seconds_device": tte_dev, "seconds_wall": tte_wall, "E0": float(evals[0]), "matvecs": st["matvecs"],
                       "residual": st["residual"], "tol": E_TOL, "solve_seconds": st["seconds"],
                       "spmv_seconds": st["spmv_seconds"]},
        "operator_bytes_per_gpu": int(op_bytes),
    }
    if rdm:
        note("RDMs")
        B.barrier()
        t0 = time.perf_counter()
        d1, d2 = cabi.compute_rdms(B.ctx, dwfn, kind, host["n"], evecs[0])
        torch.cuda.synchronize()
        out["rdm"] = {"seconds_wall": B.max(time.perf_counter() - t0),
                      "energy_identity_abs_error": abs(rdm_energy_arrays(spec, host, d1, d2) - float(evals[0])),
                      "call": "pyci_compute_rdms(ctx, wfn, c0): coefficients up, contraction, all-reduce, tensors back"}
        del d1, d2
    live = {"dham": dham, "dwfn": dwfn, "kind": kind, "host": host, "evals": evals, "evecs": evecs}
    return out, live


def rdm_energy_arrays(spec, host, d1, d2):
    """E = ecore + sum h gamma + 1/4 sum <pq||rs> Gamma in the spin-orbital basis (test_routines.py:130-133)."""
    if spec["kind"] == "genci":
        h2, g2, r1, r2 = host["one"], host["two"], d1, d2
    else:
        import pyci_b200 as pyci
        h2, g2 = _synthetic().spin_orbital_integrals(host["one"], host["two"])
        r1, r2 = pyci.spinize_rdms(d1, d2)
    e2 = np.einsum("ijkl,ijkl", g2, r2) - np.einsum("ijlk,ijkl", g2, r2)
    return host["ecore"] + np.einsum("ij,ij", h2, r1) + 0.25 * e2


def cpu_time_to_e0(B, names):
    """BASELINE.md 4.3 on the host, beside the same workloads on the GPU: t_build_cpu = the compiled reference's
    sparse_op(ham, wfn) (1 core: its row loop is serial), t_eigsh = scipy eigsh(k=1, which='SA', tol=1e-12, ncv=30) on
    the full symmetric matrix with its matvec count; GPU: construction + Davidson through the C ABI."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    ref, kind = load_reference(False)
    out = []
    for name in names:
        spec = workload_spec(name)
        spec["key"] = name
        row = {"workload": spec["label"]}
        res, live = measure_case(B, spec, 2, 1, gate_target=200)
        row.update(gpu_seconds_to_e0=res["time_to_e0"]["seconds_device"], gpu_matvecs=res["time_to_e0"]["matvecs"],
                   gpu_E0=res["time_to_e0"]["E0"], gpu_build_ms=res["ms_per_step"], parity=res["parity"])
        live["dwfn"].close()
        live["dham"].close()
        if ref is not None:
            ham, wfn = make_problem(ref, spec)
            t0 = time.perf_counter()
            op = ref.sparse_op(ham, wfn)
            t1 = time.perf_counter()
            L = sp.csr_matrix((op.data(), op.indices(), op.indptr()), shape=(len(wfn),) * 2)
            A = (L + sp.tril(L, -1).T).tocsr()
            count = [0]

            def mv(x, A=A, count=count):
                count[0] += 1
                return A @ x

            t2 = time.perf_counter()
            w, _ = spla.eigsh(spla.LinearOperator(A.shape, matvec=mv, dtype=np.float64), k=1, which="SA", tol=1e-12,
                              ncv=min(30, len(wfn) - 1))
            t3 = time.perf_counter()
            row.update(cpu_kind=kind, cpu_cores=1, t_build_cpu=t1 - t0, t_eigsh_cpu=t3 - t2,
                       t_e0_cpu=(t1 - t0) + (t3 - t2), arpack_matvecs=count[0], cpu_E0=float(w[0] + ham.ecore),
                       dE0_gpu_vs_cpu=abs(float(w[0] + ham.ecore) - row["gpu_E0"]))
        out.append(row)
    return out


def run_b200(args, spec, rank, world, local):
    B = Bench(args, rank, world, local)
    cabi, pyci, torch = B.cabi, B.pyci, B.torch
    res, live = measure_case(B, spec, args.steps, args.warmup, gate_target=args.gate_rows, rdm=False, clocks=True)
    ndet = res["ndet"]
    traffic = {}
    try:  # dram bytes per launch from the committed ncu --set full captures (single-GPU workloads only)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(spec.get("key", "") if world == 1 else "", {})
    except (OSError, ValueError):
        pass
    res["roofline"]["traffic"] = traffic.get(res["roofline"]["kernel"], {}).get("bytes")
    res["roofline_spmv"]["traffic"] = traffic.get(res["roofline_spmv"]["kernel"], {}).get("bytes")
    live["dwfn"].close()
    live["dham"].close()
    evals, evecs = live["evals"], live["evecs"]

    # ---- host objects for the public-API legs (built once, untimed)
    note("building host inputs: " + spec["label"])
    ham, wfn = make_problem(pyci, spec)
    full_space = "cfg5" not in spec
    h2d_bytes = ham.one_mo.nbytes + ham.two_mo.nbytes + ham.h.nbytes + ham.v.nbytes + ham.w.nbytes
    if not full_space:
        h2d_bytes += wfn.to_det_array().nbytes

    # ---- 1- and 2-RDM of the ground state through the public API (collective when row-sharded), checked by
    # the energy identity of the reference's test_compute_rdms (test_routines.py:115-133)
    note("RDMs")
    B.barrier()
    t0 = time.perf_counter()
    d1, d2 = pyci.compute_rdms(wfn, evecs[0])
    torch.cuda.synchronize()
    rdm_wall = B.max(time.perf_counter() - t0)
    rdm_err = abs(rdm_energy(pyci, ham, spec, wfn, d1, d2) - float(evals[0])) if rank == 0 else 0.0
    del evecs, d1, d2

    # ---- selected CI on a thinned FullCI space: one heat-bath iteration (add_hci) and the ENPT2 energy, device
    # seconds of the walk + merge; the reference's own routines timed beside them on a row sample (rank 0)
    B.barrier()
    note("selected-CI leg")
    sel = selected_ci_leg(cabi, B.ctx, rank, world, not args.no_cpu_baseline, spec["kind"] == "genci")
    if not args.no_extras:
        try:
            sel["large"] = selected_ci_big_leg(cabi, B.ctx)
        except Exception as exc:  # an extra leg must not take the headline line down; it is reported
            sel["large"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- end to end through the public API, host objects in, row pointer out
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
        return int(ip[-1])  # = op.size: entries of this rank's rows in the reference's storage

    note("end-to-end leg")
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    B.barrier()
    split[0] = split[1] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sz = step_e2e()
    torch.cuda.synchronize()
    e2e_s = B.max(time.perf_counter() - t0)
    e2e_value = B.sum(sz) * args.steps / e2e_s

    line = {
        "metric": "sparse_op_build_nnz_per_s", "value": res["value"], "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic" if "n" in spec else "reference test FCIDUMP",
        "config": {"workload": spec["label"], "ndet": ndet, "nnz_reference_format": res["nnz_reference_format"],
                   "nnz_streamed_full_rows": res["nnz_streamed_full_rows"], "parallelism": "row-shard x%d" % world,
                   "l2": "operator (%.1f GB/GPU) is larger than the 126 MB L2: no flush between iterations"
                         % (res["operator_bytes_per_gpu"] / 1e9) if res["operator_bytes_per_gpu"] > 4 * 126e6
                         else "operator fits L2: numbers are L2-resident"},
        "e2e": {"value": e2e_value, "unit": "nnz/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * e2e_s / args.steps,
                "ms_sparse_op": 1e3 * split[0] / args.steps, "ms_indptr": 1e3 * split[1] / args.steps,
                "call": "pyci_b200.sparse_op(ham, wfn); op.indptr()  (host objects in, pageable; a wave function filled "
                        "by add_all_dets is defined by (nbasis, nocc_up, nocc_dn) and is unranked on the device, so the "
                        "bytes that cross PCIe are the integrals in and the row pointer out)" if full_space else
                        "pyci_b200.sparse_op(ham, wfn); op.indptr()  (host arrays in, pageable)"},
        "gpu_launches": res["gpu_launches"],
        "clocks": res["clocks"],
        "parity": res["parity"],
        # the dominant kernel of the timed step (one construction) is the fill kernel; the SpMV kernel that the
        # solve spends its time in is reported beside it
        "roofline": res["roofline"], "roofline_spmv": res["roofline_spmv"], "spmv": res["spmv"], "build": res["build"],
        "time_to_e0": res["time_to_e0"],
        "rdm": {"seconds_wall": rdm_wall, "energy_identity_abs_error": rdm_err,
                "call": "pyci_b200.compute_rdms(wfn, c0): wfn upload + index + contraction + tensors back"},
        "selected_ci": sel,
    }
    # ---- N = 1: the multi-GPU workload on one GPU (the base of a true strong-scaling curve) and the CPU
    # time-to-E0 legs; N = 8: BASELINE config 5 at full size
    if world == 1 and spec.get("key") == "cfg3" and not args.no_extras:
        _, total_mem = torch.cuda.mem_get_info(local)  # (the stream-ordered pool keeps freed blocks: "free" is no guide)
        if total_mem > 170e9:
            note("scaling base: config 4 on one GPU")
            s4 = workload_spec("cfg4")
            s4["key"] = "cfg4"
            r4, l4 = measure_case(B, s4, max(2, args.steps // 4), 1, gate_target=args.gate_rows)
            l4["dwfn"].close()
            l4["dham"].close()
            line["scaling_base"] = {k: r4[k] for k in ("workload", "ndet", "ms_per_step", "value", "nnz_reference_format",
                                                        "parity", "roofline", "spmv", "build", "time_to_e0")}
        if rank == 0 and not args.no_cpu_baseline:
            note("CPU time-to-E0 legs")
            line["cpu_time_to_e0"] = cpu_time_to_e0(B, ["cfg1", "cfg2", "syn10"])
    # (PYCI_B200_FORCE_CFG5=1 runs the leg at any rank count, with PYCI_B200_CFG5="K,P,ndet" sizing it: 2-GPU rehearsals)
    if (world == 8 or os.environ.get("PYCI_B200_FORCE_CFG5")) and spec.get("key") == "cfg4" and not args.no_extras:
        note("config 5")
        s5 = workload_spec("cfg5")
        s5["key"] = "cfg5"
        # the headline line is complete at this point: a leg that hangs (a rank-local failure leaves the others inside a
        # collective) must not lose it
        guard = leg_watchdog(float(os.environ.get("PYCI_B200_LEG_TIMEOUT", "420")), rank, line, "cfg5", s5["label"])
        try:
            r5, l5 = measure_case(B, s5, 1, 1, gate_target=max(200, args.gate_rows // 4), rdm=True)
            l5["dwfn"].close()
            l5["dham"].close()
            line["cfg5"] = {k: r5[k] for k in ("workload", "ndet", "ms_per_step", "value", "nnz_reference_format",
                                                "nnz_streamed_full_rows", "parity", "roofline", "roofline_spmv", "spmv",
                                                "build", "time_to_e0", "rdm")}
        except Exception as exc:  # the headline line must survive a failure of this extra leg; it is reported, not hidden
            line["cfg5"] = {"workload": s5["label"], "error": "%s: %s" % (type(exc).__name__, exc)}
        guard.cancel()
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        note("CPU baseline (reference, host cores)")
        line["cpu_baseline"] = cpu_sample(spec, args.cpu_seconds)
    note("done")
    if rank == 0:
        emit(line)
    B.close()
    if not line["parity"]["ok"]:
        note("PARITY GATE FAILED: the timed operator differs from the oracle; the numbers above are not valid")
        return 1
    return 0


HCI_N, HCI_OCC, HCI_STRIDE, HCI_EPS, HCI_EPS_UPDATE = 16, (4, 4), 33, 2.0e-4, 2.0e-2
HCI_BIG_N, HCI_BIG_STRIDE = 18, 3  # every 3rd determinant of FullCI(18, 4a4b): 3.1 M determinants, 1.4e10 candidates


def selected_ci_big_leg(cabi, ctx):
    """add_hci / compute_enpt2 on a case big enough to divide over 8 GPUs (3.1 M determinants): device seconds of walk +
    exchange to the owning ranks + merge, the same case at every rank count."""
    syn = _synthetic()
    _, one, two = syn.synthetic_integrals(HCI_BIG_N, SEED)
    full = cabi.Wfn(ctx, cabi.FULLCI, HCI_BIG_N, *HCI_OCC)  # generated on the device
    dets = np.ascontiguousarray(full.download_dets()[::HCI_BIG_STRIDE])
    full.close()
    c = np.random.default_rng(1).standard_normal(len(dets))
    c /= np.linalg.norm(c)
    ham = cabi.Ham(ctx, HCI_BIG_N, 0.0, one, two)
    out = {"workload": "FullCI(%d, %da%db), every %drd determinant, eps %g" % (HCI_BIG_N, HCI_OCC[0], HCI_OCC[1], HCI_BIG_STRIDE, HCI_EPS),
           "ndet": int(len(dets))}
    for rep in range(2):  # second pass: warm allocations
        wfn = cabi.Wfn(ctx, cabi.FULLCI, HCI_BIG_N, HCI_OCC[0], HCI_OCC[1], dets)
        pt, nt = wfn.compute_enpt2(ham, c, -10.0, HCI_EPS)
        out["enpt2_seconds_device"], out["external_determinants"], out["enpt2"] = wfn.ext_seconds(), int(nt), pt
        new = wfn.add_hci(ham, c, HCI_EPS)
        out["add_hci_seconds_device"], out["added"] = wfn.ext_seconds(), int(len(new))
        wfn.close()
    ham.close()
    return out



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


def leg_watchdog(seconds, rank, line, key, label):
    """Timer around an extra leg that runs after the headline numbers are in `line`: if the leg has not finished after
    `seconds`, rank 0 prints the line with the leg reported as timed out and every rank leaves the process (a hung
    collective cannot be interrupted from Python).  Returns the timer; cancel() it when the leg is done."""
    def fire():
        if rank == 0:
            line[key] = {"workload": label, "error": "watchdog: leg not finished after %.0f s; line printed without it" % seconds}
            note("watchdog: %s leg abandoned" % key)
            emit(line)
        else:
            time.sleep(5.0)  # rank 0 prints first
        os._exit(0 if line.get("parity", {}).get("ok", False) else 1)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


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
    ap.add_argument("--ref-procs", type=int, default=0, help="processes of the reference arm (0 = one per core, <= 32)")
    ap.add_argument("--gate-rows", type=int, default=1000, help="rows per rank compared with the oracle (parity gate)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra legs (config 4 on one GPU and the CPU time-to-E0 legs at N=1, config 5 at N=8)")
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
