#!/usr/bin/env python
"""bench.py -- the headline measurement (contract in the task statement; DESIGN.md section "Measurement").

N = 1 (config.workload = BASELINE.json configs[3]): a synthetic 16384 x 16384 FP32 depth grid (the seamount field of the
reference's generate_csv_grids.cpp:32-70) upsampled 4x in both axes with the bicubic Catmull-Rom stencil -> 65533 x 65533
output cells.  The largest single-GPU configuration of BASELINE.json and the one its HBM roofline is quoted on.

N > 1 (config.workload = BASELINE.json configs[4], as stated): ONE synthetic 65536 x 65536 FP32 grid at a 70 % mask,
IDW gap fill, output rows sharded over the N ranks (strong scaling).  Every rank holds its rows + a 14-row halo of the
known-point input, draws its part of the one global mask from a counter hash (no communication), fills its rows; no
data-path collective.  One process per GPU under torchrun; time = max over ranks between barriers.

One JSON line on rank 0:
  value       output Mcells/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e         same metric through the host-buffer C-ABI calls (auvi_grid_create_slab + auvi_lattice + auvi_grid_destroy):
              host->device copy of the grid (slab) and device->host copy of every output cell inside the timed region
  roofline    dominant kernel: algorithmic bytes / event time vs the measured copy bandwidth in MEASURED_PEAKS.json;
              `traffic` (and, for the gap-fill kernel, issue-slot / pipe utilisation) from the committed ncu captures,
              labelled with their source file
  methods     per method: Mcells/s, ms, fraction of the HBM roofline, depth RMSE (BASELINE metric: "per method")
  cpu_baseline  N = 1 only: the reference's own CPU class (oracle/_ref, GridH::batchCubicInterpolate) on a bounded row
              block of the same lattice, all host threads, warmed up; plus the as-shipped single-thread figure
  gather      N > 1, outside `value`: NCCL gather of the row shards to rank 0, and the same gather with no collective --
              the kernel stores into rank 0's buffer through peer memory (auvi_peer_*)
  same_workload_on_one_gpu  N > 1, rank 0, outside `value`: the whole 65536^2 job on one GPU in the same run (the N = 1 line
              runs configs[3], a different job: this is the strong-scaling reference of the N > 1 lines)
  single_process_multi_gpu  N > 1, rank 0, outside `value`: the same kind of job through auvi_multi_* (one process
              driving all N GPUs behind the C ABI)
  extra       other BASELINE configs (Mariana 50 % and Grid A through the Point-list API, the reference's own GPU
              code recompiled for sm_100a beside it, Grid A 2x lattice) -- informational
`--impl reference` times only the reference CPU implementation on the same config and metric.
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
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))
sys.path.insert(0, ROOT)

N_GRID = 16384
FACTOR = 4
BOUNDS = (-180.0, -160.0, 20.0, 30.0)          # test_interpolation.cpp:143-144
ALGO_BYTES_PER_CELL = 4.0 + 4.0 / (FACTOR * FACTOR)   # f32 out + f32 in / 16
CPU_SAMPLE_ROWS = 256                           # lattice rows timed on the CPU (x 65533 columns)
METRIC = "output Mcells/s (bicubic 4x upsample, 16384^2 f32 -> 65533^2)"

N_FILL = 65536                                  # BASELINE configs[4]
FILL_MASK = 0.70
FILL_BOUNDS = (100.0, 110.0, -10.0, 0.0)
FILL_BYTES_PER_CELL = 8.0                       # f32 read + f32 written per grid cell
FILL_CPU_ROWS = 128                             # grid rows the CPU arm fills per step (x 65536 columns)
METRIC_FILL = "output Mcells/s (IDW gap fill, one 65536^2 f32 grid at 70 % mask, rows sharded)"
IDW_NOTE = ("IDW is an extension: the reference has no IDW (SURVEY.md section 0 fact 1).  Its neighbour SELECTION is the "
            "reference's search, pinned bit-exactly; its VALUE is parity-unpinned by construction (checked against "
            "oracle/interp_oracle.c only)")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _profiled(key):
    """Numbers taken from committed ncu captures (profiles/roofline_traffic.json), with their source label."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index`, so that the pinned host buffers
    the end-to-end leg allocates are first-touched on the GPU's own NUMA node (with one rank per GPU the
    device->host copies of all ranks otherwise cross the socket link).  Returns the core list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def synth_grid_device(torch, n_lat_global, n_lon, row_lo, row_hi, device):
    """Rows [row_lo,row_hi) of the seamount field (generate_csv_grids.cpp:32-70) in FP32, on device.
    Evaluated in FP64 in 1024-row blocks so that the temporaries stay small next to the 17 GB outputs."""
    z = torch.empty((row_hi - row_lo, n_lon), dtype=torch.float32, device=device)
    i = torch.arange(n_lon, device=device, dtype=torch.float64) * (100.0 / (n_lon - 1))
    base = -(10.0 + 2.0 * i)[None, :]
    gx = ((i - 75.0) ** 2)[None, :]
    for r in range(row_lo, row_hi, 1024):
        r1 = min(row_hi, r + 1024)
        j = torch.arange(r, r1, device=device, dtype=torch.float64) * (100.0 / (n_lat_global - 1))
        z[r - row_lo:r1 - row_lo] = (base + 100.0 * torch.exp(-(gx + ((j - 50.0) ** 2)[:, None]) / 450.0)).to(torch.float32)
    return z


# ---- CPU arms -----------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_rows, threads, steps=1, warmup=0, method=1):
    """Reference CPU path (oracle/_ref GridH, else the C port) on `sample_rows` rows of configs[3]'s output lattice.
    -> (Mcells/s, kind, cores, ms per step, sample description)"""
    from oracle import binding as ob
    n = N_GRID
    z = ob.synth_grid(n, n, csv_round=False).astype(np.float32).astype(np.float64)
    rows_out = FACTOR * (n - 1) + 1
    lat_ax = ob.lattice_axis(BOUNDS[2], BOUNDS[3], rows_out)
    lon_ax = ob.lattice_axis(BOUNDS[0], BOUNDS[1], rows_out)
    r0 = rows_out // 2 - sample_rows // 2
    pts = np.zeros((sample_rows * rows_out, 3))
    pts[:, 0] = np.tile(lon_ax, sample_rows)
    pts[:, 1] = np.repeat(lat_ax[r0:r0 + sample_rows], rows_out)
    if ob.ref_available():
        eng, kind = ob.Reference(z, *BOUNDS), "reference"
        run = lambda: eng.batch(method, pts, threads=threads)
        cores = threads
    else:
        eng, kind = ob.Oracle(z, *BOUNDS), "port"
        run = lambda: eng.batch(method, pts)
        cores = 1
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(out).all()
    sample = (f"{sample_rows} consecutive output rows x {rows_out} columns ({pts.shape[0]} cells) of the same "
              f"65533^2 lattice, GridH::batchCubicInterpolate, {cores} thread(s), {warmup} warm-up + {steps} timed step(s)")
    return pts.shape[0] / dt / 1e6, kind, cores, dt * 1e3, sample


def cpu_fill_rate(sample_rows, threads, steps=1, warmup=0):
    """configs[4] on the CPU: IDW gap fill of `sample_rows` rows (all 65536 columns) out of the middle of the one
    65536^2 grid at its 70 % hash mask.  The reference has no IDW, so this is the C port (oracle/interp_oracle.c: the
    reference's search + selection restated, then power-2 weights), run on `threads` slices of the query list.
    -> (Mcells/s counting every cell of the rows like the GPU arm, kind, cores, ms per step, sample description)"""
    from oracle import binding as ob
    n = N_FILL
    r0 = n // 2 - sample_rows // 2
    lo, hi = r0 - 16, r0 + sample_rows + 16                       # the rows a radius-10 search can reach
    z = ob.synth_rows(n, n, lo, hi).astype(np.float32).astype(np.float64)
    z[ob.hash_mask(lo, hi, n, FILL_MASK, 42)] = np.nan
    # the window as a grid of its own: same steps as the global grid, bounds moved to the window's rows
    lat_step = (FILL_BOUNDS[3] - FILL_BOUNDS[2]) / (n - 1)
    wb = (FILL_BOUNDS[0], FILL_BOUNDS[1], FILL_BOUNDS[2] + lo * lat_step, FILL_BOUNDS[2] + (hi - 1) * lat_step)
    meta = dict(n_lat=hi - lo, n_lon=n, min_lon=wb[0], max_lon=wb[1], min_lat=wb[2], max_lat=wb[3])
    rr, cc = np.nonzero(np.isnan(z[16:16 + sample_rows]))
    pts = ob.node_queries(rr + 16, cc, meta)
    eng = ob.Oracle(z, *wb)
    chunks = np.array_split(np.arange(pts.shape[0]), max(1, threads))
    out = np.empty(pts.shape[0])

    def run():
        def work(idx):
            if idx.size:
                out[idx] = eng.batch(ob.IDW, pts[idx[0]:idx[-1] + 1])
        ts = [threading.Thread(target=work, args=(c,)) for c in chunks]     # ctypes releases the GIL inside orc_batch
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(out).all()
    cells = sample_rows * n
    sample = (f"{sample_rows} consecutive grid rows x {n} columns ({cells} cells, {pts.shape[0]} masked) out of the middle of "
              f"the same 65536^2 grid and mask; IDW = oracle/interp_oracle.c (the reference has no IDW), {threads} thread(s), "
              f"{warmup} warm-up + {steps} timed step(s)")
    return cells / dt / 1e6, "port", threads, dt * 1e3, sample


def upsample_config(world):
    import shard
    out_rows = FACTOR * (N_GRID * world - 1) + 1
    out_cols = FACTOR * (N_GRID - 1) + 1
    return {"workload": "BASELINE configs[3]: synthetic 16384x16384 FP32 depth grid, 4x bicubic upsample on 1 B200",
            "grid": [N_GRID, N_GRID], "factor": FACTOR, "method": "bicubic Catmull-Rom",
            "out_cells": -(-out_rows // world) * out_cols,
            "parallelism": f"1 GPU (row-sharding unit: halo {shard.HALO} rows, no collective)",
            "l2_policy": "inputs (1.07 GB) and outputs (17.2 GB) per step exceed the 126 MB L2"}


def fill_config(world):
    import shard
    return {"workload": "BASELINE configs[4]: synthetic 65536x65536 FP32 grid at 70 % mask, IDW k-neighbour gap fill, "
                        f"row-sharded over {world} B200",
            "grid": [N_FILL, N_FILL], "mask_fraction": FILL_MASK, "method": "IDW (k = 4, power 2) over the reference's ring search",
            "out_cells": N_FILL * N_FILL,
            "parallelism": f"output rows sharded x{world}, replicated halo {shard.HALO} rows, one global hash mask, no collective",
            "l2_policy": "inputs and outputs per rank and step (2 x 17.2 GB / N) exceed the 126 MB L2",
            "note": IDW_NOTE}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.gpus > 1:
        rate, kind, cores, ms, sample = cpu_fill_rate(FILL_CPU_ROWS, threads, steps=args.steps, warmup=args.warmup)
        metric, config, scaling = METRIC_FILL, fill_config(args.gpus), "strong"
    else:
        rate, kind, cores, ms, sample = cpu_reference_rate(CPU_SAMPLE_ROWS, threads, steps=args.steps, warmup=args.warmup)
        metric, config, scaling = METRIC, upsample_config(1), "weak"
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": "Mcells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, reference_step="bounded sample: " + sample),
            "cpu_baseline": {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- shared timing helpers ----------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx(args):
    import torch
    import auvi
    auvi.load()
    if not torch.cuda.is_available() or auvi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libauvi has no CPU fallback")
    c = Ctx()
    c.torch, c.auvi, c.args = torch, auvi, args
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    c.all_cores = sorted(os.sched_getaffinity(0))
    c.numa = bind_to_gpu_numa_node(c.local)                         # host buffers of this rank live next to its GPU
    c.dist = None
    if c.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=c.dev)
        c.dist = dist
        c.cpu_group = dist.new_group(backend="gloo")               # host-side waits: no kernel spins on a GPU meanwhile
    assert c.world == args.gpus or c.world == 1, (c.world, args.gpus)
    c.stream = torch.cuda.current_stream().cuda_stream
    c.peak, c.peak_src = _peaks()

    def barrier():
        torch.cuda.synchronize()
        if c.dist is not None:
            c.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if c.dist is None:
            return v
        t = torch.tensor([v], device=c.dev, dtype=torch.float64)
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """-> (ms per step: CUDA events on the launching stream, max over ranks, between barriers; launches)"""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = auvi.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = auvi.launch_count() - l0
        barrier()
        return ms / steps, launches

    def host_barrier():
        """Ranks meet on the CPU (gloo): unlike an NCCL barrier no kernel waits on any GPU while a rank is late."""
        torch.cuda.synchronize()
        if c.dist is not None:
            c.dist.barrier(group=c.cpu_group)

    c.barrier, c.max_over_ranks, c.timed, c.host_barrier = barrier, max_over_ranks, timed, host_barrier
    return c


def host_alloc(torch, shape):
    try:
        return torch.empty(shape, dtype=torch.float32, pin_memory=True)
    except Exception:
        return torch.empty(shape, dtype=torch.float32)


def e2e_through_c_abi(c, h_z, h_out, n_lat_global, n_lon, in_lo, bounds, call, cells_total_of_rows, api):
    """The timed end-to-end region: upload the slab from host memory, run, every result cell back to host memory, release.
    `call(handle)` makes the auvi_lattice call into h_out.  -> the e2e object (value filled by the caller's cell count)."""
    import ctypes as C
    torch, auvi = c.torch, c.auvi
    lib = auvi.load()
    pieces = {"upload_ms": 0.0, "lattice_ms": 0.0, "release_ms": 0.0}

    def e2e_step():
        h = C.c_void_p()
        t0 = time.perf_counter()
        rc = lib.auvi_grid_create_slab(h_z.data_ptr(), auvi.F32, n_lat_global, n_lon, in_lo, h_z.shape[0], *bounds, c.local, C.byref(h))
        assert rc == 0, lib.auvi_last_error()
        t1 = time.perf_counter()
        rc = call(h)
        assert rc == 0, lib.auvi_last_error()
        t2 = time.perf_counter()
        lib.auvi_grid_destroy(h)
        t3 = time.perf_counter()
        pieces["upload_ms"] += (t1 - t0) * 1e3; pieces["lattice_ms"] += (t2 - t1) * 1e3; pieces["release_ms"] += (t3 - t2) * 1e3

    steps = max(1, min(c.args.steps, 3))
    e2e_step()
    c.barrier()
    for k in pieces:
        pieces[k] = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = c.max_over_ranks((time.perf_counter() - t0) / steps)
    return {"value": cells_total_of_rows / dt / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": int(h_z.numel() * 4),
            "d2h_bytes_per_step": int(h_out.numel() * 4), "ms_per_step": dt * 1e3, "steps": steps,
            "pinned_host": bool(h_out.is_pinned()), "pieces_ms": {k: v / steps for k, v in pieces.items()}, "api": api,
            "cpu_affinity": (f"{c.numa[0]}-{c.numa[-1]} ({len(c.numa)} cores, NVML GPU-local)" if c.numa else "unbound"),
            "host_threads": os.environ.get("AUVI_HOST_THREADS", "default: CPUs of this process / LOCAL_WORLD_SIZE, <= 16")}


def d2h_probe(c, e2e, dt_s):
    """Raw pinned device->host copy bandwidth of this rank: alone, and with every rank copying at the same moment."""
    torch = c.torch
    try:
        probe_d = torch.empty(1 << 28, dtype=torch.float32, device=c.dev)
        probe_h = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)
        probe_h.copy_(probe_d); torch.cuda.synchronize()
        t0 = time.perf_counter()
        probe_h.copy_(probe_d, non_blocking=True); torch.cuda.synchronize()
        e2e["pcie_d2h_gbs_raw"] = probe_d.numel() * 4 / (time.perf_counter() - t0) / 1e9
        if c.dist is not None:                                      # all ranks at once: what the host can sink in aggregate
            c.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                probe_h.copy_(probe_d, non_blocking=True)
            torch.cuda.synchronize()
            dtc = c.max_over_ranks((time.perf_counter() - t0) / 3)
            e2e["pcie_d2h_gbs_all_ranks_concurrent"] = probe_d.numel() * 4 / dtc / 1e9
        e2e["pcie_d2h_gbs_in_e2e"] = e2e["d2h_bytes_per_step"] / dt_s / 1e9
        del probe_d, probe_h
    except Exception:
        pass


# ---- N = 1: BASELINE configs[3] -------------------------------------------------------------------------------------------
def run_upsample(c):
    torch, auvi, args = c.torch, c.auvi, c.args
    import shard
    n_lon = n_lat = N_GRID
    out_rows = FACTOR * (n_lat - 1) + 1
    out_cols = FACTOR * (n_lon - 1) + 1
    z = synth_grid_device(torch, n_lat, n_lon, 0, n_lat, c.dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=0, rows=n_lat, keep=z),
                  min_lon=BOUNDS[0], max_lon=BOUNDS[1], min_lat=BOUNDS[2], max_lat=BOUNDS[3], device=c.local)
    out_ld = (out_cols + 3) // 4 * 4                                # 16-byte row pitch
    out = torch.empty((out_rows, out_ld), dtype=torch.float32, device=c.dev)
    cells = out_rows * out_cols

    def step(method=auvi.CUBIC):
        g.lattice_device(method, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, 0, out_rows, out.data_ptr(), out_ld, None, c.stream)

    sampler = ClockSampler(c.local)
    sampler.start()
    ms_step, launches = c.timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    value = cells / (ms_step * 1e-3) / 1e6
    achieved = ALGO_BYTES_PER_CELL * cells / (ms_step * 1e-3) / 1e9
    prof = _profiled("upsample_cubic_f32") or {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": c.peak, "unit": "GB/s", "frac": achieved / c.peak,
                "traffic": prof.get("dram_bytes_per_launch"), "traffic_source": prof.get("source"),
                "kernel": "upsample_tiled_kernel<float,CUBIC>", "peak_source": c.peak_src,
                "algorithmic_bytes_per_cell": ALGO_BYTES_PER_CELL, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL * cells,
                "tma": bool(g.uses_tma)}

    # ---- end to end: host buffers through the C-ABI (grid upload + every output cell back) ----------
    e2e = None
    if not args.no_e2e:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable")) * 1024
        e2e_rows = shard.e2e_row_budget(out_rows, out_cols * 4, avail, 1)
        h_out = host_alloc(torch, (e2e_rows, out_cols))
        h_z = host_alloc(torch, (n_lat, n_lon))
        h_z.copy_(z)
        lib = auvi.load()
        e2e = e2e_through_c_abi(c, h_z, h_out, n_lat, n_lon, 0, BOUNDS,
                                lambda h: lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, 0, e2e_rows, h_out.data_ptr()),
                                e2e_rows * out_cols, "auvi_grid_create_slab + auvi_lattice(host_out) + auvi_grid_destroy")
        e2e["rows"] = e2e_rows; e2e["rows_of_lattice"] = out_rows
        d2h_probe(c, e2e, e2e["ms_per_step"] * 1e-3)
        # the same call with a PAGEABLE destination (what a std::vector / numpy caller has), on a bounded row range
        try:
            import ctypes as C
            p_rows = min(e2e_rows, 8192)
            pageable = np.zeros((p_rows, out_cols), dtype=np.float32)             # touched, like a value-initialised vector
            h = C.c_void_p()
            assert lib.auvi_grid_create_slab(h_z.data_ptr(), auvi.F32, n_lat, n_lon, 0, n_lat, *BOUNDS, c.local, C.byref(h)) == 0
            pp = pageable.ctypes.data
            assert lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, 0, p_rows, pp) == 0
            t0 = time.perf_counter()
            for _ in range(2):
                assert lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, 0, p_rows, pp) == 0
            dtp = (time.perf_counter() - t0) / 2
            lib.auvi_grid_destroy(h)
            e2e["pageable_host"] = {"rows": p_rows, "ms": dtp * 1e3, "GBps_d2h": p_rows * out_cols * 4 / dtp / 1e9,
                                    "Mcells_per_s": p_rows * out_cols / dtp / 1e6,
                                    "equals_pinned_result": bool(np.array_equal(pageable[:64], h_out[:64].numpy()))}
            del pageable
        except Exception as exc:
            e2e["pageable_host"] = {"unavailable": repr(exc)[:200]}
        chk = slice(e2e_rows // 2, e2e_rows // 2 + 8)
        assert torch.equal(h_out[chk], out[chk, :out_cols].cpu())      # the host result equals the device-resident result
        del h_out, h_z

    # ---- per method (BASELINE metric: "output Mcells/s per method ... (% roofline) ...; depth RMSE") ----------------------
    methods, extra = {}, {}
    if not args.no_extra:
        ups = {"bicubic": {"Mcells_per_s": value, "ms": ms_step, "hbm_frac": achieved / c.peak}}
        ms_b, _ = c.timed(lambda: step(auvi.BILINEAR), max(3, args.steps // 2), 2)
        ups["bilinear"] = {"Mcells_per_s": cells / (ms_b * 1e-3) / 1e6, "ms": ms_b,
                           "hbm_frac": ALGO_BYTES_PER_CELL * cells / (ms_b * 1e-3) / 1e9 / c.peak}
        # latitude-only variant: 4x more rows, the same columns -> 4 + 4/4 = 5 B per output cell
        out_lat = out.view(-1)[: out_rows * n_lon].view(out_rows, n_lon)
        ms_l, _ = c.timed(lambda: g.lattice_device(auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, 1, 0, 0, out_rows, out_lat.data_ptr(), n_lon, None, c.stream),
                          max(3, args.steps // 2), 2)
        ups["bicubic_latitude_only"] = {"Mcells_per_s": out_rows * n_lon / (ms_l * 1e-3) / 1e6, "ms": ms_l,
                                        "hbm_frac": 5.0 * out_rows * n_lon / (ms_l * 1e-3) / 1e9 / c.peak}
        # the 2x lattice of the same grid (the factor the reference's driver uses): window-load form of the FP32 bicubic kernel,
        # 4 + 4/4 = 5 B per output cell
        r2, c2 = 2 * (n_lat - 1) + 1, 2 * (n_lon - 1) + 1
        ld2 = (c2 + 3) // 4 * 4
        out2 = out.view(-1)[: r2 * ld2].view(r2, ld2)
        for name2, meth2 in (("bicubic_2x", auvi.CUBIC), ("bilinear_2x", auvi.BILINEAR)):
            ms2, _ = c.timed(lambda: g.lattice_device(meth2, auvi.AXIS_EXPANDED, 2, 2, 0, 0, r2, out2.data_ptr(), ld2, None, c.stream),
                             max(3, args.steps // 4), 2)
            ups[name2] = {"Mcells_per_s": r2 * c2 / (ms2 * 1e-3) / 1e6, "ms": ms2, "hbm_frac": 5.0 * r2 * c2 / (ms2 * 1e-3) / 1e9 / c.peak}
        methods["upsample_4x_16384sq_f32 (configs[3])"] = ups
        del out
        torch.cuda.empty_cache()
        methods.update(methods_gap_fill(c))
        methods["mariana_50pct_point_list (configs[1])"] = methods_mariana(c)
        extra.update(extra_grid_a_points(c))
        methods["grid_a_2x_lattice_4000x3200_f64 (configs[0])"] = extra_grid_a_lattice(c)
        extra["reference_gpu_sm100a"] = extra_reference_gpu(c)

    cpu = None
    os.sched_setaffinity(0, c.all_cores)                             # the CPU baseline gets every core of the box back
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, kind, cores, ms_cpu, sample = cpu_reference_rate(CPU_SAMPLE_ROWS, threads, steps=2, warmup=1)
        cpu = {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": kind, "sample": sample}
        r1, _, _, _, s1 = cpu_reference_rate(32, 1, steps=1, warmup=1)
        cpu["as_shipped_1_thread"] = {"value": r1, "unit": "Mcells/s", "cores": 1, "sample": s1,
                                      "what": "GridH::batchCubicInterpolate as the reference ships it: one thread (GridH.cpp:422-448)"}

    line = {"metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": upsample_config(1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "methods": methods, "extra": extra}
    print(json.dumps(line), flush=True)
    g.close()


def methods_gap_fill(c):
    """BASELINE configs[4] on one GPU: FP32 grid at 70 % mask, full-grid gap fill -- every method at 16384^2 (with the depth
    RMSE against the unmasked field, reduced on the device) and the IDW headline of that config at the full 65536^2."""
    torch, auvi = c.torch, c.auvi
    res = {}
    pipes = _profiled("fill_idw_f32")
    for n, methods in ((16384, (("idw", auvi.IDW), ("nn", auvi.NN), ("kriging", auvi.KRIGING), ("bilinear", auvi.BILINEAR),
                                ("nearest4_mean(cubic fallback)", auvi.CUBIC))), (N_FILL, (("idw", auvi.IDW), ("nn", auvi.NN)))):
        z = synth_grid_device(torch, n, n, 0, n, c.dev)
        truth = z.clone() if n <= 16384 else None
        g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
                      min_lon=FILL_BOUNDS[0], max_lon=FILL_BOUNDS[1], min_lat=FILL_BOUNDS[2], max_lat=FILL_BOUNDS[3], device=c.dev.index)
        g.mask_hash(FILL_MASK, seed=42, count=False, stream=c.stream)   # counter-hash mask drawn on the device (csrc/ingest.cu)
        out = torch.empty((n, n), dtype=torch.float32, device=c.dev)
        tab = {}
        for name, meth in methods:
            fn = lambda: g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, c.stream)
            ms, _ = c.timed(fn, 5 if n <= 16384 else 3, 2)
            row = {"Mcells_per_s": n * n / (ms * 1e-3) / 1e6, "ms": ms, "hbm_frac": FILL_BYTES_PER_CELL * n * n / (ms * 1e-3) / 1e9 / c.peak,
                   "nan_left": int(torch.isnan(out).sum().item())}
            if truth is not None:
                t0 = time.perf_counter()
                mae, rmse, mx, n_nan, cnt = g.fill_metrics_device(out.data_ptr(), n, truth.data_ptr(), n, 0, n, c.stream)
                t_met = (time.perf_counter() - t0) * 1e3           # synchronous call: two kernels + the read-back (csrc/metrics.cu)
                row.update({"rmse_m": rmse, "mae_m": mae, "max_m": mx, "filled_cells": cnt,
                            "metrics_call_ms": t_met, "metrics_hbm_frac": 12.0 * n * n / (t_met * 1e-3) / 1e9 / c.peak})
            if name == "idw":
                row["note"] = IDW_NOTE
                if pipes:
                    row["pipes_from_ncu"] = pipes
            tab[name] = row
        res[f"gap_fill_70pct_{n}sq_f32 (configs[4]{' on one GPU' if n == N_FILL else ' scaled'})"] = tab
        g.close()
        del z, out, truth
        torch.cuda.empty_cache()
    return res


def methods_mariana(c):
    """BASELINE configs[1]: Mariana tile at 50 % removal through the Point-list API (host buffers, what GridD::batch* calls),
    all methods, with RMSE against the unmasked GEBCO truth computed on the device."""
    torch, auvi = c.torch, c.auvi
    from oracle import binding as ob          # fixture loader only (tile + seed-42 mask), not the computation
    case = ob.masked_case("mariana", 0.5)
    g = auvi.Grid(case["z"], *case["bounds"], device=c.local)
    d_truth = torch.from_numpy(case["truth"]).cuda()
    res = {}
    for name, meth in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC), ("kriging", auvi.KRIGING),
                       ("nn", auvi.NN), ("idw", auvi.IDW), ("bilinear_search(opt-in)", auvi.BILINEAR_SEARCH),
                       ("idw_true_4_nearest(opt-in)", auvi.IDW_KNN), ("kriging_fitted_variogram(opt-in)", auvi.KRIGING_FITTED)):
        g.interp_points(meth, case["pts"])
        t0 = time.perf_counter()
        for _ in range(5):
            est = g.interp_points(meth, case["pts"])
        dt = (time.perf_counter() - t0) / 5
        d_est = torch.from_numpy(est).cuda()
        mae, rmse, mx, n_nan = auvi.error_metrics_device(d_truth.data_ptr(), d_est.data_ptr(), auvi.F64, est.size)
        res[name] = {"Mpts_per_s_e2e": est.size / dt / 1e6, "ms_e2e": dt * 1e3, "kernel_ms": g.last_kernel_ms, "rmse_m": rmse,
                     "mae_m": mae, "max_m": mx, "n_nan": n_nan, "n": int(est.size)}
    g.close()
    return res


def extra_grid_a_points(c):
    """The reference's Grid-A benchmark row (results/grid_A_runtimes_averaged.csv:8): 5,000,000 random query points
    on the 4000 x 3200 synthetic grid through the Point-list API with host buffers (what GridD::batch* calls)."""
    auvi = c.auvi
    from oracle import binding as ob          # synthetic-field generator only
    z = ob.synth_grid(3200, 4000, csv_round=False)
    g = auvi.Grid(z, *BOUNDS, device=c.local)
    rng = np.random.RandomState(1)
    n = 5_000_000
    pts = np.zeros((n, 3))
    pts[:, 0] = rng.uniform(BOUNDS[0], BOUNDS[1], n)
    pts[:, 1] = rng.uniform(BOUNDS[2], BOUNDS[3], n)
    res = {}
    for name, meth in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC), ("kriging", auvi.KRIGING)):
        g.interp_points(meth, pts)
        t0 = time.perf_counter()
        for _ in range(3):
            g.interp_points(meth, pts)
        dt = (time.perf_counter() - t0) / 3
        res[f"grid_a_5M_random_points_{name}"] = {"Mpts_per_s_e2e": n / dt / 1e6, "ms_e2e": dt * 1e3,
                                                  "kernel_ms": g.last_kernel_ms}
    g.close()
    return res


def extra_grid_a_lattice(c):
    """BASELINE configs[0] at the generator's shipped size: the 4000 x 3200 FP64 Grid A, 2x expanded lattice
    (7999 x 6399 = 51.2 M cells, test_interpolation.cpp:283-297), device-resident, every method."""
    torch, auvi = c.torch, c.auvi
    from oracle import binding as ob          # synthetic-field generator only
    z = torch.from_numpy(ob.synth_grid(3200, 4000, csv_round=False)).to(c.dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F64, n_lat=3200, n_lon=4000, ld=4000, row0=0, rows=3200, keep=z),
                  min_lon=BOUNDS[0], max_lon=BOUNDS[1], min_lat=BOUNDS[2], max_lat=BOUNDS[3], device=c.dev.index)
    rows, cols = g.lattice_dims(auvi.AXIS_EXPANDED, 2, 2)
    out = torch.empty((rows, 8000), dtype=torch.float64, device=c.dev)
    res = {}
    for name, meth in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC), ("kriging", auvi.KRIGING), ("nn", auvi.NN),
                       ("idw", auvi.IDW)):
        fn = lambda: g.lattice_device(meth, auvi.AXIS_EXPANDED, 2, 2, 0, 0, rows, out.data_ptr(), 8000, None, c.stream)
        ms, _ = c.timed(fn, 5, 1)
        res[name] = {"Mcells_per_s": rows * cols / (ms * 1e-3) / 1e6, "ms": ms,
                                                 "hbm_frac": 10.0 * rows * cols / (ms * 1e-3) / 1e9 / c.peak}
    g.close()
    return res


def extra_reference_gpu(c):
    """The secondary comparator: the reference's OWN GPU code -- kernels.cu +
    GridD.cu, unmodified, compiled for sm_100a by oracle/Makefile -- on the same box, same inputs, beside points_kernel.
    Runs in a child process (tools/ref_gpu_compare.py): the reference's kriging kernel printf()s from the device
    (kernels.cu:469-476) and its host class exit()s on errors (GridD.h:9-16); neither may touch this process's one JSON line."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_compare.py"), str(c.local)], capture_output=True,
                           text=True, timeout=900)
        for line in reversed(r.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": "no result line; exit code %d; %s" % (r.returncode, r.stderr[-200:])}
    except Exception as exc:
        return {"unavailable": repr(exc)[:200]}


# ---- N > 1: BASELINE configs[4] as stated --------------------------------------------------------------------------------
def run_fill_sharded(c):
    torch, auvi, args, dist = c.torch, c.auvi, c.args, c.dist
    import shard
    n, world, rank = N_FILL, c.world, c.rank
    plan = shard.plan_rows(n, 1, world, rank)
    lo, hi, in_lo, in_hi = plan.row_lo, plan.row_hi, plan.in_lo, plan.in_hi
    my_rows = hi - lo
    z = synth_grid_device(torch, n, n, in_lo, in_hi, c.dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=in_lo, rows=in_hi - in_lo, keep=z),
                  min_lon=FILL_BOUNDS[0], max_lon=FILL_BOUNDS[1], min_lat=FILL_BOUNDS[2], max_lat=FILL_BOUNDS[3], device=c.dev.index)
    g.mask_hash(FILL_MASK, seed=42, count=False, stream=c.stream)
    out = torch.empty((my_rows, n), dtype=torch.float32, device=c.dev)
    cells_total, cells_rank = n * n, my_rows * n

    def step(method=auvi.IDW):
        g.lattice_device(method, auvi.AXIS_NODES, 1, 1, 1, lo, hi, out.data_ptr(), n, None, c.stream)

    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    ms_step, launches = c.timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = cells_total / (ms_step * 1e-3) / 1e6
    achieved = FILL_BYTES_PER_CELL * cells_rank / (ms_step * 1e-3) / 1e9
    nan_left = torch.tensor([int(torch.isnan(out).sum().item())], device=c.dev, dtype=torch.int64)
    dist.all_reduce(nan_left)
    pipes = _profiled("fill_idw_f32") or {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": c.peak, "unit": "GB/s", "frac": achieved / c.peak,
                "traffic": pipes.get("dram_bytes_per_launch"), "traffic_source": pipes.get("source"),
                "kernel": "fill_tiled_kernel<float,IDW,true>", "peak_source": c.peak_src, "per": "GPU (every rank runs the same kernel on its rows)",
                "algorithmic_bytes_per_cell": FILL_BYTES_PER_CELL, "algorithmic_bytes_per_launch": FILL_BYTES_PER_CELL * cells_rank,
                "pipes_from_ncu": pipes or None,
                "reading": "the gap-fill kernel is bound by instruction issue (integer / bit work of the ring search + FP64 distance "
                           "compares), not by HBM: the HBM fraction is small by construction, the pipe utilisation explains it",
                "tma": bool(g.uses_tma), "nan_left": int(nan_left.item())}

    # ---- end to end: the masked slab from host memory, every filled cell back to host memory ----
    e2e = None
    if not args.no_e2e:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable")) * 1024
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        e2e_rows = shard.e2e_row_budget(my_rows, 2 * n * 4, avail, local_world)
        e2e_rows_all = int(c.max_over_ranks(-e2e_rows) * -1)          # every rank uses the smallest budget
        e2e_rows = e2e_rows_all
        slab_hi = min(in_hi, lo + e2e_rows + 2 + shard.HALO)
        h_z = host_alloc(torch, (slab_hi - in_lo, n))
        h_z.copy_(z[:slab_hi - in_lo])                                # the MASKED slab (NaN cells included)
        h_out = host_alloc(torch, (e2e_rows, n))
        lib = auvi.load()
        e2e = e2e_through_c_abi(c, h_z, h_out, n, n, in_lo, FILL_BOUNDS,
                                lambda h: lib.auvi_lattice(h, auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, lo, lo + e2e_rows, h_out.data_ptr()),
                                cells_total * (e2e_rows / my_rows), "auvi_grid_create_slab + auvi_lattice(fill, host_out) + auvi_grid_destroy")
        e2e["rows_per_rank"] = e2e_rows; e2e["rows_of_shard"] = my_rows
        d2h_probe(c, e2e, e2e["ms_per_step"] * 1e-3)
        chk = slice(e2e_rows // 2, e2e_rows // 2 + 8)
        assert torch.equal(h_out[chk].nan_to_num(nan=7.0), out[chk].cpu().nan_to_num(nan=7.0))
        del h_out, h_z

    # ---- per method on the same sharded grid ----
    methods = {}
    if not args.no_extra:
        tab = {"idw": {"Mcells_per_s": value, "ms": ms_step, "hbm_frac_per_gpu": achieved / c.peak, "note": IDW_NOTE}}
        for name, meth in (("nn", auvi.NN), ("kriging", auvi.KRIGING), ("nearest4_mean(cubic fallback)", auvi.CUBIC), ("bilinear", auvi.BILINEAR)):
            ms, _ = c.timed(lambda: step(meth), 3, 1)
            tab[name] = {"Mcells_per_s": cells_total / (ms * 1e-3) / 1e6, "ms": ms,
                         "hbm_frac_per_gpu": FILL_BYTES_PER_CELL * cells_rank / (ms * 1e-3) / 1e9 / c.peak}
        methods[f"gap_fill_70pct_65536sq_f32_sharded_x{world} (configs[4])"] = tab
        step()                                                       # `out` holds the IDW result again
        torch.cuda.synchronize()

    gather = gather_to_root(c, g, out, lo, my_rows, n, ms_step)
    single, one_gpu = None, None
    if not args.no_extra:
        single = single_process_multi_gpu(c)
        one_gpu = same_workload_on_one_gpu(c)

    if rank == 0:
        line = {"metric": METRIC_FILL, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": fill_config(world),
                "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "methods": methods, "same_workload_on_one_gpu": one_gpu, "gather": gather, "single_process_multi_gpu": single}
        print(json.dumps(line), flush=True)
    g.close()


def gather_to_root(c, g, out, lo, my_rows, n, ms_step):
    """The only collective near the path -- gathering output shards to one consumer (not in `value`): NCCL gather, and the
    same gather with NO collective call: every rank's fill kernel stores its rows straight into rank 0's buffer over NVLink
    (peer memory mapped through the C ABI's CUDA-IPC entries; the kernel only sees a pointer)."""
    torch, auvi, dist, world, rank = c.torch, c.auvi, c.dist, c.world, c.rank
    import ctypes as C
    g_rows = min(my_rows, (2 << 30) // (n * 4))                      # bounded: at most 2 GiB per rank
    sendbuf = out[:g_rows]
    recv = [torch.empty_like(sendbuf) for _ in range(world)] if rank == 0 else None
    dist.gather(sendbuf, recv, dst=0)                                 # warm-up (NCCL over NVLink)
    c.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.gather(sendbuf, recv, dst=0)
    e1.record()
    torch.cuda.synchronize()
    g_ms = c.max_over_ranks(e0.elapsed_time(e1))
    g_bytes = g_rows * n * 4 * (world - 1)
    gather = {"api": "torch.distributed.gather (NCCL)", "rows_per_rank": g_rows, "bytes_into_root": g_bytes, "ms": g_ms,
              "GBps_into_root": g_bytes / (g_ms * 1e-3) / 1e9, "full_gather_ms_estimate": g_ms * my_rows / g_rows}
    try:
        lib = auvi.load()
        good, slab, opened = 1.0, None, None
        hbuf = (C.c_ubyte * 72)()
        if rank == 0:
            try:
                slab = torch.empty((world * g_rows, n), dtype=torch.float32, device=c.dev)
                if lib.auvi_peer_export(slab.data_ptr(), hbuf) != 0:
                    good = 0.0
            except Exception:
                good = 0.0
        ht = torch.tensor(list(hbuf), dtype=torch.uint8, device=c.dev)
        dist.broadcast(ht, src=0)
        base_ptr = 0
        if rank == 0:
            base_ptr = slab.data_ptr() if slab is not None else 0
        else:
            hbuf = (C.c_ubyte * 72)(*ht.cpu().tolist())
            opened = C.c_void_p()
            if lib.auvi_peer_open(hbuf, C.byref(opened)) != 0:
                good, opened = 0.0, None
            else:
                base_ptr = opened.value
        flag = torch.tensor([good], device=c.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) == 0.0:
            raise RuntimeError("peer mapping through CUDA IPC failed on some rank")
        dst = base_ptr + rank * g_rows * n * 4
        fused = lambda: g.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, lo, lo + g_rows, dst, n, None, c.stream)
        fused()
        c.barrier()
        e0.record()
        fused()
        e1.record()
        torch.cuda.synchronize()
        f_ms = c.max_over_ranks(e0.elapsed_time(e1))
        c.barrier()
        ok = True
        if rank == 0:
            for r in range(world):
                a, b = slab[r * g_rows:(r + 1) * g_rows], recv[r]
                ok = ok and bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
        gather["fused_peer_store"] = {
            "what": "the fill kernel writes its rows into rank 0's buffer over NVLink (no NCCL call, no staging copy)",
            "ms_compute_plus_transfer": f_ms, "ms_kernel_then_nccl_gather": ms_step * g_rows / my_rows + g_ms,
            "equals_nccl_gather": bool(ok), "peer_mapping": "cudaIpc (auvi_peer_export / auvi_peer_open)"}
        if opened is not None:
            lib.auvi_peer_close(opened)
        del slab
    except Exception as exc:
        gather["fused_peer_store"] = {"unavailable": repr(exc)[:200]}
    del recv
    return gather


def same_workload_on_one_gpu(c):
    """The strong-scaling reference measured in the same run: the WHOLE 65536^2 grid, same mask, IDW, on rank 0's GPU alone
    (the other ranks wait on the CPU).  The N = 1 bench line runs BASELINE configs[3], not this workload, so the per-N
    `value`s of this line and that one are not the same job; this object is what N = 1 of THIS job costs."""
    torch, auvi = c.torch, c.auvi
    res = None
    c.barrier()
    c.host_barrier()
    if c.rank == 0:
        try:
            n = N_FILL
            z = synth_grid_device(torch, n, n, 0, n, c.dev)
            g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
                          min_lon=FILL_BOUNDS[0], max_lon=FILL_BOUNDS[1], min_lat=FILL_BOUNDS[2], max_lat=FILL_BOUNDS[3], device=c.dev.index)
            g.mask_hash(FILL_MASK, seed=42, count=False, stream=c.stream)
            out = torch.empty((n, n), dtype=torch.float32, device=c.dev)
            fn = lambda: g.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, c.stream)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            res = {"n_gpus": 1, "ms_per_step": ms, "Mcells_per_s": n * n / (ms * 1e-3) / 1e6, "steps": 5,
                   "what": "the same 65536^2 grid, mask and method on one GPU of this box, in this run"}
            g.close()
            del z, out
            torch.cuda.empty_cache()
        except Exception as exc:
            res = {"unavailable": repr(exc)[:200]}
    c.host_barrier()
    return res


def single_process_multi_gpu(c):
    """The multi-GPU entry of the C ABI (auvi_multi_*: ONE process, a host thread per device) on rank 0 while the other ranks
    wait on the CPU (a gloo barrier: an NCCL barrier would leave a spinning kernel on their GPUs, which rank 0 is about to
    use): a 16384^2 grid of the same field and mask fraction from host memory, IDW fill, result back to host memory and --
    device form -- left sharded / gathered into device 0 by the kernels' own peer stores."""
    torch, auvi, world, rank = c.torch, c.auvi, c.world, c.rank
    res = None
    c.barrier()
    c.host_barrier()
    if rank == 0:
        try:
            n = 16384
            z = synth_grid_device(torch, n, n, 0, n, c.dev).cpu().numpy()
            m = auvi.MultiGrid(z, *FILL_BOUNDS, n_gpus=world)
            m.mask_hash(FILL_MASK, seed=42)
            host = np.empty((n, n), dtype=np.float32)
            m.lattice(auvi.IDW, auvi.AXIS_NODES, 1, 1, fill=1, out=host)
            t0 = time.perf_counter()
            m.lattice(auvi.IDW, auvi.AXIS_NODES, 1, 1, fill=1, out=host)
            dt = time.perf_counter() - t0
            bufs, ptrs = [], []
            for k in range(world):
                dev, r_lo, r_hi = m.shard(k, 1)
                b = torch.empty((r_hi - r_lo, n), dtype=torch.float32, device=f"cuda:{dev}")
                bufs.append(b); ptrs.append(b.data_ptr())
            m.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, ptrs, n); m.sync()
            t0 = time.perf_counter()
            for _ in range(3):
                m.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, ptrs, n)
            m.sync()
            wall = (time.perf_counter() - t0) / 3
            res = {"api": "auvi_multi_create + auvi_multi_mask_hash + auvi_multi_lattice[_device]", "grid": [n, n], "devices": world,
                   "host_to_host_ms": dt * 1e3, "host_to_host_Mcells_per_s": n * n / dt / 1e6,
                   "device_resident_wall_ms": wall * 1e3, "device_resident_kernel_ms_max_over_devices": m.last_kernel_ms,
                   "device_resident_Mcells_per_s": n * n / wall / 1e6,
                   "nan_left": int(np.isnan(host).sum())}
            try:
                m.enable_peer(0)
                root = torch.empty((n, n), dtype=torch.float32, device=f"cuda:{m.shard(0, 1)[0]}")
                ptrs = [root.data_ptr() + m.shard(k, 1)[1] * n * 4 for k in range(world)]
                m.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, ptrs, n); m.sync()
                t0 = time.perf_counter()
                m.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, ptrs, n); m.sync()
                res["gathered_on_device0_by_peer_stores_ms"] = (time.perf_counter() - t0) * 1e3
                res["gathered_equals_host_result"] = bool(np.array_equal(np.nan_to_num(root.cpu().numpy(), nan=7.0), np.nan_to_num(host, nan=7.0)))
                del root
            except Exception as exc:
                res["gathered_on_device0_by_peer_stores_ms"] = "unavailable: " + repr(exc)[:160]
            m.close()
        except Exception as exc:
            res = {"unavailable": repr(exc)[:300]}
    c.host_barrier()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # 0.56 s timed region at N = 1: several clock samples under load
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the per-method table and the informational extras")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    c = make_ctx(args)
    if c.world > 1:
        run_fill_sharded(c)
        c.dist.destroy_process_group()
    else:
        run_upsample(c)


if __name__ == "__main__":
    main()
