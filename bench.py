#!/usr/bin/env python
"""bench.py -- the headline measurement (contract in the task statement; DESIGN.md section "Measurement").

Workload (config.workload): BASELINE.json configs[3] -- a synthetic 16384 x 16384 FP32 depth grid
(the seamount field of the reference's generate_csv_grids.cpp:32-70), upsampled 4x in both axes with
the bicubic Catmull-Rom stencil -> 65533 x 65533 output cells per GPU.  It is the largest
single-GPU configuration of BASELINE.json and the one its roofline is quoted on (HBM-bound).
N > 1: weak scaling -- the global grid is (16384*N) x 16384, output rows are sharded across ranks,
each rank holds its input row slab + halo; no data-path collective (DESIGN.md "Multi-GPU").

One JSON line on rank 0:
  value       output Mcells/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e         same metric through the host-buffer C-ABI calls (auvi_grid_create_slab + auvi_lattice + auvi_grid_destroy):
              host->device copy of the grid and device->host copy of every output cell inside the timed region
              (pinned buffers; `pageable_host` = the same call into ordinary host memory)
  gather      N > 1 only, outside `value`: NCCL gather of the row shards to rank 0, and the same gather with no
              collective -- the kernel stores into rank 0's buffer through peer memory (auvi_peer_*)
  roofline    dominant kernel (upsample_tiled_kernel<float,CUBIC>): algorithmic bytes / event time vs
              the measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU class (oracle/_ref, GridH::batchCubicInterpolate) on a bounded
              row block of the same lattice, all host threads
  extra       the other methods / BASELINE configs (bilinear and latitude-only upsample, every gap-fill method on a
              70 % masked grid incl. the stated 65536^2 one -- row-sharded over the ranks when N > 1 --, Mariana 50 %
              through the Point-list API with RMSE, Grid A points and 2x lattice) -- informational
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


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("upsample_cubic_f32_dram_bytes_per_launch")
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


def cpu_reference_rate(sample_rows, threads, steps=1, warmup=0, method=1):
    """Reference CPU path (oracle/_ref GridH, else the C port) on `sample_rows` rows of the workload's
    output lattice.  -> (Mcells/s, kind, cores, ms per step, sample description)"""
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
              f"65533^2 lattice, GridH::batchCubicInterpolate")
    return pts.shape[0] / dt / 1e6, kind, cores, dt * 1e3, sample


def workload_config(world):
    """The `config` object both arms print: BASELINE configs[3], weak-scaled by rows over `world` GPUs."""
    import shard
    out_rows = FACTOR * (N_GRID * world - 1) + 1
    out_cols = FACTOR * (N_GRID - 1) + 1
    return {"workload": "BASELINE configs[3]: synthetic 16384x16384 FP32 depth grid, 4x bicubic upsample "
                        "(per GPU; N GPUs hold a (16384*N) x 16384 grid, output rows sharded)",
            "grid_per_gpu": [N_GRID, N_GRID], "factor": FACTOR, "method": "bicubic Catmull-Rom",
            "out_cells_per_gpu": -(-out_rows // world) * out_cols,
            "parallelism": f"row-sharded x{world}, halo {shard.HALO} rows, no collective",
            "l2_policy": "inputs (1.07 GB) and outputs (17.2 GB) per step exceed the 126 MB L2"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, kind, cores, ms, sample = cpu_reference_rate(CPU_SAMPLE_ROWS, threads, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "Mcells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.gpus), reference_step="bounded sample: " + sample),
            "cpu_baseline": {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # 0.56 s timed region: several clock samples under load
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the informational per-method table")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import auvi
    auvi.load()
    if not torch.cuda.is_available() or auvi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libauvi has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cores = sorted(os.sched_getaffinity(0))
    numa = bind_to_gpu_numa_node(local)                              # host buffers of this rank live next to its GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, (world, args.gpus)

    # ---- the rank's share of the global problem ---------------------------------------------------
    n_lon = N_GRID
    n_lat_global = N_GRID * world
    out_rows_global = FACTOR * (n_lat_global - 1) + 1
    out_cols = FACTOR * (n_lon - 1) + 1
    import shard
    plan = shard.plan_rows(n_lat_global, FACTOR, world, rank)
    row_lo, row_hi, in_lo, in_hi, halo = plan.row_lo, plan.row_hi, plan.in_lo, plan.in_hi, shard.HALO
    bounds = (BOUNDS[0], BOUNDS[1], BOUNDS[2], BOUNDS[2] + (BOUNDS[3] - BOUNDS[2]) * world)
    z = synth_grid_device(torch, n_lat_global, n_lon, in_lo, in_hi, dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n_lat_global, n_lon=n_lon, ld=n_lon, row0=in_lo,
                             rows=in_hi - in_lo, keep=z), min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2],
                  max_lat=bounds[3], device=local)
    out_ld = (out_cols + 3) // 4 * 4                                # 16-byte row pitch
    my_rows = row_hi - row_lo
    out = torch.empty((my_rows, out_ld), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    cells_rank = my_rows * out_cols
    cells_total = out_rows_global * out_cols

    def step(method=auvi.CUBIC):
        g.lattice_device(method, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, row_lo, row_hi, out.data_ptr(), out_ld, None, stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
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
        ms = e0.elapsed_time(e1)
        launches = auvi.launch_count() - l0
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms / steps, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = cells_total / (ms_step * 1e-3) / 1e6
    peak, peak_src = _peaks()
    achieved = ALGO_BYTES_PER_CELL * cells_rank / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _traffic(), "kernel": "upsample_tiled_kernel<float,CUBIC>", "peak_source": peak_src,
                "algorithmic_bytes_per_cell": ALGO_BYTES_PER_CELL, "tma": bool(g.uses_tma)}

    # ---- end to end: host buffers through the C-ABI (grid upload + every output cell back) ----------
    e2e = None
    if not args.no_e2e:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable")) * 1024
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        e2e_rows = shard.e2e_row_budget(my_rows, out_cols * 4, avail, local_world)
        try:
            h_out = torch.empty((e2e_rows, out_cols), dtype=torch.float32, pin_memory=True)
            h_z = torch.empty((in_hi - in_lo, n_lon), dtype=torch.float32, pin_memory=True)
        except Exception:
            h_out = torch.empty((e2e_rows, out_cols), dtype=torch.float32)
            h_z = torch.empty((in_hi - in_lo, n_lon), dtype=torch.float32)
        h_z.copy_(z)
        lib = auvi.load()
        import ctypes as C

        pieces = {"upload_ms": 0.0, "lattice_ms": 0.0, "release_ms": 0.0}

        def e2e_step():
            # what a caller with host arrays does: upload the slab, run, get every cell back, release
            h = C.c_void_p()
            t0 = time.perf_counter()
            rc = lib.auvi_grid_create_slab(h_z.data_ptr(), auvi.F32, n_lat_global, n_lon, in_lo, in_hi - in_lo, *bounds, local, C.byref(h))
            assert rc == 0, lib.auvi_last_error()
            t1 = time.perf_counter()
            rc = lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, row_lo, row_lo + e2e_rows, h_out.data_ptr())
            assert rc == 0, lib.auvi_last_error()
            t2 = time.perf_counter()
            lib.auvi_grid_destroy(h)
            t3 = time.perf_counter()
            pieces["upload_ms"] += (t1 - t0) * 1e3; pieces["lattice_ms"] += (t2 - t1) * 1e3; pieces["release_ms"] += (t3 - t2) * 1e3

        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(1):
            e2e_step()
        barrier()
        for k in pieces:
            pieces[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_cells = cells_total * (e2e_rows / my_rows)      # == cells_total unless host memory forced a row cut
        e2e = {"value": e2e_cells / dt / 1e6, "unit": "Mcells/s", "rows_per_rank": e2e_rows, "rows_of_shard": my_rows, "h2d_bytes_per_step": int(h_z.numel() * 4),
               "d2h_bytes_per_step": int(e2e_rows * out_cols * 4), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "pinned_host": bool(h_out.is_pinned()), "pieces_ms": {k: v / e2e_steps for k, v in pieces.items()},
               "api": "auvi_grid_create_slab + auvi_lattice(host_out) + auvi_grid_destroy",
               "cpu_affinity": (f"{numa[0]}-{numa[-1]} ({len(numa)} cores, NVML GPU-local)" if numa else "unbound")}
        # raw pinned device->host copy bandwidth of this box, for context (the e2e step moves 16x more bytes D2H than H2D)
        try:
            probe_d = torch.empty(1 << 28, dtype=torch.float32, device=dev)
            probe_h = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)
            probe_h.copy_(probe_d); torch.cuda.synchronize()
            t0 = time.perf_counter()
            probe_h.copy_(probe_d, non_blocking=True); torch.cuda.synchronize()
            e2e["pcie_d2h_gbs_raw"] = probe_d.numel() * 4 / (time.perf_counter() - t0) / 1e9
            e2e["pcie_d2h_gbs_in_e2e"] = e2e["d2h_bytes_per_step"] / dt / 1e9
            del probe_d, probe_h
        except Exception:
            pass
        # the same call with a PAGEABLE destination (what a std::vector / numpy caller has), on a bounded row range
        try:
            p_rows = min(e2e_rows, 8192)
            pageable = np.zeros((p_rows, out_cols), dtype=np.float32)             # touched, like a value-initialised vector
            h = C.c_void_p()
            assert lib.auvi_grid_create_slab(h_z.data_ptr(), auvi.F32, n_lat_global, n_lon, in_lo, in_hi - in_lo, *bounds, local, C.byref(h)) == 0
            pp = pageable.ctypes.data
            assert lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, row_lo, row_lo + p_rows, pp) == 0
            t0 = time.perf_counter()
            for _ in range(2):
                assert lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, row_lo, row_lo + p_rows, pp) == 0
            dtp = (time.perf_counter() - t0) / 2
            lib.auvi_grid_destroy(h)
            e2e["pageable_host"] = {"rows": p_rows, "ms": dtp * 1e3, "GBps_d2h": p_rows * out_cols * 4 / dtp / 1e9,
                                    "Mcells_per_s_this_rank": p_rows * out_cols / dtp / 1e6,
                                    "equals_pinned_result": bool(np.array_equal(pageable[:64], h_out[:64].numpy()))}
            del pageable
        except Exception as exc:
            e2e["pageable_host"] = {"unavailable": repr(exc)[:200]}
        # spot check: the host result equals the device-resident result
        chk = slice(e2e_rows // 2, e2e_rows // 2 + 8)
        assert torch.equal(h_out[chk], out[chk, :out_cols].cpu())
        del h_out, h_z

    # ---- informational: other methods / configs ----------------------------------------------------
    extra = {}
    if not args.no_extra:
        ms_b, _ = timed(lambda: step(auvi.BILINEAR), max(3, args.steps // 2), 2)
        extra["bilinear_4x_upsample_f32"] = {"Mcells_per_s": cells_total / (ms_b * 1e-3) / 1e6, "ms": ms_b,
                                             "hbm_frac": ALGO_BYTES_PER_CELL * cells_rank / (ms_b * 1e-3) / 1e9 / peak}
        # latitude-only variant: 4x more rows, the same columns -> 4 + 4/4 = 5 B per output cell
        lat_rows = my_rows
        out_lat = out.view(-1)[: lat_rows * n_lon].view(lat_rows, n_lon)

        def step_lat():
            g.lattice_device(auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, 1, 0, row_lo, row_hi, out_lat.data_ptr(), n_lon, None, stream)

        ms_l, _ = timed(step_lat, max(3, args.steps // 2), 2)
        extra["bicubic_4x_latitude_only_f32"] = {"Mcells_per_s": out_rows_global * n_lon / (ms_l * 1e-3) / 1e6, "ms": ms_l,
                                                 "out_cells_per_gpu": lat_rows * n_lon,
                                                 "hbm_frac": 5.0 * lat_rows * n_lon / (ms_l * 1e-3) / 1e9 / peak}
        if rank == 0:
            extra.update(extra_gap_fill(torch, auvi, dev, stream, peak))
            extra.update(extra_mariana(torch, auvi, local))
            extra.update(extra_grid_a_points(torch, auvi, local))
            extra.update(extra_grid_a_lattice(torch, auvi, dev, stream, peak))

    if dist is not None and not args.no_extra:
        res = extra_gap_fill_sharded(torch, auvi, dist, world, rank, dev, stream, peak)
        if rank == 0:
            extra.update(res)

    # ---- N > 1: the only collective on the path -- gathering output shards to one consumer (not in `value`) ----
    gather = None
    if dist is not None:
        step()                                                      # `out` holds the bicubic result again (extras reused it)
        torch.cuda.synchronize()
        g_rows = min(my_rows, (4 << 30) // (out_ld * 4))            # bounded: at most 4 GiB per rank
        sendbuf = out[:g_rows]
        recv = [torch.empty_like(sendbuf) for _ in range(world)] if rank == 0 else None
        dist.gather(sendbuf, recv, dst=0)                             # warm-up (NCCL over NVLink)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.gather(sendbuf, recv, dst=0)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        g_ms = float(t.item())
        g_bytes = g_rows * out_ld * 4 * (world - 1)
        gather = {"api": "torch.distributed.gather (NCCL)", "rows_per_rank": g_rows, "bytes_into_root": g_bytes, "ms": g_ms,
                  "GBps_into_root": g_bytes / (g_ms * 1e-3) / 1e9,
                  "full_gather_ms_estimate": g_ms * my_rows / g_rows,
                  "Mcells_per_s_gathered_to_root_estimate": cells_total / ((ms_step + g_ms * my_rows / g_rows) * 1e-3) / 1e6}
        if rank == 0:
            assert torch.equal(recv[0], sendbuf)
        # The same gather with NO collective call: every rank's upsample kernel stores its tiles straight into rank 0's
        # buffer over NVLink (peer memory mapped through the C ABI's CUDA-IPC entries, torch's symmetric memory as the
        # fallback; the kernel only sees a pointer).
        # Compute and transfer are one kernel: the 16-byte streaming stores of a tile go to the peer while the next
        # tile is computed.
        try:
            mapping = "cudaIpc (auvi_peer_export / auvi_peer_open)"
            opened = None
            root_view = None
            # rank 0's buffer mapped into every rank through the C ABI (CUDA IPC), no torch object involved; the two
            # collectives below are unconditional so that a failure on one rank cannot leave the others waiting
            import ctypes as C
            lib = auvi.load()
            good = 1.0
            slab = None
            hbuf = (C.c_ubyte * 72)()
            if rank == 0:
                try:
                    slab = torch.empty((world * g_rows, out_ld), dtype=torch.float32, device=dev)
                    if lib.auvi_peer_export(slab.data_ptr(), hbuf) != 0:
                        good = 0.0
                except Exception:
                    good = 0.0
            ht = torch.tensor(list(hbuf), dtype=torch.uint8, device=dev)
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
            flag = torch.tensor([good], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if float(flag.item()) == 0.0:                                              # any rank failed: torch's mapping instead
                import torch.distributed._symmetric_memory as symm_mem
                mapping = "torch symmetric memory"
                if opened is not None:
                    lib.auvi_peer_close(opened)
                opened = None
                slab = symm_mem.empty((world * g_rows, out_ld), dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(slab, dist.group.WORLD)
                root_view = hdl.get_buffer(0, slab.shape, slab.dtype)              # rank 0's buffer, mapped here
                base_ptr = root_view.data_ptr()
            dst = base_ptr + rank * g_rows * out_ld * 4

            def fused():
                g.lattice_device(auvi.CUBIC, auvi.AXIS_EXPANDED, FACTOR, FACTOR, 0, row_lo, row_lo + g_rows, dst, out_ld, None, stream)

            fused()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fused()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            barrier()
            f_ms = float(t.item())
            ok = True
            if rank == 0:
                for r in range(world):
                    ok = ok and torch.equal(slab[r * g_rows:(r + 1) * g_rows, :out_cols], recv[r][:, :out_cols])
            gather["fused_peer_store"] = {
                "what": "upsample kernel writes its rows into rank 0's buffer over NVLink (no NCCL call, no staging copy)",
                "ms_compute_plus_transfer": f_ms, "ms_kernel_then_nccl_gather": ms_step * g_rows / my_rows + g_ms,
                "GBps_into_root": g_bytes / (f_ms * 1e-3) / 1e9, "equals_nccl_gather": bool(ok), "peer_mapping": mapping}
            if opened is not None:
                lib.auvi_peer_close(opened)
            del slab, root_view
        except Exception as exc:                                                    # no peer mapping on this box
            gather["fused_peer_store"] = {"unavailable": repr(exc)[:200]}
        del recv

    cpu = None
    os.sched_setaffinity(0, all_cores)                               # the CPU baseline gets every core of the box back
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, kind, cores, ms_cpu, sample = cpu_reference_rate(CPU_SAMPLE_ROWS, threads)
        cpu = {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "gather": gather, "extra": extra}
        print(json.dumps(line), flush=True)
    g.close()
    if dist is not None:
        dist.destroy_process_group()


def extra_gap_fill(torch, auvi, dev, stream, peak):
    """BASELINE configs[4] on one GPU: FP32 grid at 70 % mask, full-grid gap fill -- every method at 16384^2 and
    the IDW headline of that config at the full 65536^2 (17.2 GB in + 17.2 GB out on one B200)."""
    res = {}
    for n, methods in ((16384, (("idw", auvi.IDW), ("nn", auvi.NN), ("kriging", auvi.KRIGING), ("bilinear", auvi.BILINEAR),
                                ("nearest4_mean(cubic fallback)", auvi.CUBIC))), (65536, (("idw", auvi.IDW),))):
        z = synth_grid_device(torch, n, n, 0, n, dev)
        g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
                      min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0, device=dev.index)
        g.mask_hash(0.70, seed=42, count=False, stream=stream)   # counter-hash mask drawn on the device (csrc/ingest.cu)
        out = torch.empty((n, n), dtype=torch.float32, device=dev)
        for name, meth in methods:
            fn = lambda: g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, stream)
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            res[f"gap_fill_70pct_{name}_{n}sq_f32"] = {"Mcells_per_s": n * n / (ms * 1e-3) / 1e6, "ms": ms,
                                                        "hbm_frac": 8.0 * n * n / (ms * 1e-3) / 1e9 / peak,
                                                        "nan_left": int(torch.isnan(out).sum().item())}
        g.close()
        del z, out
        torch.cuda.empty_cache()
    return res


def extra_gap_fill_sharded(torch, auvi, dist, world, rank, dev, stream, peak):
    """BASELINE configs[4] as it is stated: ONE 65536^2 FP32 grid at 70 % mask, IDW gap fill, output rows sharded over the
    ranks (strong scaling).  Each rank holds its rows + a halo of the known-point input, draws its part of the one global
    mask from the counter hash (no communication), fills its rows; the time is the max over ranks between barriers."""
    import shard
    n = 65536
    plan = shard.plan_rows(n, 1, world, rank)
    lo, hi, in_lo, in_hi = plan.row_lo, plan.row_hi, plan.in_lo, plan.in_hi
    z = synth_grid_device(torch, n, n, in_lo, in_hi, dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=in_lo, rows=in_hi - in_lo,
                             keep=z), min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0, device=dev.index)
    g.mask_hash(0.70, seed=42, count=False, stream=stream)
    out = torch.empty((hi - lo, n), dtype=torch.float32, device=dev)
    res = {}
    for name, meth in (("idw", auvi.IDW), ("nn", auvi.NN)):
        fn = lambda: g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, lo, hi, out.data_ptr(), n, None, stream)
        fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nan_left = torch.tensor([int(torch.isnan(out).sum().item())], device=dev, dtype=torch.int64)
        dist.all_reduce(nan_left)
        ms = float(t.item())
        res[f"gap_fill_70pct_{name}_65536sq_f32_sharded_x{world}"] = {
            "Mcells_per_s": n * n / (ms * 1e-3) / 1e6, "ms": ms, "scaling": "strong", "rows_per_rank": hi - lo,
            "halo_rows": shard.HALO, "hbm_frac_per_gpu": 8.0 * (hi - lo) * n / (ms * 1e-3) / 1e9 / peak,
            "nan_left": int(nan_left.item())}
    g.close()
    del z, out
    torch.cuda.empty_cache()
    return res


def extra_grid_a_points(torch, auvi, local):
    """The reference's Grid-A benchmark row (results/grid_A_runtimes_averaged.csv:8): 5,000,000 random query points
    on the 4000 x 3200 synthetic grid through the Point-list API with host buffers (what GridD::batch* calls)."""
    from oracle import binding as ob          # synthetic-field generator only
    z = ob.synth_grid(3200, 4000, csv_round=False)
    g = auvi.Grid(z, *BOUNDS, device=local)
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


def extra_grid_a_lattice(torch, auvi, dev, stream, peak):
    """BASELINE configs[0] at the generator's shipped size: the 4000 x 3200 FP64 Grid A, 2x expanded lattice
    (7999 x 6399 = 51.2 M cells, test_interpolation.cpp:283-297), device-resident, every method."""
    from oracle import binding as ob          # synthetic-field generator only
    z = torch.from_numpy(ob.synth_grid(3200, 4000, csv_round=False)).to(dev)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F64, n_lat=3200, n_lon=4000, ld=4000, row0=0, rows=3200, keep=z),
                  min_lon=BOUNDS[0], max_lon=BOUNDS[1], min_lat=BOUNDS[2], max_lat=BOUNDS[3], device=dev.index)
    rows, cols = g.lattice_dims(auvi.AXIS_EXPANDED, 2, 2)
    out = torch.empty((rows, 8000), dtype=torch.float64, device=dev)
    res = {}
    for name, meth in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC), ("kriging", auvi.KRIGING), ("nn", auvi.NN),
                       ("idw", auvi.IDW)):
        fn = lambda: g.lattice_device(meth, auvi.AXIS_EXPANDED, 2, 2, 0, 0, rows, out.data_ptr(), 8000, None, stream)
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[f"grid_a_2x_lattice_f64_{name}"] = {"Mcells_per_s": rows * cols / (ms * 1e-3) / 1e6, "ms": ms,
                                                 "hbm_frac": 10.0 * rows * cols / (ms * 1e-3) / 1e9 / peak}
    g.close()
    return res


def extra_mariana(torch, auvi, local):
    """BASELINE configs[1]: Mariana tile at 50 % removal through the Point-list API (host buffers), all
    methods, with RMSE against the unmasked truth computed on the device."""
    from oracle import binding as ob          # fixture loader only (tile + seed-42 mask), not the computation
    case = ob.masked_case("mariana", 0.5)
    g = auvi.Grid(case["z"], *case["bounds"], device=local)
    d_truth = torch.from_numpy(case["truth"]).cuda()
    res = {}
    for name, meth in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC), ("kriging", auvi.KRIGING),
                       ("nn", auvi.NN), ("idw", auvi.IDW), ("bilinear_search(opt-in)", auvi.BILINEAR_SEARCH)):
        g.interp_points(meth, case["pts"])
        t0 = time.perf_counter()
        for _ in range(5):
            est = g.interp_points(meth, case["pts"])
        dt = (time.perf_counter() - t0) / 5
        d_est = torch.from_numpy(est).cuda()
        mae, rmse, mx, n_nan = auvi.error_metrics_device(d_truth.data_ptr(), d_est.data_ptr(), auvi.F64, est.size)
        res[f"mariana50_points_{name}"] = {"Mpts_per_s_e2e": est.size / dt / 1e6, "ms_e2e": dt * 1e3,
                                           "kernel_ms": g.last_kernel_ms, "rmse_m": rmse, "mae_m": mae, "max_m": mx,
                                           "n_nan": n_nan, "n": int(est.size)}
    g.close()
    return res


if __name__ == "__main__":
    main()
