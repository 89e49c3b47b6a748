#!/usr/bin/env python
"""bench.py — headline benchmark of the cuberille hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--period P]

One "step" = one pass of the hot path (cub_count + cub_emit through the C-ABI: classify -> count +
look-back scan -> emit points + quads) over a synthetic float32 gyroid that is already resident in
HBM.  N > 1 (launched by torchrun, one rank per GPU): the image is split into z-slabs with a 2-slice
halo, every rank runs the same kernels on its slab, the only exchange is an NCCL all-gather of the
per-rank (points, cells) counts that turns local ids into global ids (weak scaling: S^3 voxels per
GPU, the image is S x S x (S*N)).  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference filter (oracle/, single-threaded like
GenerateData itself) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="voxels per axis per GPU")
    ap.add_argument("--period", type=float, default=128.0, help="gyroid period in voxels")
    ap.add_argument("--field", default="gyroid", choices=["gyroid", "marschner_lobb", "blobs"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time the all-gather of the meshes (N > 1)")
    return ap.parse_args()


FIELD_KIND = {"gyroid": 0, "marschner_lobb": 1, "blobs": 2}
FIELD_ISO = {"gyroid": 0.0, "marschner_lobb": 0.5, "blobs": 0.5}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s in sm if s > 0.5 * max(mx or [1])] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample_rate(vol, iso, triangles, project, params):
    import oracle_py as O
    t = time.perf_counter()
    m = O.cuberille(vol, iso, triangles=triangles, project=project, mode=O.LITERAL, **params)
    dt = time.perf_counter() - t
    return dt, m


def cpu_baseline_from_volume(get_slab, nx, ny, nz_max, iso, seconds):
    """time the oracle on a bounded z-sample of the workload: calibrate on 8 slices, then size the
    sample for about `seconds` of CPU work."""
    vol = get_slab(8)
    dt, _ = oracle_sample_rate(vol, iso, False, False, {})
    rate = vol.size / max(dt, 1e-6)
    nz = int(min(nz_max, max(8, seconds * rate / (nx * ny))))
    vol = get_slab(nz)
    dt, m = oracle_sample_rate(vol, iso, False, False, {})
    return {"value": vol.size / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
            "sample": f"{nx}x{ny}x{nz} z-sub-slab of the same volume (same bytes, downloaded from the GPU), "
                      f"quads, no projection, {dt:.2f} s, {m.cells.shape[0]} quads",
            "mfaces_per_s": m.cells.shape[0] / dt / 1e6, "seconds": dt,
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    """the reference's own CPU implementation of the path (the oracle port: ITK is not installable here,
    DESIGN.md §6), single-threaded like GenerateData (txx:136-206), on a bounded sample."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_py as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    S = args.size
    # same field as the CUDA arm, generated on the host (numpy float32); the sample is a z-sub-slab
    k = np.float32(2.0 * np.pi / args.period)

    def slab(nz):
        z, y, x = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(S, dtype=np.float32), np.arange(S, dtype=np.float32),
                              indexing="ij")
        g = (np.sin(k * x) * np.cos(k * y) + np.sin(k * y) * np.cos(k * z) + np.sin(k * z) * np.cos(k * x)).astype(np.float32)
        g[0] = -2; g[-1] = -2; g[:, 0] = -2; g[:, -1] = -2; g[:, :, 0] = -2; g[:, :, -1] = -2
        return np.ascontiguousarray(g)

    cal = slab(8)
    dt, _ = oracle_sample_rate(cal, 0.0, False, False, {})
    rate = cal.size / max(dt, 1e-6)
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    nz = int(min(S, max(8, per_step * rate / (S * S))))
    vol = slab(nz)
    times, faces = [], 0
    for i in range(args.warmup + args.steps):
        dt, m = oracle_sample_rate(vol, 0.0, False, False, {})
        faces = m.cells.shape[0]
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    v = vol.size / t / 1e9
    line = {
        "impl": "reference", "metric": "Gvoxels/s", "value": v, "unit": "Gvoxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "mfaces_per_s": faces / t / 1e6,
        "cpu_baseline": {"value": v, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
                         "sample": f"{S}x{S}x{nz} z-sub-slab of the gyroid per step (CPU restatement of txx:59-498, "
                                   f"single-threaded like the reference; ITK itself is not installable offline)",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": v, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    S = args.size
    nz = S * n if args.scaling == "weak" else S
    return {"workload": f"synthetic {args.field} {S}x{S}x{nz} float32"
                        + (f" period {args.period:g} voxels" if args.field == "gyroid" else "")
                        + f", iso {FIELD_ISO[args.field]}, quads, no projection, uint32 ids"
                        + (f", z-slabs over {n} GPUs (2-slice halo)" if n > 1 else ""),
            "voxels": S * S * nz, "l2_policy": "inputs larger than L2 (4.3 GB volume per GPU vs 126 MB L2), no flush needed"
            if S >= 512 else "small input: L2-resident"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = world

    P = importlib.import_module("midas-journal-740_b200")
    capi = P.capi
    # a real (non-default) stream: torch's default stream has handle 0, which the C-ABI reads as "create your
    # own stream" - and then torch events would not bracket the library's work
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h = capi.Handle(local_rank, stream.cuda_stream)

    S = args.size
    image_nz = S * N if args.scaling == "weak" else S
    slab = P.slabs.plan_slabs(image_nz, N, halo=2)[rank]
    own0, own1, lo, hi = slab.own_z0, slab.own_z1, slab.local_z0, slab.local_z1
    kind = FIELD_KIND[args.field]
    p0, p1 = (args.period, 1.0) if args.field == "gyroid" else ((48.0, 1.0) if args.field == "blobs" else (0.0, 0.0))
    h.generate(kind, (S, S, hi - lo), (S, S, image_nz), lo, p0, p1)
    if N > 1:
        h.set_slab(image_nz, lo, own0, own1)
    iso = FIELD_ISO[args.field]
    prm = capi.default_params()
    prm.iso_value, prm.generate_triangles, prm.project_vertices = iso, 0, 0

    last_counts = [None]

    side_stream = torch.cuda.Stream() if N > 1 else None

    def step(params=prm, handle=None):
        hh = handle or h
        n_pts, n_quads = hh.count(params)
        if N > 1:
            hh.emit_vertices()   # needs no id base: runs while the counts are exchanged on a side stream
        counts = P.slabs.all_gather_counts(n_pts, n_quads, dev, side_stream)   # the only exchange of the data path
        last_counts[0] = counts
        cells_per_quad = 2 if params.generate_triangles else 1
        pbase, cbase = P.slabs.exclusive_bases(counts, rank)
        hh.set_id_base(pbase, cbase * cells_per_quad)
        hh.emit(4)
        return n_pts, n_quads, (sum(c[0] for c in counts), sum(c[1] for c in counts))

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = h.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if N > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, h.launch_count() - l0, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches, (n_pts, n_quads, tot) = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    voxels_total = S * S * (own1 - own0)
    vt = torch.tensor([voxels_total], dtype=torch.int64, device=dev)
    if N > 1:
        dist.all_reduce(vt)
    voxels_total = int(vt.item())
    gvox = voxels_total / (ms_step * 1e-3) / 1e9
    mfaces = tot[1] / (ms_step * 1e-3) / 1e6

    peak, peak_src = measured_peaks()

    # ---- per-kernel device times (CUDA events inside the library, separate passes) ---------------
    h.enable_timing(True)
    kt = {"classify": [], "count_scan": [], "scan_only": [], "emit": []}
    for _ in range(max(3, min(args.steps, 10))):
        step()
        t = h.timings()
        for k in kt:
            kt[k].append(t[k])
    h.enable_timing(False)
    kavg = {k: sum(v) / len(v) for k, v in kt.items()}
    own_vox = S * S * (own1 - own0)
    local_vox = S * S * (hi - lo)
    alg_k1 = local_vox * 4  # K1 reads every voxel of the local buffer once
    k1_gbs = alg_k1 / (kavg["classify"] * 1e-3) / 1e9
    alg_pipe = own_vox * 4 + n_pts * 12 + n_quads * 16
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("k_classify_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_classify<float>", "achieved": k1_gbs, "peak": peak, "unit": "GB/s",
                "frac": k1_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_k1, "avg_launch_ms": kavg["classify"],
                "kernel_ms": kavg,
                "pipeline": {"algorithmic_bytes_per_step": alg_pipe, "achieved": alg_pipe / (ms_step * 1e-3) / 1e9,
                             "frac": alg_pipe / (ms_step * 1e-3) / 1e9 / peak,
                             "input_only_frac": own_vox * 4 / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- extras: raster vertex order; the full default filter (triangles + projection) -------------
    extras = {}
    if not args.no_extras:
        prm1 = capi.default_params()
        prm1.iso_value, prm1.generate_triangles, prm1.project_vertices = iso, 0, 0
        prm1.vertex_order = capi.ORDER_RASTER
        ms1, _, (_, _, tot1) = timed(lambda: step(prm1), max(3, args.steps // 2), 2)
        extras["raster_vertex_order"] = {"ms_per_step": ms1, "gvoxels_per_s": voxels_total / (ms1 * 1e-3) / 1e9,
                                         "mfaces_per_s": tot1[1] / (ms1 * 1e-3) / 1e6,
                                         "note": "CUB_ORDER_RASTER: same mesh up to vertex renumbering (canonical ordering)"}
        prm2 = capi.default_params()
        prm2.iso_value, prm2.generate_triangles, prm2.project_vertices = iso, 1, 1
        prm2.surface_distance_threshold = 0.01 if args.field == "gyroid" else 0.005
        ms2, _, (np2, nq2, tot2) = timed(lambda: step(prm2), max(2, args.steps // 3), 1)
        h.enable_timing(True)
        step(prm2)
        t2 = h.timings()
        h.enable_timing(False)
        extras["triangles_projection"] = {"ms_per_step": ms2, "gvoxels_per_s": voxels_total / (ms2 * 1e-3) / 1e9,
                                          "mtriangles_per_s": 2 * tot2[1] / (ms2 * 1e-3) / 1e6,
                                          "mvertices_per_s_project_kernel": np2 / (t2["project"] * 1e-3) / 1e6 if t2["project"] else None,
                                          "kernel_ms": t2, "threshold": prm2.surface_distance_threshold}

    # ---- e2e: host volume in, host mesh out, through the public call sequence ----------------------
    e2e = None
    if not args.no_e2e:
        vol_host = torch.empty((hi - lo, S, S), dtype=torch.float32, pin_memory=True)
        h._check(h._L.cub_download_volume(h._h, vol_host.data_ptr(), vol_host.numel() * 4))
        pts_host = torch.empty((max(n_pts, 1), 3), dtype=torch.float32, pin_memory=True)
        cells_host = torch.empty((max(n_quads, 1), 4), dtype=torch.int32, pin_memory=True)
        he = capi.Handle(local_rank, stream.cuda_stream)

        def e2e_step():
            he.set_volume_ptr(vol_host.data_ptr(), np.float32, (S, S, hi - lo), capi.MEM_HOST)
            if N > 1:
                he.set_slab(image_nz, lo, own0, own1)
            a, b, _ = step(prm, he)
            he.fetch_into(pts_host.data_ptr(), cells_host.data_ptr())
            return a, b, None

        ms_e, _, _ = timed(e2e_step, max(2, min(args.steps, 5)), 1)
        e2e = {"value": voxels_total / (ms_e * 1e-3) / 1e9, "unit": "Gvoxels/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": int(vol_host.numel() * 4), "d2h_bytes_per_step": int(n_pts * 12 + n_quads * 16),
               "mfaces_per_s": tot[1] / (ms_e * 1e-3) / 1e6, "mode": "one slab per GPU: copy in, run, copy out"}
        if N == 1:
            # streamed: the volume travels through 3 handles (3 streams) as 16 z-slabs, so the PCIe copies in both
            # directions and the kernels overlap (slabs.run_streamed: the same public calls, ids stay global)
            # The volume fits in HBM beside its mesh, so it is copied once (16 consecutive pieces on a copy stream)
            # into a device buffer and the handles borrow windows of it: no halo slice crosses PCIe twice.
            sts = [torch.cuda.Stream() for _ in range(3)]
            hs = [capi.Handle(local_rank, st.cuda_stream) for st in sts]
            vol_dev = torch.empty(vol_host.numel() * 4, dtype=torch.uint8, device=dev)

            def e2e_streamed():
                a, b = P.slabs.run_streamed(hs, vol_host.data_ptr(), np.float32, (S, S, hi - lo), prm, 16,
                                            pts_host.data_ptr(), cells_host.data_ptr(), device_volume=vol_dev, streams=sts)
                assert (a, b) == (n_pts, n_quads)
                return a, b, None

            ms_s, _, _ = timed(e2e_streamed, max(2, min(args.steps, 5)), 1)
            for x in hs:
                x.close()
            del vol_dev
            unstreamed = e2e
            e2e = {"value": voxels_total / (ms_s * 1e-3) / 1e9, "unit": "Gvoxels/s", "ms_per_step": ms_s,
                   "h2d_bytes_per_step": int(vol_host.numel() * 4),
                   "d2h_bytes_per_step": int(n_pts * 12 + n_quads * 16), "mfaces_per_s": tot[1] / (ms_s * 1e-3) / 1e6,
                   "mode": "streamed: volume copied once in 16 pieces, 16 z-slabs through 3 handles / streams borrow windows of it",
                   "unstreamed": unstreamed}
        he.close()
        del vol_host

    # ---- optional: all-gather of the meshes over NVLink (reported separately) -----------
    gather = None
    if args.gather and N > 1:
        step(prm)  # the handle holds the headline mesh again (quads), whatever the extras ran last
        maxp, maxq = max(c[0] for c in last_counts[0]), max(c[1] for c in last_counts[0])
        # padded all-gather straight from the result buffers (uneven sizes -> pad to the max)
        src_p = torch.zeros(maxp * 3, dtype=torch.float32, device=dev)
        src_c = torch.zeros(maxq * 4, dtype=torch.int32, device=dev)
        h.fetch_into(src_p.data_ptr(), src_c.data_ptr(), 0, capi.MEM_DEVICE)
        dst_p = torch.empty(N * maxp * 3, dtype=torch.float32, device=dev)
        dst_c = torch.empty(N * maxq * 4, dtype=torch.int32, device=dev)

        def gstep():
            dist.all_gather_into_tensor(dst_p, src_p)
            dist.all_gather_into_tensor(dst_c, src_c)
            return 0, 0, None
        ms_g, _, _ = timed(gstep, 3, 1)
        gather = {"ms": ms_g, "bytes_per_rank": int(maxp * 12 + maxq * 16), "collective": "nccl all_gather (padded)"}

    # ---- cpu baseline (rank 0, N == 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        full = None

        def get_slab(nz):
            nonlocal full
            if full is None:
                full = h.download_volume()
            z0 = (full.shape[0] - nz) // 2
            v = np.ascontiguousarray(full[z0:z0 + nz])
            return v
        cpu = cpu_baseline_from_volume(get_slab, S, S, hi - lo, iso, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": "Gvoxels/s", "value": gvox, "unit": "Gvoxels/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, N),
            "mfaces_per_s": mfaces, "n_points": tot[0], "n_quads": tot[1],
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "extras": extras,
        }
        if gather:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    h.close()
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON result: anything a library prints there meanwhile (NCCL's version
    # banner, for one) is sent to stderr instead
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)
    _result = sys.stdout = os.fdopen(_stdout_fd, "w")
    try:
        main()
    finally:
        _result.flush()
