#!/usr/bin/env python
"""bench.py — sim timesteps/s of the floe-interaction hot path (BASELINE.json metric).

One "step" = one reference timestep restricted to the replaced calls (simulation.jl:94-170):
add_ghosts! -> timestep_collisions! -> ghost removal -> timestep_coupling! -> timestep_floe_properties!
on a synthetic Voronoi-packed floe field; coupling runs EVERY step here (the
reference default is every 10th), so the number is a lower bound for default settings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--floes 100000] [--npoints 1000] [--walls collision|periodic|shear]

N = 1: BASELINE config 3 (100k floes, collisions + ocean/atmosphere coupling) on one B200.
N > 1 (torchrun, one rank per GPU): weak scaling, every rank owns a slab of `--floes` floes; the slab data plane
(partition, per-step halo update over peer memory, staleness test, rebuild) is the library's sz_slab_* API.
--impl reference: the CPU oracle (the reference is pure Julia and cannot run in this image, so the
reference arm is the oracle "port") on the host cores of rank 0, on a bounded sample of the workload.
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
import szload  # noqa: E402,F401  (registers the package directory `subzero.jl_b200` as subzero_jl_b200)

# `value` is FLOE-NORMALISED: simulated timesteps/s x floes_total / 100 000, i.e. timesteps/s of a 100k-floe field
# (BASELINE config 3).  At N = 1 with the default 100k floes it IS timesteps/s; under weak scaling (100k floes per
# GPU, N GPUs) it is the whole-job aggregate that grows with N, so value_N / (N value_1) is the parallel efficiency.
METRIC = "sim timesteps/sec per 100k floes (collisions + one-way ocean/atmosphere coupling + state update; steps/s x floes_total/100000)"
NORM_FLOES = 100000.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--floes", type=int, default=100000, help="floes per GPU")
    ap.add_argument("--npoints", type=int, default=1000, help="Monte-Carlo draws per floe (about 59 %% are kept)")
    ap.add_argument("--walls", default="collision", choices=["collision", "periodic", "shear"])
    ap.add_argument("--scale", type=float, default=1.01)
    ap.add_argument("--flow", default="random", choices=["random", "converging"],
                    help="initial floe velocities: U(-0.1, 0.1) m/s or a flow converging on the domain centre (BASELINE config 5)")
    ap.add_argument("--cpu-sample", type=int, default=25000, help="floes of the cpu_baseline sample field")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the one-step CUDA-vs-oracle comparison on the benchmark field")
    ap.add_argument("--two-way", action="store_true", help="also turn on two-way coupling (not part of the headline config)")
    ap.add_argument("--skin", type=float, default=6000.0, help="halo-list skin in metres (N > 1): lists stay valid while floes moved < skin/2")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 5 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index, self.sm, self.bits, self.run, self.t, self.smax = index, [], 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is None:
            return
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.run = False
        self.t.join(timeout=2)
        reasons = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                "samples": len(self.sm)}


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off (torchrun does not bind): the pinned host arrays of the
    end-to-end leg are then first-touched on that node and the copies do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # nvml prints an 8-digit domain, sysfs a 4-digit one
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        pass
    return None


def pin_floe_arrays(fa):
    """Move every array of a FloeArrays into pinned host memory (torch allocator)."""
    import torch
    keep = []
    for name, v in list(fa.__dict__.items()):
        if isinstance(v, np.ndarray) and v.size > 0:
            v = np.ascontiguousarray(v)
            t = torch.empty(v.shape, dtype=torch.from_numpy(np.empty(0, dtype=v.dtype)).dtype, pin_memory=True)
            t.numpy()[...] = v
            keep.append(t)
            setattr(fa, name, t.numpy())
    fa._pinned = keep
    return fa


def dyn_bytes(fa):
    """Bytes of the dynamic state crossing PCIe in one direction (scalars + tensors + status + rings)."""
    from subzero_jl_b200 import capi
    b = 0
    for name in capi.DOUBLE_FIELDS:
        b += getattr(fa, name).nbytes
    return b + fa.status_tag.nbytes + fa.vert_xy.nbytes


def oracle_steps_per_s(args, n_sample, steps, warmup, threads=0):
    """Time the CPU oracle on a field of n_sample floes with the same generator/statistics;
    returns (steps/s on the sample, threads used, counts)."""
    from subzero_jl_b200 import synth
    from oracle import szo
    lib = szo.oracle()
    f = synth.make_field(n_sample, scale=args.scale, walls=args.walls, npoints=args.npoints, flow=args.flow)
    threads = threads or os.cpu_count()  # explicit: torchrun exports OMP_NUM_THREADS=1
    h = synth.setup_handle(f, lib, threads=threads)
    for t in range(warmup):
        h.step(t, True)
    t0 = time.perf_counter()
    for t in range(steps):
        h.step(warmup + t, True)
    dt = time.perf_counter() - t0
    c = h.counts()
    h.close()
    return steps / dt, threads, c


def parity_vs_oracle(args, prod, device, world):
    """One timestep of the benchmark field (this rank's tile at N > 1) on the CUDA library and on the CPU oracle, phase by
    phase, from the same initial state: pair sets and interaction rows bit-exact, state within 1e-9 (BASELINE north_star).
    Runs AFTER the timed region; the oracle is the checker, not the thing measured."""
    from subzero_jl_b200 import synth
    from oracle import szo
    try:
        f = synth.make_field(args.floes, scale=args.scale, walls=args.walls if world == 1 else "collision", npoints=args.npoints,
                             seed=args.floes, flow=args.flow)
        hg = synth.setup_handle(f, prod, device=device)
        ho = synth.setup_handle(f, szo.oracle(), threads=os.cpu_count())
        for h in (hg, ho):
            h.add_ghosts()
            h.step_collisions()
        pairs_equal = all(np.array_equal(hg.pairs(w), ho.pairs(w)) for w in range(4))
        og, rg = hg.interactions()
        oo, ro = ho.interactions()
        rows_equal = bool(np.array_equal(og, oo) and np.array_equal(rg, ro))
        for h in (hg, ho):
            h.remove_ghosts()
            h.step_coupling()
            h.step_floe_properties(0)
        a, b = hg.download_floes(mc=False), ho.download_floes(mc=False)
        worst, worst_name = 0.0, ""
        for name in ("centroid_x", "centroid_y", "alpha", "u", "v", "xi", "fxOA", "fyOA", "trqOA", "collision_force", "collision_trq",
                     "overarea", "stress_accum", "stress_instant", "strain", "vert_xy", "mass", "moment", "height"):
            x, y = np.asarray(getattr(a, name)), np.asarray(getattr(b, name))
            scale = max(float(np.sqrt(np.mean(y * y))), 1e-300)
            err = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), scale))) if x.size else 0.0
            if err > worst:
                worst, worst_name = err, name
        c = hg.counts()
        out = {"pairs_equal": bool(pairs_equal), "rows_equal": rows_equal, "max_rel_err": worst, "max_rel_err_field": worst_name,
               "tolerance": 1e-9, "ok": bool(pairs_equal and rows_equal and worst < 1e-9),
               "n_floes": int(f.floes.n), "n_candidates": c["n_candidates"], "n_overlap": c["n_overlap"], "n_rows": c["n_rows"],
               "how": "one timestep phase by phase on the benchmark field, CUDA library vs CPU oracle (candidate / filtered / overlap / "
                      "fuse pair lists and interaction rows compared bit for bit, state fields relative to max(|ref|, rms))"}
        hg.close()
        ho.close()
        return out
    except Exception as e:  # evidence, not part of the measurement
        return {"error": repr(e)}


def run_reference(args, rank):
    if rank != 0:
        return
    # calibrate on a small field, then size the sample so that K + W steps take about two minutes
    n0 = min(5000, args.floes)
    sps0, cores, _ = oracle_steps_per_s(args, n0, 1, 1)
    budget = 120.0
    per_floe = 1.0 / (sps0 * n0)
    n_sample = int(min(args.floes, max(n0, budget / ((args.steps + args.warmup) * per_floe))))
    n_sample = max(1000, (n_sample // 1000) * 1000)
    sps, cores, c = oracle_steps_per_s(args, n_sample, args.steps, args.warmup)
    world = max(1, args.gpus)
    total = args.floes * world
    steps_per_s = sps * n_sample / total  # O(N) path: steps/s scale inversely with the floe count
    value = steps_per_s * total / NORM_FLOES  # the same floe-normalised unit as the GPU arm
    sample = ("oracle port (OpenMP, %d threads) on a %d-floe field of the same generator; steps/s scaled by %d/%d "
              "to the %d-floe workload (%d GPUs x %d floes), then floe-normalised like the GPU arm"
              % (cores, n_sample, n_sample, total, total, world, args.floes))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world), "floes_total": total, "steps_per_s": steps_per_s,
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference (Subzero.jl) is pure Julia; no julia binary exists in this image, so the reference arm "
                "is the CPU oracle restating its algorithm (O(N) grid broad phase instead of the reference's O(N^2) loop)",
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "synthetic Voronoi-packed %d floes per GPU (BASELINE config 3), scale %.2f dense contacts, "
                        "%s walls, %s initial flow, collisions + one-way ocean/atmosphere coupling every step + state update"
                        % (args.floes, args.scale, args.walls, args.flow),
            "floes_per_gpu": args.floes, "mc_draws_per_floe": args.npoints, "coupling_every": 1, "dt_s": 10,
            "parallelism": "1 GPU" if world == 1 else "%d spatial slabs, one rank per GPU" % world,
            "l2": "Monte-Carlo points (%.2f GB per GPU at 100k floes) exceed the 126 MB L2; no explicit flush" % 0.95}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import szload  # noqa: F401
    from subzero_jl_b200 import capi, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created: send fd 1 to stderr until then, so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    prod = capi.product()

    sl = None
    gloo = None
    if world == 1:
        f = synth.make_field(args.floes, scale=args.scale, walls=args.walls, npoints=args.npoints, seed=args.floes, flow=args.flow)
        h = synth.setup_handle(f, prod, device=local_rank, two_way_coupling_on=int(args.two_way))
        fa0 = f.floes
    else:
        # weak scaling: rank r generates the tile [r L, (r+1) L) x [0, L) and hands it to sz_slab_build; the library
        # finds the neighbours' boundary floes (halo copies), wires the peer-memory arenas (cudaIpc) and from then on
        # every sz_slab_step pushes this rank's boundary floes straight into its neighbours' stores
        from subzero_jl_b200 import slab
        gloo = dist.new_group(backend="gloo")  # byte transport of the library's set-up / rebuild messages
        tile = synth.make_field(args.floes, scale=args.scale, walls="collision", npoints=args.npoints, seed=args.floes + rank)
        slab.shift_x(tile.floes, rank * tile.L)
        tile.floes.id = rank * args.floes + np.arange(1, args.floes + 1, dtype=np.int64)  # unique over all tiles (collisions.jl:751-758)
        if args.flow == "converging":  # towards the centre of the whole (world x 1 tiles) domain
            tile.floes.u = -0.2 * (tile.floes.centroid_x - 0.5 * world * tile.L) / (world * tile.L)
            tile.floes.v = -0.2 * (tile.floes.centroid_y - 0.5 * tile.L) / tile.L
        ew = "shear" if args.walls in ("shear", "periodic") else "collision"
        f = synth.tiled_model(tile, world, ew)
        sl = slab.Slab(prod, f, world, rank=rank, group=gloo, devices=[local_rank], skin=args.skin, device=local_rank)
        sl.set_edges(slab.tile_edges(world, tile.L, ew == "shear"))
        sl.build([tile.floes], [rank * args.floes + np.arange(args.floes, dtype=np.int64)])
        h = sl.handles[0]
        fa0 = tile.floes
    N, M, V = fa0.n, int(fa0.mc_offsets[-1]), int(fa0.vert_offsets[-1])
    rebuild_s, polls = [], []

    # staleness poll without a stall: the 8-byte MAX all-reduce of step t is enqueued asynchronously (NCCL's own stream) and
    # read ONE step later, when it has long finished — a blocking poll cost ~0.9 ms per 25 steps at N = 8 (r3m).  Every
    # rank reads the same reduced value at the same step, so the collective rebuild is entered by all of them.
    poll = {"h_in": None, "dev": None, "h_out": None, "ev": None, "t": -1}
    if sl is not None:
        poll["h_in"] = torch.zeros(1, dtype=torch.float64).pin_memory()
        poll["h_out"] = torch.zeros(1, dtype=torch.float64).pin_memory()
        poll["dev"] = torch.zeros(1, dtype=torch.float64, device="cuda")
        poll["ev"] = torch.cuda.Event()

    def maybe_rebuild(t):
        # the lists are valid while no owned floe travelled more than skin / 2; polled every 25 steps against 0.4 skin
        # (the library keeps the displacement current without a synchronisation), so floes may move another skin / 10
        # between a poll and the rebuild it triggers (1.2 m/s at the default skin)
        if sl is None:
            return
        if t % 25 == 24:
            poll["h_in"][0] = sl.max_displacement()
            poll["dev"].copy_(poll["h_in"], non_blocking=True)
            dist.all_reduce(poll["dev"], op=dist.ReduceOp.MAX, async_op=True).wait()  # stream-level wait, not a host wait
            poll["h_out"].copy_(poll["dev"], non_blocking=True)
            poll["ev"].record()
            poll["t"] = t
        elif poll["t"] >= 0 and t == poll["t"] + 1:
            poll["ev"].synchronize()
            d = float(poll["h_out"][0])
            polls.append((poll["t"], d))
            poll["t"] = -1
            if d > 0.4 * args.skin:
                t0 = time.perf_counter()
                sl.rebuild()
                rebuild_s.append(time.perf_counter() - t0)

    def do_step(t):
        if sl is not None:
            sl.step(t, True)  # unpack the neighbours' records, step, push mine: one C call, no NCCL on the data path
            maybe_rebuild(t)
        else:
            h.step(t, True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ---------------------------------------------------------------
    for t in range(args.warmup):
        do_step(t)
    if world > 1:  # the collectives the timed region uses (displacement poll) have run once before it starts
        d = torch.zeros(1, dtype=torch.float64, device="cuda")
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
    # everything that takes host time (NVML initialisation of the clock sampler: milliseconds) happens BEFORE the barrier:
    # a rank that enters the timed loop late makes its neighbours wait for its boundary floes inside THEIR timed region
    sampler = ClockSampler(local_rank)
    phase = np.zeros(8)
    launches = 0
    step_wall = np.zeros(args.steps)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    tp = t0
    for t in range(args.steps):
        do_step(args.warmup + t)
        ms = h.timings_raw()
        phase += ms
        launches += int(ms[7])
        tn = time.perf_counter()
        step_wall[t] = tn - tp
        tp = tn
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    c = h.counts()
    dev_ms = phase[6] / args.steps  # CUDA events on the library's stream, first to last kernel of a step
    t_rank = torch.tensor([wall, phase[6] / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_rank, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(t_rank[0]), float(t_rank[1])
    wall_ranks = [wall]
    if world > 1:
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([wall], dtype=torch.float64, device="cuda"))
        wall_ranks = [float(x[0]) for x in g]
    steps_per_s = args.steps / wall_max
    norm = N * world / NORM_FLOES
    value = steps_per_s * norm

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_fa = pin_floe_arrays(h.download_floes(mc=False))
        h2d = dyn_bytes(host_fa)
        def e2e_step(t, fused):
            if sl is not None:
                sl.step_host([host_fa], t, True)         # uploads, publication to the neighbours, step, overlapped D2H
            elif fused:
                h.step_host(host_fa, t, True)            # one C-ABI call: H2D + kernels + D2H, overlapped
            else:
                h.upload_state(host_fa)                  # H2D: every per-floe scalar + ring coordinates
                h.step(t, True)
                h.download_floes(into=host_fa, mc=False)  # D2H: the same state back
            return float(host_fa.collision_force[0, 0])

        def time_e2e(fused):
            for t in range(3):
                e2e_step(t, fused)
            barrier()
            t0 = time.perf_counter()
            ne = max(3, min(args.steps, 20))
            for t in range(ne):
                e2e_step(t, fused)
            torch.cuda.synchronize()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return ne / float(te[0])

        d2h = h2d + host_fa.id.nbytes + host_fa.ghost_id.nbytes
        sep = time_e2e(False) if sl is None else None
        # sz_step_host does not upload what the step overwrites before any read: collision_force / collision_trq
        # (zeroed by timestep_collisions!) and, on a step that runs the coupling, fxOA / fyOA / trqOA / hflx_factor
        h2d_fused = (h2d - host_fa.collision_force.nbytes - host_fa.collision_trq.nbytes - host_fa.fxOA.nbytes
                     - host_fa.fyOA.nbytes - host_fa.trqOA.nbytes - host_fa.hflx_factor.nbytes)
        e2e_sps = time_e2e(True)
        # the same loop for a shim that knows what its host processes touched (sz_step_host_partial): here an observer that
        # tags floes and reads positions, velocities and forces every step — reported BESIDE the all-fields number
        up_f, dn_f = ("status_tag",), ("centroid_x", "centroid_y", "alpha", "u", "v", "xi", "collision_force", "collision_trq",
                                      "fxOA", "fyOA", "trqOA", "status_tag")

        def masked_step(t):
            if sl is not None:
                sl.step_host_partial([host_fa], t, True, upload=up_f, download=dn_f)
            else:
                h.step_host_partial(host_fa, t, True, upload=up_f, download=dn_f)

        for t in range(3):
            masked_step(t)
        barrier()
        t0 = time.perf_counter()
        ne = max(3, min(args.steps, 20))
        for t in range(ne):
            masked_step(t)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        masked = {"value": ne / float(te[0]) * norm, "unit": "steps/s", "steps_per_s": ne / float(te[0]),
                  "h2d_bytes_per_step": int(sum(getattr(host_fa, k).nbytes for k in up_f)),
                  "d2h_bytes_per_step": int(sum(getattr(host_fa, k).nbytes for k in dn_f)),
                  "upload": list(up_f), "download": list(dn_f)}
        e2e = {"value": e2e_sps * norm, "unit": "steps/s", "steps_per_s": e2e_sps,
               "h2d_bytes_per_step": int(h2d_fused), "d2h_bytes_per_step": int(d2h),
               "call": ("sz_step_host on pinned host arrays (upload of every per-floe input scalar + rings, step, download "
                        "of the whole state; copies overlap the kernels)" if world == 1 else
                        "sz_slab_step_host on pinned host arrays of every rank's local list (uploads, publication of the "
                        "uploaded boundary floes to the neighbours, step, overlapped downloads)"),
               "separate_calls_steps_per_s": sep, "masked": masked,
               "floor": ("PCIe: the 33 MB that only exist after the state update leave in 0.8 ms (41 GB/s measured) behind 1.08 ms of "
                         "kernels that cannot start before the first upload group has landed (0.16 ms): 2.04 ms = 490 steps/s is the "
                         "floor of an all-fields synchronous step on this box (profiles/README.md)" if world == 1 else
                         "host bandwidth: with 8 ranks copying at once this box gives each rank 17-23 GB/s (88 GB/s alone; "
                         "tools/pcie_concurrency.py, profiles/r2/r3a_pcie_n8.json): the 79 MB of an all-fields step take 3.7-5.0 ms "
                         "whatever the library does; `masked` (10 MB per step) is what a shim that names its fields gets")}

    # ---- parity of the benchmark field against the oracle (one step, after the timed region) --------------------
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_vs_oracle(args, prod, local_rank, world)

    halo = None
    if sl is not None:
        # transport check: every halo copy must carry its owner's state bit for bit.  Each rank publishes a checksum per
        # owned boundary floe, the holders of copies compare (outside the timed region, gloo).
        st = sl.stats()
        g, o = sl.local_index(0)
        sl.refresh_halo()  # collective: halo copies := their owners' current state
        fa = h.download_floes(mc=False)
        key = np.stack([fa.centroid_x, fa.centroid_y, fa.u, fa.v, fa.xi, fa.alpha, fa.height], axis=1).view(np.uint64)
        chk = np.bitwise_xor.reduce(key * np.arange(1, 8, dtype=np.uint64)[None, :], axis=1)
        mine = {int(a): int(b) for a, b in zip(g[o == rank], chk[o == rank])} if False else None
        tables = [None] * world
        own_mask = o == rank
        dist.all_gather_object(tables, (g[own_mask], chk[own_mask]), group=gloo)
        bad = 0
        for src in range(world):
            if src == rank:
                continue
            sel = np.nonzero(o == src)[0]
            if len(sel) == 0:
                continue
            gs, cs = tables[src]
            pos = np.searchsorted(gs, g[sel])
            ok = (pos < len(gs)) & (gs[np.minimum(pos, len(gs) - 1)] == g[sel])
            bad += int((~ok).sum()) + int((cs[pos[ok]] != chk[sel[ok]]).sum())
        hs = torch.tensor([st["halo_floes"], st["send_bytes_per_step"], sl.max_displacement(), bad, st["rebuilds"],
                           sum(rebuild_s)], dtype=torch.float64, device="cuda")
        hmax = hs.clone()
        dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(hs, op=dist.ReduceOp.SUM)
        assert int(hs[3]) == 0, "halo copies differ from their owners' state on %d floes" % int(hs[3])
        halo = {"halo_floes_max": int(hmax[0]), "send_bytes_per_step_max": int(hmax[1]), "skin_m": args.skin,
                "max_displacement_m": float(hmax[2]), "lists_stale": bool(hmax[2] > 0.5 * args.skin),
                "rebuilds": int(hmax[4]), "rebuild_seconds_total_max": float(hmax[5]), "displacement_polls": polls[-12:],
                "halo_copies_equal_owner_state": True, "rank0_cpu_binding": numa,
                "exchange": "sz_slab_step: k_slab_unpack (wait for the neighbours' flag, scatter) -> step -> k_slab_push (8 doubles + "
                            "ring per boundary floe straight into the neighbours' arenas over NVLink, cudaIpc-mapped) — no NCCL "
                            "call, no pack / unpack round trip; NCCL only for this script's barrier / all-reduce of the timings"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    g = f.grid
    per = phase / args.steps
    P, C = c["n_pairs"], c["n_rows"]
    nbar = V / max(N, 1)
    Pc = c["n_candidates"]
    kernels = {
        "k_coupling": (per[4], 16.0 * M + 120.0 * N + 40.0 * (g.Nx + 1) * (g.Ny + 1)),
        "k_narrow": (per[2], P * (16.0 * (2 * nbar) + 128.0) + 112.0 * C),
        "k_update": (per[5], 440.0 * N + 32.0 * V + 56.0 * C),
        "broad": (per[1], 56.0 * N + 8.0 * Pc),   # SURVEY §8(d) K1 + K2
        "rows": (per[3], 112.0 * C + 24.0 * N),
    }
    dom = max(kernels, key=lambda k: kernels[k][0])
    kms, kbytes = kernels[dom]
    achieved = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    try:  # ncu dram bytes per launch of the same workload (committed with the profile it comes from)
        if N == 100000 and args.npoints == 1000:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    notes = {"k_narrow": "narrow phase = k_item_count/scatter + k_narrow_ab<0> (clip, decisions) + k_narrow_ab<1> (forces) + k_narrow "
                         "(warp kernel: large rings and every rare path): FP64 latency bound, not bandwidth bound (ncu r3g: 2-4 % of DRAM "
                         "peak, 29-36 % issue slots, 12 warps per SM, 20-24 of 32 lanes active, FP64 pipe 14-16 %); the HBM fraction is reported because the "
                         "contract asks for it, roofline.fp64 gives the FP64 view (profiles/README.md)",
             "k_coupling": "streams 16 B per Monte-Carlo point once (ncu: 962 MB = the algorithmic bytes) through cp.async.bulk; FP64 pipe "
                           "55 % busy, issue slots 65 %; inside sz_step it runs on a second stream beside the broad phase",
             "k_update": "", "broad": "chain of small dependent kernels (uniform grid, neighbour lists, image filter): latency bound",
             "rows": "per-floe row assembly + sequential sums in the reference's row order (deterministic, no float atomics)"}
    roofline = {"kernel": dom if dom != "k_narrow" else "k_narrow_ab", "bound": "hbm", "achieved": achieved, "peak": hbm,
                "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "note": notes[dom],
                "all_kernels": {k: {"ms": v[0], "algorithmic_bytes": v[1], "GB/s": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0),
                                    "frac": (v[1] / (v[0] * 1e-3) / 1e9 / hbm if v[0] > 0 else 0.0)} for k, v in kernels.items()},
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": kbytes, "kernel_ms": kms,
                "phase_ms": {"ghosts": per[0], "broad": per[1], "narrow": per[2], "rows": per[3], "coupling": per[4],
                             "update": per[5], "step_device": per[6]}}

    # FP64 view of the same kernels: measured DFMA peak of this GPU (tools/fp64_peak, run after the
    # timed region) against the algorithmic FP64 operations: ~80 per in-bounds Monte-Carlo point (coupling, counted
    # from cp_point with the atmosphere field zero), ~30 per (edge, edge) test and one clip per candidate pair plus
    # one per overlapping pair (narrow phase)
    fp64 = None
    try:
        exe = os.path.join(ROOT, "tools", "fp64_peak")
        pk = json.loads(subprocess.run([exe], capture_output=True, text=True, timeout=60, check=True).stdout.strip())
        e2 = (nbar - 1.0) ** 2
        fl = {"k_coupling": 80.0 * M, "k_narrow": 30.0 * e2 * (c["n_pairs"] + c["n_overlap"])}
        fp64 = {"peak_tflops": pk["dfma_tflops"], "peak_unfused_tflops": pk["dmul_dadd_tflops"], "peak_source": "tools/fp64_peak.cu on this GPU: " + pk["how"],
                "kernels": {k: {"algorithmic_flops": v, "TFLOP/s": v / (kernels[k][0] * 1e-3) / 1e12,
                                "frac": v / (kernels[k][0] * 1e-3) / 1e12 / pk["dfma_tflops"]} for k, v in fl.items() if kernels[k][0] > 0}}
    except Exception as e:  # the microbenchmark is evidence, not part of the product
        fp64 = {"error": repr(e)}
    roofline["fp64"] = fp64

    cpu = None
    if not args.no_cpu_baseline:
        ns = min(args.cpu_sample, args.floes)
        sps, cores, _ = oracle_steps_per_s(args, ns, 2, 1)
        cpu = {"value": sps * ns / args.floes, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": "oracle port (OpenMP, %d threads), 2 timed steps on a %d-floe field of the same generator, "
                         "steps/s scaled by %d/%d" % (cores, ns, ns, args.floes)}

    line = {
        "metric": METRIC, "value": value * 1.0, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "steps_per_s": steps_per_s, "floes_total": N * world, "floe_steps_per_s": steps_per_s * N * world,
        "contacts_per_s": steps_per_s * c["n_overlap"] * world, "candidate_pairs_per_s": steps_per_s * c["n_candidates"] * world,
        "mc_points_per_s": steps_per_s * M * world, "parity": parity,
        "device_ms_per_step": 1e3 * dev_max / args.steps,
        "step_wall_ms_rank0": {"p50": 1e3 * float(np.percentile(step_wall, 50)), "p90": 1e3 * float(np.percentile(step_wall, 90)),
                               "p99": 1e3 * float(np.percentile(step_wall, 99)), "max": 1e3 * float(step_wall.max()),
                               "first10_mean": 1e3 * float(step_wall[:10].mean()), "last10_mean": 1e3 * float(step_wall[-10:].mean())},
        "counts": {k: c[k] for k in ("n_init", "n_candidates", "n_pairs", "n_overlap", "n_rows", "n_mc", "n_vertices")},
        "halo": halo, "wall_ms_per_step_ranks": [round(w / args.steps * 1e3, 4) for w in wall_ranks], "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
