#!/usr/bin/env python
"""bench.py -- throughput of the FDTD time-stepping hot path (updateE + updateH) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4]): the synthetic all-PML (CPML Num=12) vacuum
box with one point dipole, z-slab sharded, WEAK scaling: 2048 x 2048 x 256 cells per GPU, so
N = 8 is the named 2048^3 cube.  A step = one full Yee step (updateE + updateH) of the whole
grid.  Metric: Mcells/s (cells incl. PML x steps / second), whole job.

  value     device-resident stepping through the engine C ABI (alisb_step_E/H), CUDA events
            on the engine's stream, barrier + synchronize on both sides, max over ranks
  e2e       the same steps driven the way a user's time loop drives them: per step the source
            sample is produced on the host and a probe value (get(W,Ez,...)) is read back
            device->host, host wall clock around the loop
  roofline  the dominant kernel (the interior launch): algorithmic bytes (36 B per cell per
            half-step pass = 72 B per cell-step) / CUDA-event duration of that launch
            -> `frac`; beside it `frac_of_moved_bytes` on the bytes the kernel must move (a fused
            E+H launch: 48 B per cell-step) and `traffic` (ncu dram bytes, reported only while the
            kernel sources are the captured ones)
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, bounded sample

`--impl reference` times the reference CPU implementation (oracle/_ref/bin/bench_cube_ref, the
reference library built from its own sources; falls back to the oracle port) on ONE GPU's shard
of the same workload (2048 x 2048 x 256 with all six CPML sides, when the host RAM allows).
`--scaling strong` runs the named 2048^3 cube split over N GPUs (N = 1: 2048 x 2048 x 1024).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, NZ_PER_GPU, NUM = 2048, 2048, 256, 12
FUSE_DEFAULT = "1"      # fused E+H stepping (alisb_set_fusion); --fuse 0 times the two-pass path
METRIC, UNIT = "Mcells/s", "Mcells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--ny", type=int, default=NY)
    ap.add_argument("--nz-per-gpu", type=int, default=NZ_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, what the driver runs): --nz-per-gpu planes per GPU; strong: the named 2048^3 cube split over the GPUs "
                         "(--nz-total planes in all; the 1-GPU reference point is 2048x2048x1024, SURVEY 8d: the cube does not fit one GPU)")
    ap.add_argument("--nz-total", type=int, default=2048)
    ap.add_argument("--fuse", type=int, default=int(os.environ.get("ALISB_FUSE", FUSE_DEFAULT)),
                    help="1: fused E+H stepping where a step is only the two curls (48 B per cell-step instead of 72)")
    return ap.parse_args()


def workload_name(nx, ny, nz, n):
    return f"synthetic vacuum+CPML(Num={NUM}) box {nx}x{ny}x{nz} cells, point dipole Ez/Pulse, z-slab sharded over {n} GPU(s), {nz // n} planes per GPU"


def total_planes(args, world):
    """planes of the whole job: weak = nz_per_gpu per GPU; strong = the fixed cube (1 GPU: half of it, it does not fit)"""
    if args.scaling == "weak":
        return args.nz_per_gpu * world
    return args.nz_total if world > 1 else min(args.nz_total, 1024)


def config_of(args, world):
    """identical in both arms (the driver compares them)"""
    nz = total_planes(args, world)
    return {"workload": workload_name(args.nx, args.ny, nz, world),
            "l2": "inputs (6 field arrays, >= 26 GB per GPU) are far larger than the 126 MB L2"}


def pml_share(nx, ny, nz):
    """fraction of the box's cells that lie in a CPML slab (Num + 1 cells on each of the six sides)"""
    d = NUM + 1
    inner = max(nx - 2*d, 0) * max(ny - 2*d, 0) * max(nz - 2*d, 0)
    return 1.0 - inner / float(nx * ny * nz)


def host_sample_planes(nx, ny, want):
    """z extent of the CPU sample: the GPU arm's per-GPU box when the host has the memory for the reference's
    arrays (6 fields + CPML slabs, ~25 B per cell), else the largest power of two that fits"""
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        avail = 32 << 30
    nz = want
    while nz > 32 and nx * ny * nz * 26 * 1.15 > 0.8 * avail:
        nz //= 2
    return nz


# ------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref)
# ------------------------------------------------------------------------------------------
def run_reference_cube(nx, ny, nz, steps, warmup):
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "bench_cube_ref")
    cores = os.cpu_count() or 1
    if os.path.exists(exe):
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        out = subprocess.run([exe, str(nx), str(ny), str(nz), str(steps), str(warmup)], env=env, capture_output=True, text=True, timeout=3000).stdout
        m = re.search(r"seconds=([0-9.]+) mcells_per_s=([0-9.]+) threads=(\d+)", out)
        if m:
            return {"value": float(m.group(2)), "seconds": float(m.group(1)), "cores": int(m.group(3)), "kind": "reference"}
    # fall back to the oracle port (plain-C restatement, OpenMP)
    from alis_b200 import capi, synthetic
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    w = synthetic.BoxWorld(nx, ny, nz, num=NUM)
    e = w.make_engine(capi.load_oracle())
    e.run(warmup)
    t0 = time.time()
    e.run(steps)
    dt = time.time() - t0
    return {"value": w.cells() * steps / dt * 1e-6, "seconds": dt, "cores": cores, "kind": "port"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_gpu = total_planes(args, args.gpus) // args.gpus
    sample_nz = host_sample_planes(args.nx, args.ny, per_gpu)      # one GPU's shard of the workload when the host RAM allows
    r = run_reference_cube(args.nx, args.ny, sample_nz, args.steps, args.warmup)
    sample = (f"{args.nx}x{args.ny}x{sample_nz} box (one GPU's shard of the workload with all six CPML sides, "
              f"{100 * pml_share(args.nx, args.ny, sample_nz):.1f}% of its cells in CPML slabs), {args.steps} steps, all host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, args.steps),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, args.gpus), "timed_sample": sample,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for q, n in enumerate(names) if any(len(s) > 2 + q and s[2 + q].lower() == "active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from alis_b200 import capi, synthetic, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    torch.cuda.set_stream(torch.cuda.Stream())       # a real (non-default) stream shared by the engine, NCCL ordering and the timing events
    lib = capi.load_engine()
    nz = total_planes(args, world)
    w = synthetic.BoxWorld(args.nx, args.ny, nz, num=NUM)
    os.environ["ALISB_FUSE"] = str(args.fuse)        # read by the engine at creation (here and in the C driver of the e2e leg)
    eng = w.make_engine(lib, device=local, rank=rank, world=world)
    runner = sharding.SlabRunner(eng, rank, world, cuda=True, overlap=not os.environ.get("ALISB_NO_OVERLAP"), fuse=bool(args.fuse))      # engine now runs on torch's current stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    runner.run(args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- device-resident timing --------------------------------------------------------
    eng.set_profiling(True)
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        runner.step()
    runner.flush()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    fused_launches = eng.fused_launch_count()
    kinfo = lib_kernel_stats(eng)
    eng.set_profiling(False)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = w.cells() * args.steps / (ms * 1e-3) * 1e-6

    # ---- end to end: the drop-in C API (createWorld + pointDipole + updateE/updateH + get per step) -----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_through_c_api(args, world, rank, local, w, eng, runner, barrier)

    clocks = None
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=5)
        clocks = sampler.summary()

    if rank == 0:
        peaks, peak_src = load_peaks()
        roof = None
        if kinfo:
            ach = kinfo["bytes_per_launch"] / (kinfo["avg_ms"] * 1e-3) / 1e9
            # what the kernel must move: a fused E+H launch reads and writes every field once (48 B per cell-step,
            # 2/3 of the two-pass 72 B the algorithmic figure is quoted on)
            moved = kinfo["bytes_per_launch"] * (2.0 / 3.0 if fused_launches else 1.0)
            traffic, traffic_note = load_traffic("yee_fused" if fused_launches else "yee_volume")
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks, "unit": "GB/s", "frac": ach / peaks, "traffic": traffic, "traffic_source": traffic_note,
                    "moved_bytes_per_launch": moved, "achieved_moved_GBps": moved / (kinfo["avg_ms"] * 1e-3) / 1e9,
                    "frac_of_moved_bytes": moved / (kinfo["avg_ms"] * 1e-3) / 1e9 / peaks,
                    "kernel": ("yee_fused (interior, E and H half steps in one pass; algorithmic bytes are the two-pass 72 B per cell-step of SURVEY 8d, "
                               "the kernel itself moves 48 B per cell-step, so frac may exceed 1)") if fused_launches else "yee_volume (interior, no-PML instantiation)",
                    "avg_launch_ms": kinfo["avg_ms"], "launches_timed": kinfo["n"],
                    "bytes_per_launch": kinfo["bytes_per_launch"], "peak_source": peak_src,
                    "whole_step": {"algorithmic_bytes": w.algorithmic_bytes_per_step() / world, "achieved_GBps": w.algorithmic_bytes_per_step() / world / (ms / args.steps * 1e-3) / 1e9}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            snz = host_sample_planes(args.nx, args.ny, nz)
            r = run_reference_cube(args.nx, args.ny, snz, 8, 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                   "sample": f"{args.nx}x{args.ny}x{snz} box of the same workload ({100 * pml_share(args.nx, args.ny, snz):.1f}% of its cells in CPML slabs), 8 steps after 1 warm-up, all host threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_of(args, world), "per_gpu_mcells_per_s": value / world,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "fused_launches": fused_launches, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_through_c_api(args, world, rank, local, w, eng, runner, barrier):
    """The metric measured the way a user of the library drives it.  N = 1: a C driver written against
    the public ALiS API (oracle/harness/bench_cube.c linked with alis_b200/libalis.so: createWorld,
    pointDipole, then updateE; updateH; get(W,Ez,...) every step), host wall clock around the loop.
    N > 1: the same loop through the Python binding of the C ABI (the C API is single-process), with
    the probe read back every step."""
    import torch
    import torch.distributed as dist
    from alis_b200 import capi
    exe = os.path.join(ROOT, "tests", "bin", "bench_cube_mine_cuda")
    note = "FDTD state is device-resident by design: per step the host supplies the source samples (kernel arguments, 8 B) and reads one probe value back (4 B)"
    if world == 1 and os.path.exists(exe):
        eng_free = getattr(eng, "close", None)
        del runner
        eng.close()                                  # give the 27 GB back before the C driver allocates its own world
        torch.cuda.empty_cache()
        env = dict(os.environ, ALIS_DEVICE=str(local))
        out = subprocess.run([exe, str(args.nx), str(args.ny), str(total_planes(args, world) // world), str(args.steps), str(args.warmup), "1"],
                             env=env, capture_output=True, text=True, timeout=3000).stdout
        m = re.search(r"seconds=([0-9.]+) mcells_per_s=([0-9.]+)", out)
        if m:
            res = {"value": float(m.group(2)), "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 4,
                   "path": "C driver on the public API: tests/bin/bench_cube_mine_cuda (oracle/harness/bench_cube.c + alis_b200/libalis.so)", "note": note}
            # the same loop with a plane monitor at the cadence of examples/Dipole_PML.c: sliceMax(W, EE, XY) every 4th step
            out2 = subprocess.run([exe, str(args.nx), str(args.ny), str(total_planes(args, world) // world), str(args.steps), str(args.warmup), "2"],
                                  env=env, capture_output=True, text=True, timeout=3000).stdout
            m2 = re.search(r"seconds=([0-9.]+) mcells_per_s=([0-9.]+)", out2)
            if m2:
                res["with_plane_monitor"] = {"value": float(m2.group(2)), "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 4 + 1,
                                             "what": "additionally sliceMax(W, EE, XY plane z=0) every 4th step, reduced on the device (kernel reduce_absmax): 4 more bytes come back every 4th step"}
            return res
    probe_c = min(max(eng.shard[0], w.ax[2].arr(5)), eng.shard[1]) - (eng.shard[0] - 1)
    pbox = [w.ax[0].arr(0), w.ax[0].arr(0), w.ax[1].arr(0), w.ax[1].arr(0), probe_c, probe_c]
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(args.steps):
        runner.step()
        runner.flush()
        acc += float(eng.read_box(capi.ARR_FIELD, [0, 2, 0], pbox)[0, 0, 0])   # get(W, Ez, ...): 4 B D2H, syncs
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    return {"value": w.cells() * args.steps / dt * 1e-6, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 4,
            "path": "Python binding of the engine C ABI (alis_b200/capi.py), one process per GPU", "note": note}


def lib_kernel_stats(eng):
    """per-launch CUDA-event times of the dominant kernel, collected by the engine"""
    avg, n, by = eng.profile_dominant()
    if n == 0:
        return None
    return {"avg_ms": avg, "n": n, "bytes_per_launch": by}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def kernels_sha():
    import hashlib
    h = hashlib.sha256()
    for f in ("kernels.cuh", "engine.cu"):
        h.update(open(os.path.join(ROOT, "alis_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def load_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
    `ncu --set full` capture of this bench command, profiles/).  ncu cannot run inside a timed bench, so the
    figure is a recorded one: it is reported only while the kernel sources are the ones it was captured from
    (sha of alis_b200/csrc), otherwise null -- never a stale constant."""
    p = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(p):
        try:
            rec = json.load(open(p))
            k = rec[kernel]
            if rec.get("csrc_sha16") != kernels_sha():
                return None, f"capture {k.get('source', p)} predates the current kernel sources: not reported"
            return k.get("dram_bytes_per_launch"), f"{k.get('source', p)} (same kernel sources, sha {rec.get('csrc_sha16')})"
        except Exception:
            return None, "no capture"
    return None, "no capture"


def emit(line):
    """the ONE JSON line, on the process's real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # libraries (NCCL's version banner, torchrun) write to fd 1; keep stdout for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
