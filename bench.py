#!/usr/bin/env python
"""bench.py -- ELF-strip throughput of the B200 path vs the reference's host `strip` pipeline.

  python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo (CUDA kernels)
  python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference's CPU pipeline

Metric (BASELINE.json): ELF-strip GB/s of build-tree `.so` INPUT bytes.  One "step" = one pass of
the hot path over one batch: every file of this rank's shard of the synthetic corpus (config 4 of
BASELINE.json: sizes log-uniform 1 KB..128 MB, seed 0xB200, debug fraction U(0.05,0.8); 1250 files
per GPU, i.e. the 10 000-file / ~100 GB corpus at 8 GPUs -- weak scaling, files are dealt size-sorted
round-robin, no payload crosses GPUs; one NCCL allgather of per-rank byte counts per step).

  value     : device-resident.  Inputs already in HBM; timed = upload of offsets + plan kernel +
              offset scan + compaction kernel + fetch of sizes/status (+ allgather when N > 1).
  e2e       : the same batch through the C ABI with HOST buffers (lb2_strip_host): pinned H2D of
              every input byte, kernels, D2H of every output byte, inside the timed region.
  roofline  : compaction kernel, algorithmic bytes (copied extents read + output written) over its
              CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the reference's own line `find DIR/ -name "*.so" | xargs strip`
              (/root/reference/lambdipy/project_build.py:260) on /dev/shm over a size-balanced 1/8
              sample of the same corpus: serial as the reference runs it, and `xargs -P nproc -n 1`.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ELF-strip GB/s (build-tree .so bytes)"
FILES_PER_GPU = 1250
SEED = 0xB200


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------- clocks during the timed region
class ClockSampler:
    """SM clocks and throttle reasons DURING the timed regions.  NVML from a thread (a few microseconds
    per sample); `nvidia-smi -lms` as fallback -- polling nvidia-smi at 100 ms measurably slowed the
    sampled GPU's kernels in the 8-GPU runs, NVML queries do not."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.idx)
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(mx))
                r = get_reasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv
            import threading
            nv.nvmlInit()
            nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.path = tempfile.mktemp(prefix="lb2_clocks_", suffix=".csv")
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            out["source"] = "nvml"
        elif self.proc:
            out["source"] = "nvidia-smi"
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.f.close()
            try:
                with open(self.path) as f:
                    for line in f:
                        p = [x.strip() for x in line.split(",")]
                        if len(p) < 9:
                            continue
                        try:
                            self.sm.append(float(p[1])); self.mx.append(float(p[2]))
                        except ValueError:
                            continue
                        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                            if v.lower().startswith("active"):
                                self.reasons.add(name)
                os.unlink(self.path)
            except Exception:
                pass
        if self.sm:
            out.update(sm_mhz=statistics.median(self.sm), sm_max_mhz=max(self.mx), reasons=sorted(self.reasons), samples=len(self.sm))
        return out


# ---------------------------------------------------------------- the reference pipeline on host cores
REF_LINE = 'find {d}/ -name "*.so" | xargs strip'              # project_build.py:260, verbatim
PAR_LINE = 'find {d}/ -name "*.so" | xargs -P {p} -n 1 strip'  # same tool, all host cores


def _materialize(args):
    corpus, i, path = args
    with open(path, "wb") as f:
        f.write(corpus.materialize(i))
    return os.path.getsize(path)


def cpu_strip_bench(write_master, n_bytes, steps, warmup, parallel_only=False):
    """write_master(dir) populates dir with *.so inputs.  Returns GB/s of the reference line, serial
    and with -P nproc.  Each timed run works on a fresh copy (strip rewrites files in place)."""
    nproc = os.cpu_count() or 1
    base = tempfile.mkdtemp(prefix="lb2_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        master = os.path.join(base, "master")
        os.makedirs(master)
        write_master(master)

        def timed(line, reps, warm):
            best, times = None, []
            for r in range(warm + reps):
                run = os.path.join(base, "run")
                shutil.copytree(master, run)
                cmd = line.format(d=run, p=nproc)
                t0 = time.perf_counter()
                rc = subprocess.run(["bash", "-c", "set -e; set -o pipefail; " + cmd], capture_output=True)
                dt = time.perf_counter() - t0
                shutil.rmtree(run)
                if rc.returncode != 0:
                    raise RuntimeError("reference strip pipeline failed: %s" % rc.stderr.decode()[:300])
                if r >= warm:
                    times.append(dt)
            return times

        par = timed(PAR_LINE, steps, warmup)
        ser = [] if parallel_only else timed(REF_LINE, max(1, min(steps, 2)), 0)
        res = {"bytes": n_bytes, "nproc": nproc,
               "parallel_s": par, "parallel_gbs": n_bytes / 1e9 / (sum(par) / len(par)),
               "serial_s": ser, "serial_gbs": (n_bytes / 1e9 / (sum(ser) / len(ser))) if ser else None}
        return res
    finally:
        shutil.rmtree(base, ignore_errors=True)


def strip_version():
    try:
        return subprocess.run(["strip", "--version"], capture_output=True, text=True).stdout.splitlines()[0]
    except Exception:
        return "unknown"


# ---------------------------------------------------------------- arms
def run_reference(a, rank, world):
    if rank != 0:
        return 0
    from lambdipy_b200.corpus import Corpus
    from multiprocessing import Pool
    # the same 1/8 size-balanced sample of the N=1 workload that the b200 arm's cpu_baseline uses
    sample = Corpus(a.files_per_gpu, seed=SEED, rank=0, world=8)
    n_bytes = sample.total_bytes

    def write_master(d):
        jobs = [(sample, i, os.path.join(d, "f%05d.so" % i)) for i in range(len(sample))]
        with Pool(min(32, os.cpu_count() or 1)) as pool:
            pool.map(_materialize, jobs, chunksize=4)

    res = cpu_strip_bench(write_master, n_bytes, a.steps, max(a.warmup, 1))
    value = res["parallel_gbs"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sum(res["parallel_s"]) / len(res["parallel_s"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(a, world),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": res["nproc"], "kind": "reference",
                         "sample": "size-balanced 1/8 of the N=1 workload: %d files, %.3f GB on /dev/shm; `%s` (GNU strip: %s); "
                                   "serial as the reference runs it (1 process): %.3f GB/s" %
                                   (len(sample), n_bytes / 1e9, PAR_LINE.format(d="DIR", p=res["nproc"]), strip_version(), res["serial_gbs"] or 0),
                         "serial_value": res["serial_gbs"]},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(a, world):
    return {"workload": "synthetic ELF corpus (BASELINE config 4/5): %d files per GPU (%d total), sizes log-uniform 1 KB-128 MB, "
                        "seed 0x%X, dropped fraction U(0.05,0.8); files dealt size-sorted round-robin over ranks" %
                        (a.files_per_gpu, a.files_per_gpu * world, SEED),
            "files_per_gpu": a.files_per_gpu, "parallelism": "file-sharded x%d, no payload exchange" % world,
            "l2": "inputs (>10 GB per GPU) far larger than the 126 MB L2; no flush needed"}


def run_b200(a, rank, local_rank, world):
    os.environ["NCCL_DEBUG"] = os.environ.get("LB2_NCCL_DEBUG", "NONE")  # NCCL prints its version banner on stdout at any level >= VERSION; bench prints ONE JSON line
    import numpy as np
    import torch
    from lambdipy_b200 import _native as N
    from lambdipy_b200.corpus import Corpus
    from lambdipy_b200.device import DeviceBatch

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = N.Context(local_rank)
    corpus = Corpus(a.files_per_gpu * world, seed=SEED, rank=rank, world=world)
    batch = DeviceBatch.from_corpus(ctx, corpus)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    gathered = torch.zeros(4 * world, dtype=torch.int64, device="cuda")
    # The allgather of the per-rank counters runs on a side stream: the next batch does not have to
    # wait for the slowest rank's counters (files never move between GPUs); the timed region ends
    # only after every step's allgather has completed.
    side = torch.cuda.Stream() if world > 1 else None
    ring = [(torch.zeros(4, dtype=torch.int64, device="cuda"), torch.zeros(4 * world, dtype=torch.int64, device="cuda")) for _ in range(8)] if world > 1 else []
    ring_pos = [0]

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        batch.strip_async(stream=sptr)
        st = batch.results()
        if dist:  # the one collective of the path: per-rank byte counts (32 bytes per rank)
            c, g = ring[ring_pos[0] % len(ring)]
            ring_pos[0] += 1
            with torch.cuda.stream(side):
                c.copy_(torch.tensor([st["in_bytes"], st["out_bytes"], st["n_ok"], st["n_unsupported"]], dtype=torch.int64), non_blocking=True)
                dist.all_gather_into_tensor(g, c)
        return st

    for _ in range(max(a.warmup, 3)):
        st = step()
    assert st["n_unsupported"] == 0 and st["n_ok"] == len(corpus), st
    # ---- device-resident timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    compact_ms, plan_ms = [], []
    e0.record(stream)
    for _ in range(a.steps):
        st = step()
        compact_ms.append(st["compact_ms"]); plan_ms.append(st["plan_ms"])
    if side is not None:
        stream.wait_stream(side)  # every step's allgather is inside the timed region
    e1.record(stream)
    barrier()
    if dist:
        g = ring[(ring_pos[0] - 1) % len(ring)][1].cpu().numpy().reshape(world, 4)
        assert int(g[:, 2].sum()) == a.files_per_gpu * world and int(g[:, 3].sum()) == 0, g
    dev_ms = e0.elapsed_time(e1)
    if a.profile_mode:
        from lambdipy_b200.sharding import gather_counts
        per_rank = gather_counts([int(1e3 * sum(compact_ms) / len(compact_ms)), int(1e3 * sum(plan_ms) / len(plan_ms)),
                                  st["copy_bytes"] + st["out_bytes"], int(1e3 * dev_ms / a.steps)], device="cuda")
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": dev_ms / a.steps, "compact_ms": compact_ms, "plan_ms": plan_ms,
                              "per_rank_[compact_us, plan_us, alg_bytes, step_us]": per_rank.tolist(),
                              "per_rank_compact_frac": [float(r[2]) / 1e9 / (r[0] / 1e6) / peaks()[0] for r in per_rank.tolist()],
                              "note": "not a bench value"}))
        batch.close()
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- end to end through host buffers
    in_span = int(corpus.off[-1])
    h_in = ctx.pinned_alloc(in_span + 256)
    out_cap = in_span + len(corpus) * 4096 + (16 << 20)
    h_out = ctx.pinned_alloc(out_cap)
    batch.read_input_arena(h_in)
    n = len(corpus)
    out_off = np.zeros(n + 1, dtype=np.uint64); out_sizes = np.zeros(n, dtype=np.uint64); status = np.zeros(n, dtype=np.int32)
    u64p = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint64))
    hst = N.Stats()

    def e2e_step():
        ctx.check(ctx.lib.lb2_strip_host(ctx.h, h_in, u64p(batch.off), u64p(batch.sizes), n, h_out, out_cap, u64p(out_off), u64p(out_sizes),
                                         status.ctypes.data_as(C.POINTER(C.c_int32)), 0, C.byref(hst)))
        if dist:
            counts.copy_(torch.tensor([hst.in_bytes, hst.out_bytes, hst.n_ok, hst.n_unsupported], dtype=torch.int64), non_blocking=True)
            dist.all_gather_into_tensor(gathered, counts)
            torch.cuda.synchronize()

    e2e_steps = max(1, min(a.steps, a.e2e_steps))

    def time_e2e():
        e2e_step()  # warm-up (allocates workspaces / pipeline slots)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        return (time.perf_counter() - t0) * 1e3

    # default path of lb2_strip_host: zero-copy over the mapped pinned arenas; then, for comparison, the
    # staged pipeline (explicit H2D of whole files -> kernels in HBM -> D2H, 256 MB chunks on 3 streams)
    e2e_ms = time_e2e()
    zc_stats = (hst.in_bytes, hst.out_bytes, hst.copy_bytes)
    os.environ["LB2_HOST_ZEROCOPY"] = "0"
    staged_ms = time_e2e()
    os.environ.pop("LB2_HOST_ZEROCOPY")
    clocks = sampler.stop() if rank == 0 else None  # sampled across the device-resident and the e2e timed regions
    assert hst.n_ok == n and int(status.max()) == 0
    # e2e result check: host output of one mid-sized file equals the device-resident result
    probe = int(np.argsort(batch.sizes)[n // 2])
    assert C.string_at(h_out + int(out_off[probe]), int(out_sizes[probe])) == batch.read_output(probe)

    # ---- reduce over ranks: totals, max time
    local = torch.tensor([st["in_bytes"], st["out_bytes"], st["copy_bytes"], st["header_bytes"], st["n_ok"], st["n_unsupported"],
                          float(in_span)], dtype=torch.float64, device="cuda")
    times = torch.tensor([dev_ms, e2e_ms, sum(compact_ms) / len(compact_ms), sum(plan_ms) / len(plan_ms), staged_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    tot_in, tot_out, tot_copy, tot_hdr, n_ok, n_uns, tot_span = [float(x) for x in local.tolist()]
    dev_ms, e2e_ms, cms, pms, staged_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        ms_per_step = dev_ms / a.steps
        value = tot_in / 1e9 / (ms_per_step / 1e3)
        e2e_value = tot_in / 1e9 / (e2e_ms / e2e_steps / 1e3)
        peak, peak_src = peaks()
        # dominant kernel: compaction.  Algorithmic bytes per launch on THIS rank = C + OUT.
        alg = (st["copy_bytes"] + st["out_bytes"])
        achieved = alg / 1e9 / (sum(compact_ms) / len(compact_ms) / 1e3)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get("compact_dram_bytes_per_launch")
            except Exception:
                pass
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            sample = Corpus(a.files_per_gpu, seed=SEED, rank=0, world=8)
            full = {int(g): i for i, g in enumerate(corpus.global_index)}

            def write_master(d):
                for k, g in enumerate(sample.global_index):
                    with open(os.path.join(d, "f%05d.so" % k), "wb") as f:
                        f.write(batch.read_input(full[int(g)]))

            nb = sum(int(batch.sizes[full[int(g)]]) for g in sample.global_index)
            r = cpu_strip_bench(write_master, nb, 2, 1)
            cpu = {"value": r["parallel_gbs"], "unit": "GB/s", "cores": r["nproc"], "kind": "reference",
                   "sample": "size-balanced 1/8 of this workload (%d files, %.3f GB) on /dev/shm through the reference's line "
                             "`find DIR/ -name \"*.so\" | xargs strip` with -P %d -n 1 (%s); serial (1 process, as the reference runs it): %.3f GB/s"
                             % (len(sample), nb / 1e9, r["nproc"], strip_version(), r["serial_gbs"]),
                   "serial_value": r["serial_gbs"]}
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(a, world),
            "totals": {"files": int(n_ok), "unsupported_files": int(n_uns), "in_gb": tot_in / 1e9, "out_gb": tot_out / 1e9,
                       "copied_gb": tot_copy / 1e9, "header_gb": tot_hdr / 1e9},
            "roofline": {"bound": "hbm", "kernel": "lb2_compact_kernel" if os.environ.get("LB2_COMPACT_TMA") == "0" else "lb2_compact_tma_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": sum(compact_ms) / len(compact_ms),
                         "plan_kernel_ms": sum(plan_ms) / len(plan_ms),
                         "whole_pass_frac": (alg + st["header_bytes"]) / 1e9 / ((sum(compact_ms) + sum(plan_ms)) / len(compact_ms) / 1e3) / peak},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(tot_copy + tot_hdr),
                    "d2h_bytes_per_step": int(tot_out), "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "lb2_strip_host on pinned, device-mapped host arenas: kernels pull headers + kept extents over PCIe and "
                           "push stripped files back (dropped sections never cross the bus)",
                    "staged": {"value": tot_in / 1e9 / (staged_ms / e2e_steps / 1e3), "ms_per_step": staged_ms / e2e_steps,
                               "h2d_bytes_per_step": int(tot_span), "d2h_bytes_per_step": int(tot_out),
                               "api": "LB2_HOST_ZEROCOPY=0: cudaMemcpyAsync of whole files in 256 MB chunks on 3 streams"}},
            "gpu_launches": 4 * a.steps,  # plan (2 variants), scan, compaction per step
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    ctx.pinned_free(h_in); ctx.pinned_free(h_out)
    batch.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files-per-gpu", type=int, default=FILES_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="device-resident steps only (for runs under ncu; not a bench value)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return run_reference(a, rank, world)
    if world != a.gpus and world == 1 and a.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d needs torchrun (WORLD_SIZE=1 seen); running 1 GPU\n" % a.gpus)
    return run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
