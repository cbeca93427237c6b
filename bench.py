#!/usr/bin/env python
"""bench.py -- reads/s mapped+aligned on BASELINE.json config 2 (4.6 Mbp reference, 100k CLR reads, 1 GPU).

One "step" = one pass of the read-to-reference hot path (seeds -> probe -> prefix filter -> banded DP with
traceback, first success in list order; locator.cpp:70-92 semantics at R = 0.3) over one batch of synthetic reads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): every rank maps its own batch of reads against its own replica
of the index (weak scaling, no data-path collective); one NCCL all-reduce of the hit/score counters plus a gather
of the 56-byte records closes each step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one hardware queue per aligner band-class stream (must be set before CUDA is initialised in this process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
# The path's one exchange is four counters and a few MB of records per step, queued under the next step's aligner kernels.
# A collective kernel spins on its SMs until every rank has arrived: with NCCL's default channel count that took ~10 % of the
# SMs away from the aligner for most of a step (2 and 8 GPUs: K3 51.3 ms under the exchange against 45.2 ms alone).  Two
# channels carry this payload in well under a step.
os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
os.environ.setdefault("NCCL_MAX_CTAS", "2")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

MASK = 0xff3c3ffc  # seeds.txt line 1: 111**111*11*1111
R = 0.3
REF_LEN = 4_600_000
METRIC = "reads_per_sec_mapped_aligned"
UNIT = "reads/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled every 200 ms.  NVML is asked
    directly from a thread of this process; a looping nvidia-smi (the fallback when NVML cannot be loaded) takes driver locks
    long enough to stretch one aligner step in ten by ~5 ms."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.stop_flag, self.thread, self.how = index, [], None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))

            def loop():
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append([str(sm), str(mx), "0"] + ["Active" if mask & bit else "Not Active" for _, bit in names])
                    except Exception:
                        pass
                    time.sleep(0.2)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in list(self.rows):
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "how": self.how}


def make_workload(nreads: int, rank: int):
    import workload
    t0 = time.time()
    ref = workload.reference(2, REF_LEN)
    lens = workload.read_lengths(3 + 1000 * rank, nreads, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    txt, offs, lens, starts = workload.reads(3 + 1000 * rank, ref, lens)  # ins 9 / del 4 / sub 2 %
    log(f"[rank {rank}] workload: ref {len(ref)} bp, {nreads} reads, {len(txt)} bases in {time.time() - t0:.1f}s")
    return ref, txt, offs, lens


def source_sha() -> str:
    """digest of the aligner's kernel sources: ties a committed ncu capture of K3 to the code it was taken with"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "pacbioassembly_b200", "csrc")
    for f in ("pb_align.cu", "pb_align_nb.cuh", "pb_internal.cuh"):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class CpuReference:
    """The reference's own CPU implementation of the path (oracle/_ref, compiled from the unmodified sources) -- or the
    C port (oracle/pb_oracle.c) when _ref was not built.  The seed map is built once, as locator.cpp does; each call
    maps a bounded, evenly spaced sample of the same reads with all the host threads given."""

    def __init__(self, ref):
        import cpu_libs
        self.ref_txt = ref
        self.r = cpu_libs.ref()
        t0 = time.time()
        if self.r is not None:
            self.kind = "reference"
            self.h = self.r.locator_open(ref, MASK)
        else:
            self.kind = "port"
            self.o = cpu_libs.oracle()
            self.h = self.o.index_build(ref, MASK, 0)
        self.build_s = time.time() - t0

    def run(self, txt, offs, lens, sample_reads: int, nthreads: int):
        n = len(lens)
        step = max(1, n // max(sample_reads, 1))
        ids = np.arange(0, n, step)[:sample_reads]
        s_lens = lens[ids]
        s_offs = np.zeros(len(ids), dtype=np.int64)
        np.cumsum(s_lens[:-1], out=s_offs[1:])
        s_txt = np.concatenate([txt[offs[i]: offs[i] + lens[i]] for i in ids]) if len(ids) else np.zeros(0, np.uint8)
        t0 = time.time()
        if self.kind == "reference":
            recs = self.r.locator_run(self.h, s_txt, s_offs, s_lens, R=R, nthreads=nthreads)
        else:
            recs = self.o.locate(self.h, self.ref_txt, s_txt, s_offs, s_lens, MASK, R=R, nthreads=nthreads)
        dt = time.time() - t0
        kept = int((s_lens >= 500).sum())
        return kept / dt, recs, ids, dt

    def close(self):
        if self.kind == "reference":
            self.r.locator_close(self.h)
        else:
            self.o.index_free(self.h)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    ref, txt, offs, lens = make_workload(args.reads, 0)
    nthreads = min(os.cpu_count() or 1, args.cpu_threads)
    sample = args.cpu_sample or 48 * nthreads
    times = []
    cpu = CpuReference(ref)
    kind = cpu.kind
    log(f"[reference] seed map built in {cpu.build_s:.1f}s ({kind})")
    for it in range(args.warmup + args.steps):
        rps, recs, ids, dt = cpu.run(txt, offs, lens, sample, nthreads)
        if it >= args.warmup:
            times.append(dt)
        log(f"[reference] step {it}: {len(ids)} reads in {dt:.2f}s = {rps:.1f} reads/s ({kind}, {nthreads} threads)")
    cpu.close()
    kept = int((lens[ids] >= 500).sum())
    ms = 1e3 * float(np.mean(times))
    value = kept / (ms / 1e3)
    desc = (f"{len(ids)} evenly spaced reads of the {args.reads}-read batch per step; seed map prebuilt outside the "
            f"timed region (as for the GPU arm)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"config2: {REF_LEN} bp iid reference, {args.reads} CLR reads (mean 5 kbp, 15% error), "
                               f"mask {MASK:08x}, R={R}, locator.cpp semantics", "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--cpu-threads", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: fewer than 3 warm-up steps; the timing rules ask for W >= 3")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pacbioassembly_b200 import Context
    from pacbioassembly_b200.api import LOCATE_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the exchange's kernels run under the aligner's persistent CTAs of consecutive steps (no idle gap between steps): on a
        # high-priority stream they take the first SM slots a retiring CTA frees instead of queueing behind the aligner's blocks
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ref, txt, offs, lens = make_workload(args.reads, rank)
    nreads = len(lens)
    # pinned host copies: the e2e leg copies from these every step
    h_txt = torch.empty(len(txt), dtype=torch.uint8, pin_memory=True)
    h_txt.numpy()[:] = txt
    txt_pinned = h_txt.numpy()
    d_txt = h_txt.cuda(non_blocking=False)

    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    t0 = time.time()
    ref_set = ctx.seqset_one(ref)
    index = ctx.index(ref_set, MASK)
    index_ms = ctx.timings()
    log(f"[rank {rank}] index: {index.nentries} entries, {index.nkeys} keys in {time.time() - t0:.2f}s {index_ms}")

    kept = int((lens >= 500).sum())
    state = {}
    from pacbioassembly_b200 import shard

    reducer = shard.FinalReduction(nreads, device="cuda")
    recs_host = reducer.send_view(kept)  # pinned; the records are fetched straight into the exchange's send buffer

    def final_reduction(recs):
        """the path's only collective: hit/score counters all-reduced, records gathered on rank 0"""
        # records land in rank 0's pinned buffer; stitching them into one numpy array (FinalReduction.records()) is
        # output formatting, not part of the step
        reducer(recs, want_records=False, sync=False)  # queued on a side stream; joined by the next call / reducer.finish()
        return None


    phase_s = {"seqset": 0.0, "locate_run": 0.0, "fetch": 0.0, "reduction": 0.0, "n": 0}

    def step_device():
        t0 = time.perf_counter()
        s = ctx.seqset_from_device(d_txt.data_ptr(), d_txt.numel(), offs, lens)
        t1 = time.perf_counter()
        job = ctx.locate_run(index, s, R=R)
        t2 = time.perf_counter()
        reducer.wait()  # the previous step's exchange has long finished: its send buffer is this step's record buffer
        recs = job.fetch(recs=recs_host)
        t3 = time.perf_counter()
        t = ctx.timings()
        state["stats"], state["timings"], state["recs"] = job.stats(), t, recs
        tot = final_reduction(recs)
        t4 = time.perf_counter()
        job.free()
        s.free()
        for k, v in (("seqset", t1 - t0), ("locate_run", t2 - t1), ("fetch", t3 - t2), ("reduction", t4 - t3)):
            phase_s[k] += v
        phase_s["n"] += 1
        return tot

    # e2e leg: host text in, records out, through the pipelined entry points (pb_locate_submit / pb_locate_collect): the
    # host->device copy of step k+1 is queued before step k is collected, so it runs under step k's alignment
    pipe = {"prev": None, "mode": "text", "tot": None}
    bin_image = {"buf": None}

    def submit_e2e():
        if pipe["mode"] == "bin":
            return ctx.locate_submit_bin(index, bin_image["buf"], R=R)
        return ctx.locate_submit(index, txt_pinned, offs, lens, R=R)

    e2e_host = {"submit": 0.0, "collect": 0.0, "reduce": 0.0, "n": 0}

    def step_e2e():
        t0 = time.perf_counter()
        cur = submit_e2e()
        t1 = time.perf_counter()
        if pipe["prev"] is not None:
            reducer.wait()
            recs = pipe["prev"].collect(recs=recs_host)
            t2 = time.perf_counter()
            state["e2e_timings"] = ctx.timings()
            pipe["tot"] = final_reduction(recs)
            e2e_host["collect"] += t2 - t1
            e2e_host["reduce"] += time.perf_counter() - t2
        e2e_host["submit"] += t1 - t0
        e2e_host["n"] += 1
        pipe["prev"] = cur
        return pipe["tot"]

    def drain_e2e():
        if pipe["prev"] is not None:
            reducer.wait()
            recs = pipe["prev"].collect(recs=recs_host)
            state["e2e_timings"] = ctx.timings()
            state["e2e_recs"] = recs.copy()
            pipe["tot"] = final_reduction(recs)
            pipe["prev"] = None
        return pipe["tot"]

    walls = []  # host wall time of every call inside the last timed region (ms): shows where a region lost time

    def timed(fn, warmup, steps, collect=None, drain=None):
        for _ in range(warmup):
            fn()
        if drain is not None:
            drain()
        for k in phase_s:
            phase_s[k] = 0 if k == "n" else 0.0  # host-phase clocks cover the timed steps only
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record(stream)
        walls.clear()
        tw = time.perf_counter()
        for _ in range(steps):
            out = fn()
            if collect is not None:
                collect()
            walls.append(round(1e3 * (time.perf_counter() - tw), 2))
            tw = time.perf_counter()
        if drain is not None:
            drain()  # the last step in flight is collected inside the timed region
            walls.append(round(1e3 * (time.perf_counter() - tw), 2))
        out = reducer.finish()  # the last exchange is joined inside the timed region; the summed counters are the step's result
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, (ctx.launches - l0) // steps

    sampler = ClockSampler(local_rank)
    sampler.start()
    align_ms, dp_cells, band_cells, alu_instr, redone, stage = [], [], [], [], [], []

    def collect():
        align_ms.append(state["timings"]["align"])
        dp_cells.append(state["stats"]["dp_cells"])
        band_cells.append(state["stats"]["band_cells"])
        alu_instr.append(state["stats"]["alu_instr"])
        redone.append(state["stats"]["redone"])
        stage.append(dict(state["timings"]))

    # device leg: the same batches with their text resident in HBM, streamed through the pipelined entry points
    # (pb_locate_submit_device / pb_locate_collect: no copy in, records out); K steps and the drain inside the timed region.
    # The strictly serial call sequence (seqset -> locate_run -> fetch, the GPU idle while the host plans) is timed beside it.
    dpipe = {"prev": None, "tot": None}

    def collect_device(step):
        reducer.wait()
        recs = step.collect(recs=recs_host)
        state["stats"], state["timings"], state["recs"] = step.stats, ctx.timings(), recs
        collect()
        dpipe["tot"] = final_reduction(recs)

    def step_device_pipe():
        cur = ctx.locate_submit_device(index, d_txt.data_ptr(), d_txt.numel(), offs, lens, R=R)
        if dpipe["prev"] is not None:
            collect_device(dpipe["prev"])
        dpipe["prev"] = cur
        return dpipe["tot"]

    def drain_device():
        if dpipe["prev"] is not None:
            collect_device(dpipe["prev"])
            dpipe["prev"] = None
        return dpipe["tot"]

    ms_dev, tot_dev, launches = timed(step_device_pipe, args.warmup, args.steps, drain=drain_device)
    dev_walls = list(walls)
    for lst in (align_ms, dp_cells, band_cells, alu_instr, redone, stage):
        del lst[:-args.steps]  # the timed steps are the last K collected
    dev_recs = state["recs"].copy()  # the device leg's records (recs_host is reused by the other legs)
    ms_serial, _, _ = timed(step_device, 2, args.steps)
    state["recs"] = dev_recs
    log(f"[rank {rank}] host wall per serial device-leg step (ms): " + ", ".join(f"{k} {1e3 * v / max(phase_s['n'], 1):.1f}" for k, v in phase_s.items() if k != "n") + "\n")
    ms_e2e, tot_e2e, _ = timed(step_e2e, max(3, args.warmup), args.steps, drain=drain_e2e)
    e2e_walls = list(walls)
    log(f"[rank {rank}] host wall per e2e step (ms): " + ", ".join(f"{k} {1e3 * v / max(e2e_host['n'], 1):.1f}" for k, v in e2e_host.items() if k != "n"))
    e2e_same = bool((state["e2e_recs"] == state["recs"]).all()) if "e2e_recs" in state else None
    e2e_stage = state.get("e2e_timings")
    # the same batch as a .bin image (binary_test.cpp:55-63): a quarter of the bytes to copy
    ms_bin, bin_same = None, None
    try:
        import workload
        img = workload.pack_bin(txt, offs, lens)
        bin_image["buf"] = torch.empty(len(img), dtype=torch.uint8, pin_memory=True).numpy()
        bin_image["buf"][:] = img
        pipe["mode"] = "bin"
        ms_bin, tot_bin, _ = timed(step_e2e, 2, args.steps, drain=drain_e2e)
        bin_same = bool((state["e2e_recs"] == state["recs"]).all())
    except Exception as e:
        log(f"[rank {rank}] .bin e2e leg skipped: {e}")
    clocks = sampler.stop()

    total_reads = int(tot_dev[3])  # kept reads over all ranks
    value = total_reads / (ms_dev / 1e3)
    e2e_value = int(tot_e2e[3]) / (ms_e2e / 1e3)

    # roofline of the dominant kernel (K3 banded aligner): 2 parent bits written per band cell it computes.  Its first pass
    # computes a certified strip of the reference's band, so two cell counts exist: the strip cells it
    # really computes (the roofline's algorithmic bytes) and the cells the reference's recurrence would have filled for the
    # same alignments (the GCUPS the north star asks for: work delivered, in the reference's unit).
    peak, peak_src = peaks()
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one step's K3 launches, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r02_k3_traffic.json")) as f:
            tj = json.load(f)
        # the capture names the source files it was taken with; a stale one is not quoted
        if args.reads == int(tj.get("reads", 0)) and tj.get("source_sha") == source_sha():
            traffic, traffic_src = float(tj["dram_bytes_total"]), "profiles/r02_k3_traffic.json (ncu, same workload and kernel sources, one step)"
        else:
            traffic_src = "profiles/r02_k3_traffic.json is from other kernel sources or another workload: not quoted"
    except Exception:
        pass
    k3_ms = float(np.mean(align_ms))
    k3_cells = float(np.mean(dp_cells))
    k3_band = float(np.mean(band_cells))
    alg_bytes = 0.25 * k3_band
    achieved = alg_bytes / (k3_ms / 1e3) / 1e9
    gcups = k3_cells / (k3_ms / 1e3) / 1e9
    # integer pipe: ALU warp instructions of K3's row loops (rows x the per-row count read off each band class's SASS)
    # against a measured LOP3/SHF peak on the same GPU
    int_peak = ctx.int_pipe_peak()
    int_ach = float(np.mean(alu_instr)) / (k3_ms / 1e3)
    int_pipe = {"achieved_warp_instr_s": int_ach, "peak_warp_instr_s": int_peak, "frac": int_ach / int_peak if int_peak else None,
                "achieved_ops_s": 32.0 * int_ach, "peak_ops_s": 32.0 * int_peak,
                "how": "achieved = integer-ALU warp instructions of K3's row loops (11 per band word + 12 per row in the strip pass, "
                       "19 + 23 in the full-band pass, counted in the SASS) / K3 time; peak = pb_int_pipe_peak, a register-only "
                       "LOP3/SHF kernel timed on this GPU in this run"}

    # K1 bulk seed extraction over every position of the read set: 0.25 B read + 4 B written per position
    rs = ctx.seqset_from_device(d_txt.data_ptr(), d_txt.numel(), offs, lens)
    seed_gbs = None
    for _ in range(4):
        nk, ms = rs.seeds_device(MASK)
        seed_gbs = 4.25 * nk / (ms / 1e3) / 1e9
    # K2 bulk: every read position probed against the index (count pass and gather pass timed separately)
    probe = None
    try:
        pbk = index.probe_bulk(rs)
        pbk = index.probe_bulk(rs)
        q, cnd = pbk["queries"], pbk["candidates"]
        count_gbs = 12.0 * q / (pbk["count_ms"] / 1e3) / 1e9
        gather_gbs = (12.0 * q + 12.0 * cnd) / (pbk["gather_ms"] / 1e3) / 1e9
        rg = ctx.random_gather_peak(64 << 20)  # pure random 4-byte reads of a 64 MB table: the request-rate ceiling
        probe = {"queries": q, "candidates": cnd, "count_ms": pbk["count_ms"], "gather_ms": pbk["gather_ms"],
                 "count_gbs": count_gbs, "gather_gbs": gather_gbs, "count_frac_of_peak": count_gbs / peak,
                 "gather_frac_of_peak": gather_gbs / peak,
                 "random_gather_peak_reads_s": rg, "count_probes_s": q / (pbk["count_ms"] / 1e3),
                 "count_frac_of_random_gather_peak": q / (pbk["count_ms"] / 1e3) / rg,
                 "bound": "the SMs' random-request rate (one divergent 32-byte sector per probe), measured by "
                          "pb_random_gather_peak: independent random reads of a 64 MB table with no other traffic",
                 "algorithmic": "count: 4 B key + 8 B bucket header per query; gather: the same + 4 B per position read + "
                                "8 B per candidate written; random 8-byte reads of a 64 MB bucket table (L2-resident)"}
    except Exception as e:  # the bulk leg needs ~10 GB of scratch; never let it take the headline down
        probe = {"error": str(e)}
    rs.free()

    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:  # every N: rank 0 checks a sample of ITS shard against the reference on the host
        nthreads = min(os.cpu_count() or 1, args.cpu_threads)
        sample = args.cpu_sample or 48 * nthreads
        cr = CpuReference(ref)
        rps, crecs, ids, dt = cr.run(txt, offs, lens, sample, nthreads)
        cr.close()
        if world == 1:
            cpu = {"value": rps, "unit": UNIT, "cores": nthreads, "kind": cr.kind,
                   "sample": f"{len(ids)} evenly spaced reads of the batch in {dt:.1f}s (seed map prebuilt in {cr.build_s:.1f}s, not timed)"}
        # The same reads through the GPU path must give the same records.  The compiled reference is only a valid
        # witness where its band fits its matrix row (2*max_dst < MAXM = 6000): beyond that its cells alias the next
        # row and its answers depend on what earlier alignments left there (SURVEY Q-D1), so those reads are checked
        # against the C port of the clean recurrence instead.
        import cpu_libs
        kept_rank = np.cumsum(lens >= 500) - 1
        sid = ids[lens[ids] >= 500]
        g = state["recs"][kept_rank[sid]]
        fields = ["found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand"]
        in_domain = 2 * (1 + (lens[sid] * R).astype(np.int64)) < 6000
        same_ref = all((g[n][in_domain] == crecs[n][in_domain]).all() for n in fields) if cr.kind == "reference" else None
        o = cpu_libs.oracle()
        oix = o.index_build(ref, MASK, 0)
        s_lens = lens[ids]
        s_offs = np.zeros(len(ids), dtype=np.int64)
        np.cumsum(s_lens[:-1], out=s_offs[1:])
        s_txt = np.concatenate([txt[offs[i]: offs[i] + lens[i]] for i in ids])
        precs = o.locate(oix, ref, s_txt, s_offs, s_lens, MASK, R=R, nthreads=nthreads)
        o.index_free(oix)
        same_port = all((g[n] == precs[n]).all() for n in fields + ["cells"])
        parity = {"reads_checked": int(len(g)), "bit_exact_vs_port": bool(same_port),
                  "reads_in_reference_domain": int(in_domain.sum()), "bit_exact_vs_reference": same_ref,
                  "e2e_records_equal_device_leg": e2e_same, "rank": 0, "n_gpus": world}

    if rank == 0:
        t = stage[-1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"config2: {REF_LEN} bp iid reference, {args.reads} CLR reads per GPU (mean 5 kbp, "
                                   f"ins 9/del 4/sub 2 %), mask {MASK:08x}, R={R}, ntrial 50, locator.cpp semantics",
                       "l2": "inputs (0.5 GB of reads per step) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"reads sharded over {world} GPU(s), index replicated"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(len(txt) + offs.nbytes + lens.nbytes),
                    "d2h_bytes_per_step": int(recs_host.nbytes), "ms_per_step": ms_e2e,
                    "how": "pb_locate_submit / pb_locate_collect from pinned host text: step k+1 is submitted before step k is "
                           "collected, so its host->device copy runs under step k's alignment; every step's records are "
                           "copied back and reduced inside the timed region",
                    "stage_ms": e2e_stage,
                    "call_wall_ms": e2e_walls,  # host wall of each step's submit+collect (and of the final drain) in the timed region
                    "bin_input": None if ms_bin is None else {
                        "value": total_reads / (ms_bin / 1e3), "ms_per_step": ms_bin,
                        "h2d_bytes_per_step": int(len(bin_image["buf"])), "records_equal_device_leg": bin_same,
                        "what": "the same batch handed over as a .bin image (binary_test.cpp:55-63) through pb_locate_submit_bin"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "align_locate_nb_kernel<S> + align_locate_kernel<S> (K3: banded bit-parallel DP + traceback, "
                                   "certified strip pass then full-band pass for what it could not certify)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_ms_per_step": [round(float(x), 2) for x in align_ms],
                         "peak_source": peak_src,
                         "algorithmic": "0.25 B (2 parent bits) per band cell K3 computes (strip width x rows, summed over the "
                                        "alignments it ran)",
                         "kernel_ms": k3_ms, "band_cells_per_step": k3_band, "ref_equiv_cells_per_step": k3_cells,
                         "ref_equiv_frac": 0.25 * k3_cells / (k3_ms / 1e3) / 1e9 / peak,
                         "reads_redone_full_band_per_step": float(np.mean(redone))},
            "gcups": gcups,
            "gcups_note": "reference-equivalent: cells seq_aligner.h:151-190 fills for the alignments K3 ran / K3 time",
            "gcups_band": k3_band / (k3_ms / 1e3) / 1e9,
            "int_pipe": int_pipe,
            "seed_extract": {"achieved_gbs": seed_gbs, "frac_of_peak": seed_gbs / peak if seed_gbs else None,
                             "algorithmic": "0.25 B read + 4 B written per position, every position of the read set"},
            "probe_bulk": probe,
            "stage_ms": t,
            "device_leg": {"how": "text resident in HBM, pb_locate_submit_device / pb_locate_collect: step k+1 is queued before "
                                  "step k is collected; K steps and the drain inside the timed region, every step's records "
                                  "copied back and reduced", "call_wall_ms": dev_walls,
                           "serial": {"ms_per_step": ms_serial, "value": total_reads / (ms_serial / 1e3),
                                      "how": "pb_seqset_from_device_text -> pb_locate_run -> pb_locate_fetch, one after the "
                                             "other: the GPU idles while the host plans and reduces"}},
            "mapped_reads": int(tot_dev[0]), "kept_reads": total_reads, "sum_cost": int(tot_dev[1]),
            "ref_equiv_cells": int(tot_dev[2]),
            "cpu_baseline": cpu,
            "parity_vs_reference_cpu": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
