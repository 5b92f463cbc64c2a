#!/usr/bin/env python
"""Benchmark of the block-decode hot path (BASELINE.json metric: 4x4 blocks/s & HBM GB/s, BC1 + BC7,
1/2/4/8 GPU vs CPU ref).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|bc1|bc7|bc7u|bc6h|mixed|bc1_1e8]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's own CPU decode on the host cores

A "step" is one pass of the hot path over one batch of synthetic blocks (the whole texture of the
workload).  The headline leg is BASELINE.json configs[1]: BC1 8192x8192 (4 194 304 blocks, distribution U)
through the LINEAR layout, native pixel format; with the default `--workload all` the same run also times the
other configurations of the metric and reports them inside the same JSON line:

  value        whole-job blocks/s of the headline leg with inputs and outputs resident in HBM, timed on the
               device (CUDA events on the launching stream, barrier + synchronize on both sides, max over
               ranks).  Buffers rotate over sets totalling >> L2 so nothing is served from cache.
  roofline     HBM roofline of the headline decode kernel: algorithmic bytes per launch (72 B/block for
               BC1) / average launch duration, against MEASURED_PEAKS.json.
  per_config   the same measurement (value, ms, roofline, parity, e2e) for BC7 16384x16384 (configs[2],
               distribution M, LINEAR) and BC6H 8192x8192 (configs[3]) -- weak scaling like the headline.
  strong_1e8   strong scaling: 10^8 blocks in total, index-range sharded over the ranks, for
               BC1 (the north-star's scaling target) and for the mixed ETC2 + EAC_RG11 + RGTC2 batch
               (configs[4]); every rank's shard is hashed on the device and compared with the hash of the CPU
               checker's output for the same shard, at every N.
  e2e          the headline metric through the reference-facing C-ABI call detexDecompressTextureLinear with
               HOST buffers: host->device copy of the blocks and device->host copy of the pixels inside the
               timed region, every step.  `value` is the pinned-buffer figure; `pageable` is the same call on
               malloc'd (numpy) buffers -- what the reference's own callers pass (detex-convert.c:303-310) --
               and `pcie_ceiling` a plain cudaMemcpy of the same bytes on the same box.
  cpu_baseline the reference's CPU decode (oracle/_ref, else the oracle port) on this box's host cores, warm
               output buffer, persistent threads, on the same texture (rank 0, N=1 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload -> parts: (format, width, height, layout, seed, force_modes)
WORKLOADS = {
    "bc1": dict(name="BC1 8192x8192 synthetic (BASELINE configs[1]), distribution U, LINEAR",
                parts=[("BC1", 8192, 8192, "linear", 1, 0)]),
    "bc7": dict(name="BC7/BPTC 16384x16384 synthetic (BASELINE configs[2]), distribution M, LINEAR",
                parts=[("BPTC", 16384, 16384, "linear", 1, 8)]),
    "bc7u": dict(name="BC7/BPTC 16384x16384 synthetic, distribution U, LINEAR",
                 parts=[("BPTC", 16384, 16384, "linear", 1, 0)]),
    "bc6h": dict(name="BC6H/BPTC_FLOAT 8192x8192 synthetic (BASELINE configs[3]), distribution M, LINEAR",
                 parts=[("BPTC_FLOAT", 8192, 8192, "linear", 1, 14)]),
    "mixed": dict(name="ETC2 + EAC_RG11 + RGTC2, 1e8 blocks total over all ranks (BASELINE configs[4]), TILED",
                  parts=[("ETC2", 0, 0, "tiled", 1, 0), ("EAC_RG11", 0, 0, "tiled", 2, 0), ("RGTC2", 0, 0, "tiled", 3, 0)],
                  total_blocks=100_000_000, strong=True),
    "bc1_1e8": dict(name="BC1 1e8 blocks total over all ranks (scaling config), TILED",
                    parts=[("BC1", 0, 0, "tiled", 1, 0)], total_blocks=100_000_000, strong=True),
}
HEADLINE = "bc1"
ALGO_BYTES = {"BC1": 72, "BC1A": 72, "ETC1": 72, "ETC2": 72, "ETC2_PUNCHTHROUGH": 72, "BC2": 80, "BC3": 80,
              "BPTC": 80, "ETC2_EAC": 80, "BPTC_FLOAT": 144, "BPTC_SIGNED_FLOAT": 144, "RGTC1": 24, "RGTC2": 48,
              "SIGNED_RGTC1": 40, "EAC_R11": 40, "EAC_SIGNED_R11": 40, "SIGNED_RGTC2": 80, "EAC_RG11": 80,
              "EAC_SIGNED_RG11": 80}
SORTED_FORMATS = ("BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key, build_id):
    """(dram bytes per launch, note) from the committed ncu capture of this kernel -- only when the capture was
    taken on the build that is loaded now (profiles/ncu_traffic.json records the build id of every figure,
    tools/ncu_traffic.py writes it); a figure from another build is dropped, with the reason."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        e = json.load(open(p)).get(kernel_key)
    except Exception:
        return None, "profiles/ncu_traffic.json missing or unreadable"
    if not isinstance(e, dict):
        return None, "no ncu capture of this kernel recorded"
    if e.get("build") != build_id:
        return None, "stale: captured on build %s, loaded library is build %s" % (e.get("build"), build_id)
    return e.get("bytes"), "ncu --set full capture %s on this build" % e.get("source", "?")


def kernel_name(fmt, layout_name):
    if fmt in SORTED_FORMATS:       # mode-sorted persistent kernel
        return "decode_sorted_kernel<%s,%s>" % (fmt, layout_name)
    return "decode_%s_kernel<%s>" % (layout_name, fmt)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load: power above the idle floor of this window
        thr = min(power) + 0.3 * (max(power) - min(power))
        loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power),
                "window": "warm-up + timed region of the headline leg + ~1 s sustained loop of the same kernel"}


def shard(total, rank, world):
    """Contiguous index range [lo, hi) of rank."""
    return total * rank // world, total * (rank + 1) // world


def part_share(wl, index):
    """blocks of part `index` of a flat (strong-scaling) workload: equal shares, the remainder to the first parts"""
    k = len(wl["parts"])
    return wl["total_blocks"] // k + (1 if index < wl["total_blocks"] % k else 0)


def config_dict(wl_key, world):
    """`config` of the JSON line: BOTH arms print exactly this dict for the same command line."""
    wl = WORKLOADS[wl_key]
    fmt, w, h, layout, _, _ = wl["parts"][0]
    if wl.get("strong"):
        per_gpu = wl["total_blocks"] // world
    else:
        per_gpu = (w // 4) * (h // 4)
    return {"workload": wl["name"], "blocks_per_step_per_gpu": int(per_gpu), "pixel_format": "native",
            "layout": layout, "sharding": "index range per rank, no data-path collective",
            "cache": "rotating buffer sets totalling > 600 MB per rank (>> 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world, out):
    """bench.py --impl reference: the reference's own CPU implementation on the host cores, warm buffers,
    persistent threads."""
    if rank != 0:
        return
    from oracle import harness as H
    lib, kind = H.best_cpu_decoder()
    wl_key = HEADLINE if args.workload == "all" else args.workload
    wl = WORKLOADS[wl_key]
    threads = os.cpu_count() or 1
    parts = []
    total_blocks = 0
    for (fmt, w, h, layout, seed, force) in wl["parts"]:
        if w == 0:      # flat arrays: bounded sample of 2^22 blocks per format
            w, h = 8192, 8192 // (2 if H.FORMATS[fmt][1] == 16 else 1)
        if fmt == "BPTC" and w == 16384:
            w, h = 8192, 8192       # bounded sample: a quarter of the 16384^2 texture
        n = (w // 4) * (h // 4)
        b = H.synth_blocks(fmt, n, seed=seed)
        if force:
            b = H.force_modes(fmt, b)
        ps = H.FORMATS[fmt][2]
        parts.append((fmt, b, w, h, layout == "linear", np.zeros(n * 16 * ps, dtype=np.uint8)))
        total_blocks += n

    def step():
        t = 0.0
        for (fmt, b, w, h, linear, buf) in parts:
            _, _, s = H.cpu_decode(lib, fmt, b, w, h, linear=linear, threads=threads, out=buf)
            t += s
        return t
    for _ in range(args.warmup):
        step()
    secs = sum(step() for _ in range(args.steps))
    value = total_blocks * args.steps / secs
    sample = "; ".join("%s %dx%d (%d blocks)" % (p[0], p[2], p[3], (p[2] // 4) * (p[3] // 4)) for p in parts)
    line = {"impl": "reference", "metric": "4x4 blocks/s", "value": value, "unit": "blocks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_dict(wl_key, world),
            "cpu_baseline": {"value": value, "unit": "blocks/s", "cores": threads, "kind": kind,
                             "sample": sample + "; warm output buffer, persistent threads, every step decodes this sample"},
            "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    out.emit(json.dumps(line))


class QuietStdout:
    """Everything that libraries print to fd 1 (e.g. the NCCL version banner) goes to stderr, so
    that stdout carries exactly one line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    with QuietStdout() as out:
        run(out)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import detex_b200 as D
        self.torch, self.dist, self.D = torch, dist, D
        self.args, self.rank, self.world = args, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.lib = D.load()
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.peak, self.peak_src = measured_peak()
        self.build = D.build_id()
        self.cores = os.cpu_count() or 1
        if world > 1:       # the ranks share the host cores: each library instance gets its share of copy threads
            D.set_host_threads(max(4, self.cores // world))

    # ---- device-resident inputs of one workload: rotating buffer sets totalling >> L2 (126 MB) -------------------
    def build_parts(self, wl):
        torch, D = self.torch, self.D
        parts = []
        for index, (fmt, w, h, layout, seed, force) in enumerate(wl["parts"]):
            tf = D.FORMATS[fmt]
            bs, ps = D.block_size(tf), D.pixel_size(tf & 0xFFFF)
            if w == 0:          # flat array, index-range sharded over the ranks (strong scaling)
                lo, hi = shard(part_share(wl, index), self.rank, self.world)
                lo, hi = lo & ~1, (hi & ~1) if self.rank + 1 < self.world else hi      # 16-byte units between shards
                n, first = hi - lo, lo
                wib, hib, w_px, h_px = n, 1, 4 * n, 4
            else:               # one texture per rank (weak scaling)
                wib, hib = w // 4, h // 4
                n, first = wib * hib, self.rank * wib * hib
                w_px, h_px = w, h
            n_gen = n + (n & 1 if bs == 8 else 0)       # the generator works in 16-byte units
            set_bytes = n * (bs + 16 * ps)
            n_sets = max(1 if set_bytes > 600e6 else 2, min(8, int(np.ceil(600e6 / set_bytes))))
            sets = []
            for k in range(n_sets):
                blocks = torch.empty(n_gen * bs, dtype=torch.uint8, device=self.dev)
                D.generate_blocks_device(blocks.data_ptr(), n_gen, bs, seed=seed + 1000 * k, first_block=first,
                                         force_modes=force, stream=self.sp)
                pixels = torch.empty(n * 16 * ps, dtype=torch.uint8, device=self.dev)
                sets.append((blocks, pixels))
            parts.append(dict(fmt=fmt, tf=tf, n=n, wib=wib, hib=hib, w=w_px, h=h_px, bs=bs, ps=ps, sets=sets, seed=seed,
                              force=force, first=first, lname=layout,
                              layout=D.LAYOUT_LINEAR if layout == "linear" else D.LAYOUT_TILED))
        torch.cuda.synchronize()
        return parts

    def decode_part(self, p, k):
        blocks, pixels = p["sets"][k % len(p["sets"])]
        self.D.decode_device(p["tf"], blocks.data_ptr(), p["w"], p["h"], pixels.data_ptr(), layout=p["layout"],
                             d_invalid=self.flag.data_ptr(), stream=self.sp, wib=p["wib"], hib=p["hib"])

    def reduce_max(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def reduce_all(self, ok):
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- device-timed leg: W warm-up steps, K timed steps, barrier + synchronize on both sides, max over ranks ----
    def time_device(self, parts, steps, warmup):
        torch, D = self.torch, self.D

        def one_step(k):
            for p in parts:
                self.decode_part(p, k)
        for k in range(warmup):
            one_step(k)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        launches0 = D.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for k in range(steps):
            one_step(warmup + k)
        e1.record(self.stream)
        torch.cuda.synchronize()
        launches = D.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            self.dist.barrier()
        ms_max = self.reduce_max(ms)
        job_blocks = self.reduce_sum(sum(p["n"] for p in parts))
        # multi-part steps: time each part's kernel on its own (events on the launching stream) so that the
        # roofline of the dominant kernel uses its measured duration, not a byte-share estimate
        part_ms = None
        if len(parts) > 1:
            part_ms = []
            for p in parts:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(self.stream)
                for k in range(steps):
                    self.decode_part(p, k)
                a1.record(self.stream)
                torch.cuda.synchronize()
                part_ms.append(self.reduce_max(a0.elapsed_time(a1) / steps))
        return dict(ms_per_step=ms_max / steps, job_blocks=job_blocks, launches=int(launches), part_ms=part_ms,
                    value=job_blocks * steps / (ms_max / 1e3))

    def roofline(self, parts, timing):
        """roofline of the dominant kernel of the leg: per-rank algorithmic bytes per launch over its average
        launch duration"""
        dom = max(parts, key=lambda p: p["n"] * ALGO_BYTES[p["fmt"]])
        per_launch_ms = timing["part_ms"][parts.index(dom)] if timing["part_ms"] is not None else timing["ms_per_step"]
        achieved = dom["n"] * ALGO_BYTES[dom["fmt"]] / (per_launch_ms / 1e3) / 1e9
        kname = kernel_name(dom["fmt"], dom["lname"])
        traffic, note = ncu_traffic(kname, self.build)
        return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                "frac": achieved / self.peak, "peak_source": self.peak_src,
                "algorithmic_bytes_per_block": ALGO_BYTES[dom["fmt"]],
                "algorithmic_bytes_per_launch": int(dom["n"] * ALGO_BYTES[dom["fmt"]]),
                "traffic": traffic, "traffic_note": note}

    # ---- parity of a leg: this rank's shard against the CPU checker ---------------------------------------------
    def parity(self, parts, full_compare):
        """Decode set 0 of every part once more and compare with the checker's output for the same blocks: by the
        device-side hash of hashes (detexb200HashDevice vs the same construction on the host), and byte for byte
        when full_compare.  Every rank checks its own shard; the verdict is AND-reduced."""
        from oracle import harness as H
        torch, D = self.torch, self.D
        cpu_lib, kind = H.best_cpu_decoder()
        threads = max(1, self.cores // self.world)
        ok_all, checked, cpu_secs, cpu_blocks = True, 0, 0.0, 0
        for p in parts:
            blocks, pixels = p["sets"][0]
            self.flag.zero_()
            self.decode_part(p, 0)
            torch.cuda.synchronize()
            gpu_valid = bool(self.flag.item() == 0)
            n, bs, ps, fmt = p["n"], p["bs"], p["ps"], p["fmt"]
            # the checker's input is generated on the host from the same counter-based generator
            hb = H.synth_blocks(fmt, n, seed=p["seed"], first_block=p["first"])
            if p["force"]:
                hb = self._force_modes_at(H, fmt, hb, p["first"])
            same_input = torch.equal(blocks[: n * bs][: 1 << 22].cpu(), torch.from_numpy(hb[: 1 << 22].copy()))
            linear = p["layout"] == D.LAYOUT_LINEAR
            if linear:
                want, cpu_valid, s = H.cpu_decode(cpu_lib, fmt, hb, p["w"], p["h"], linear=True, threads=threads)
            else:   # flat array: rows of 1024 blocks + a ragged tail
                rows = n // 1024
                want = np.empty(n * 16 * ps, dtype=np.uint8)
                cpu_valid, s = True, 0.0
                if rows:
                    _, cpu_valid, s = H.cpu_decode(cpu_lib, fmt, hb[: rows * 1024 * bs], 4096, 4 * rows, linear=False,
                                                   threads=threads, out=want[: rows * 1024 * 16 * ps])
                if n > rows * 1024:
                    _, v2, s2 = H.cpu_decode(cpu_lib, fmt, hb[rows * 1024 * bs:], 4 * (n - rows * 1024), 4, linear=False,
                                             out=want[rows * 1024 * 16 * ps:])
                    cpu_valid, s = cpu_valid and v2, s + s2
            cpu_secs += s
            cpu_blocks += n
            n_bytes = want.size
            same = same_input and gpu_valid == cpu_valid and \
                D.hash_device(pixels.data_ptr(), n_bytes, stream=self.sp) == H.tree_hash(want, threads=threads)
            if same and full_compare:
                same = bool(np.array_equal(pixels[:n_bytes].cpu().numpy(), want))
            ok_all = ok_all and same
            checked += n
            del want, hb
        return {"bit_exact": self.reduce_all(ok_all), "against": kind, "checked_blocks": int(self.reduce_sum(checked)),
                "how": ("per-rank shard: device hash of hashes vs the checker's" + (", and byte for byte" if full_compare else "")),
                "ranks": self.world}, (cpu_blocks / cpu_secs if cpu_secs > 0 else None), threads

    @staticmethod
    def _force_modes_at(H, fmt, hb, first):
        """distribution M for a shard that starts at global block `first`: block i has mode (first + i) % n_modes"""
        n_modes = 8 if fmt == "BPTC" else 14
        shift = first % n_modes
        modes = [(shift + k) % n_modes for k in range(n_modes)]
        return H.force_modes(fmt, hb, modes)

    def host_copy_gbs(self):
        """GB/s of plain host copies, one 64 MiB buffer per thread, all host threads at once (rank 0's share of the box
        when several ranks run): the ceiling of any path that has to copy the pixels once on the host"""
        threads = max(1, self.cores // self.world)
        src = [np.ones(64 << 20, dtype=np.uint8) for _ in range(threads)]
        dst = [np.empty(64 << 20, dtype=np.uint8) for _ in range(threads)]

        def work(i):
            for _ in range(3):
                np.copyto(dst[i], src[i])
        for i in range(threads):        # touch the pages
            dst[i][:] = 0
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return threads * 3 * (64 << 20) / (time.perf_counter() - t0) / 1e9

    # ---- end-to-end legs: reference-facing API, host buffers, copies inside the timed region ---------------------
    def e2e(self, parts, steps):
        torch, D, lib = self.torch, self.D, self.lib
        h2d = d2h = 0
        geo = []
        for p in parts:
            n_out = (p["w"] * p["h"] if p["layout"] == D.LAYOUT_LINEAR else p["n"] * 16) * p["ps"]
            geo.append((p, int(p["n"] * p["bs"]), int(n_out)))
            h2d += int(p["n"] * p["bs"])
            d2h += int(n_out)

        def make_calls(pinned):
            calls = []
            for p, n_in, n_out in geo:
                if pinned:
                    hb = torch.empty(n_in, dtype=torch.uint8).pin_memory()
                    hb.copy_(p["sets"][0][0][:n_in])
                    hp = torch.empty(n_out, dtype=torch.uint8).pin_memory()
                    in_ptr, out_ptr, view = hb.data_ptr(), hp.data_ptr(), hp.numpy()
                else:       # pageable: plain malloc'd memory, as the reference's callers pass (ktx.c:142)
                    hb = np.empty(n_in, dtype=np.uint8)
                    hb[:] = p["sets"][0][0][:n_in].cpu().numpy()
                    hp = np.zeros(n_out, dtype=np.uint8)            # (touched: warm pages, like the CPU baseline's)
                    in_ptr, out_ptr, view = hb.ctypes.data, hp.ctypes.data, hp
                tex = D.DetexTexture(p["tf"], in_ptr, p["w"], p["h"], p["wib"], p["hib"])
                fn = lib.detexDecompressTextureLinear if p["layout"] == D.LAYOUT_LINEAR else lib.detexDecompressTextureTiled
                calls.append((fn, tex, out_ptr, p["tf"] & 0xFFFF, view, (hb, hp)))
            return calls

        def timed(calls):
            def step():
                for fn, tex, out_ptr, pf, _, _ in calls:
                    if not fn(C.byref(tex), out_ptr, pf) and not D.detexGetErrorMessage().startswith("detexDecompressBlock"):
                        raise RuntimeError("e2e call failed: " + D.detexGetErrorMessage())
            for _ in range(3):
                step()
            if self.world > 1:
                self.dist.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            return self.reduce_max(time.perf_counter() - t0)

        job_blocks = self.reduce_sum(sum(p["n"] for p in parts))
        host_copy = self.host_copy_gbs()
        res = {}
        views = {}
        for kind in ("pinned", "pageable"):
            calls = make_calls(kind == "pinned")
            t = timed(calls)
            res[kind] = {"value": job_blocks * steps / t, "ms_per_step": 1e3 * t / steps}
            views[kind] = calls[0][4]
        # plain cudaMemcpy of the same bytes (pinned, H2D and D2H concurrently on two streams): what the link gives
        hin = torch.empty(h2d, dtype=torch.uint8).pin_memory()
        hout = torch.empty(d2h, dtype=torch.uint8).pin_memory()
        din = torch.empty(h2d, dtype=torch.uint8, device=self.dev)
        dout = torch.empty(d2h, dtype=torch.uint8, device=self.dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def memcpy_step():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
        for _ in range(2):
            memcpy_step()
        if self.world > 1:
            self.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            memcpy_step()
        t_copy = self.reduce_max(time.perf_counter() - t0)
        ceiling = job_blocks * steps / t_copy
        p0 = parts[0]
        out = {"value": res["pinned"]["value"], "unit": "blocks/s", "steps": steps,
               "ms_per_step": res["pinned"]["ms_per_step"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "detexDecompressTexture%s per part, pinned host buffers, wall clock around the synchronous calls" %
                      ("Linear" if p0["layout"] == D.LAYOUT_LINEAR else "Tiled"),
               "pageable": {"value": res["pageable"]["value"], "ms_per_step": res["pageable"]["ms_per_step"],
                            "buffers": "malloc'd (numpy) block and pixel arrays, warm pages; the library stages them "
                                       "through the pinned bounce buffers of its copy pool",
                            "frac_of_pinned": res["pageable"]["value"] / res["pinned"]["value"],
                            "gbs_out": d2h * self.world / (res["pageable"]["ms_per_step"] / 1e3) / 1e9,
                            "host_copy_ceiling_gbs": host_copy,
                            "bound": "host memory: every pixel is written by the DMA engine into a pinned bounce buffer and copied "
                                     "once more by the pool's threads; host_copy_ceiling_gbs is what plain multi-threaded memcpy "
                                     "reaches on this box with all cores and no DMA beside it"},
               "pcie_ceiling": {"value": ceiling, "ms_per_step": 1e3 * t_copy / steps,
                                "gbs_d2h": d2h * self.world * steps / t_copy / 1e9,
                                "how": "cudaMemcpyAsync of the same bytes, pinned, H2D and D2H on two streams, same box, all ranks "
                                       "at once; gbs_d2h = pixels of all ranks per second"},
               "frac_of_pcie": res["pinned"]["value"] / ceiling,
               "pageable_frac_of_pcie": res["pageable"]["value"] / ceiling}
        return out, views


def run(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS),
                    help="all (default): the headline leg (bc1) plus per_config (bc7, bc6h) and strong_1e8 (bc1_1e8, mixed)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer legs (default min(steps, 10))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and the parity legs")
    ap.add_argument("--no-extra", action="store_true", help="with --workload all: headline leg only")
    ap.add_argument("--gather", action="store_true",
                    help="N > 1: also time the optional final-image gather (NCCL all-gather of the decoded slabs), reported apart")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world, out)
        return
    head_key = HEADLINE if args.workload == "all" else args.workload
    # The CPU baseline is the --impl reference arm run as a child process FIRST, while this process is still small and
    # the box idle: measured later, beside gigabytes of pinned and device allocations of the parent, the same child gave
    # anything between 0.6x and 1.0x of the arm the driver runs on its own (round 2: 285 M vs 480 M blocks/s on one box).
    pre_ref = reference_child(head_key) if (not args.no_cpu and rank == 0 and world == 1) else None

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = Bench(args, rank, world, local_rank)
    D = B.D
    wl = WORKLOADS[head_key]
    strong = bool(wl.get("strong"))
    e2e_steps = args.e2e_steps or min(args.steps, 10)

    # ---- headline leg ---------------------------------------------------------------------------------------------
    parts = B.build_parts(wl)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timing = B.time_device(parts, args.steps, args.warmup)
    gather = gather_leg(B, parts, timing, args) if (args.gather and world > 1) else None
    if rank == 0:       # sustained loop for the clock sampler (not part of any reported number)
        t_end = time.time() + 1.0
        k = 0
        while time.time() < t_end:
            for _ in range(20):
                for p in parts:
                    B.decode_part(p, k)
                k += 1
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    roof = B.roofline(parts, timing)
    e2e, e2e_views = B.e2e(parts, e2e_steps)

    # ---- CPU baseline (rank 0, N = 1 only) + parity of the timed configuration (every rank, every N) ---------------
    cpu_baseline = None
    parity = None
    if not args.no_cpu:
        parity, _, _ = B.parity(parts, full_compare=not strong)
        if rank == 0 and world == 1:
            cpu_baseline, want0 = cpu_baseline_leg(B, parts[0], pre_ref)
            if want0 is not None:       # the e2e paths wrote host buffers: compare them too
                parity["e2e_pinned_bit_exact"] = bool(np.array_equal(e2e_views["pinned"][: want0.size], want0))
                parity["e2e_pageable_bit_exact"] = bool(np.array_equal(e2e_views["pageable"][: want0.size], want0))
    del e2e_views
    step_bytes = sum(p["n"] * ALGO_BYTES[p["fmt"]] for p in parts)
    per_part = None if timing["part_ms"] is None else [
        {"format": p["fmt"], "blocks": int(p["n"]), "ms": m, "gblocks_s": p["n"] / m / 1e6,
         "gbs": p["n"] * ALGO_BYTES[p["fmt"]] / m / 1e6} for p, m in zip(parts, timing["part_ms"])]
    del parts
    torch.cuda.empty_cache()

    # ---- the rest of the metric: per_config (weak) and strong_1e8 (strong), same run, same JSON line --------------
    per_config, strong_1e8 = None, None
    if args.workload == "all" and not args.no_extra:
        per_config = []
        for key in ("bc7", "bc6h"):
            per_config.append(extra_leg(B, key, max(5, min(args.steps, 20)), args, with_e2e=(key == "bc7"), e2e_steps=min(e2e_steps, 5)))
        strong_1e8 = {}
        for key, name in (("bc1_1e8", "bc1"), ("mixed", "mixed")):
            strong_1e8[name] = extra_leg(B, key, max(5, min(args.steps, 10)), args, with_e2e=False, e2e_steps=0)

    if rank == 0:
        line = {
            "metric": "4x4 blocks/s", "value": timing["value"], "unit": "blocks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": timing["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(head_key, world), "per_part": per_part, "roofline": roof,
            "e2e": e2e, "cpu_baseline": cpu_baseline, "gpu_launches": timing["launches"], "clocks": clocks,
            "parity": parity, "hbm_gbs_per_gpu": step_bytes / (timing["ms_per_step"] / 1e3) / 1e9,
            "per_config": per_config, "strong_1e8": strong_1e8, "library": D.load().detexb200Version().decode(),
        }
        if gather is not None:
            line["gather"] = gather
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_leg(B, key, steps, args, with_e2e, e2e_steps):
    """One more configuration of the metric, measured like the headline leg."""
    wl = WORKLOADS[key]
    torch = B.torch
    parts = B.build_parts(wl)
    timing = B.time_device(parts, steps, 3)
    leg = {"workload": wl["name"], "scaling": "strong" if wl.get("strong") else "weak", "value": timing["value"],
           "unit": "blocks/s", "ms": timing["ms_per_step"], "steps": steps, "gpu_launches": timing["launches"],
           "blocks_per_step_per_gpu": int(sum(p["n"] for p in parts)), "job_blocks_per_step": int(timing["job_blocks"]),
           "roofline": B.roofline(parts, timing),
           "hbm_gbs_per_gpu": sum(p["n"] * ALGO_BYTES[p["fmt"]] for p in parts) / (timing["ms_per_step"] / 1e3) / 1e9}
    if timing["part_ms"] is not None:
        leg["per_part"] = [{"format": p["fmt"], "blocks": int(p["n"]), "ms": m, "gblocks_s": p["n"] / m / 1e6,
                            "gbs": p["n"] * ALGO_BYTES[p["fmt"]] / m / 1e6} for p, m in zip(parts, timing["part_ms"])]
    if with_e2e:
        leg["e2e"], views = B.e2e(parts, e2e_steps)
        del views
    if not args.no_cpu:
        leg["parity"], cpu_rate, threads = B.parity(parts, full_compare=not wl.get("strong"))
        if B.world == 1 and cpu_rate:
            leg["cpu_baseline"] = {"value": cpu_rate, "unit": "blocks/s", "cores": threads, "kind": leg["parity"]["against"],
                                   "sample": "the whole leg, decoded once for the parity check (single pass, cold output pages)"}
    del parts
    torch.cuda.empty_cache()
    return leg


def reference_child(key):
    """bench.py --impl reference --workload key in a child process -> its cpu_baseline dict (None when the workload has
    no reference arm of its own shape or the child fails)"""
    wl = WORKLOADS[key]
    if wl.get("strong") or wl["parts"][0][3] != "linear":
        return None
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", key,
                            "--steps", "10", "--warmup", "3"], capture_output=True, text=True, timeout=300)
        ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
        base = dict(ref["cpu_baseline"])
        base["how"] = "bench.py --impl reference --steps 10 --warmup 3, run as a child process before this run touched the GPU"
        return base
    except Exception:
        return None


def cpu_baseline_leg(B, p0, pre_ref=None):
    """The reference's CPU decode of the headline texture on this box's host cores (warm output buffer, persistent
    threads, mean over the timed passes), a single-thread figure (the reference as shipped), and the pixels of one pass
    for the parity check."""
    from oracle import harness as H
    D = B.D
    cpu_lib, kind = H.best_cpu_decoder()
    threads = B.cores
    fmt = p0["fmt"]
    n_s, w_s, h_s = p0["n"], p0["w"], p0["h"]
    linear = p0["layout"] == D.LAYOUT_LINEAR
    if not linear:      # flat arrays: bounded sample of the first 2^22 blocks as a 2048-block-wide texture
        n_s = min(p0["n"], 1 << 22)
        n_s -= n_s % 2048
        w_s, h_s = 8192, n_s // 2048 * 4
    hb = p0["sets"][0][0][: n_s * p0["bs"]].cpu().numpy()
    buf = np.zeros((w_s * h_s if linear else n_s * 16) * p0["ps"], dtype=np.uint8)
    # One pass here gives the pixels for the parity check.  The number itself comes from the --impl reference arm run as a
    # child process at the start of the run (pre_ref; same procedure and conditions as the arm the driver runs: inside this
    # process, after gigabytes of pinned and device allocations, the same decode measured anywhere between 0.55x and 1.2x).
    want, ok, s_once = H.cpu_decode(cpu_lib, fmt, hb, w_s, h_s, linear=linear, threads=threads, out=buf)
    small = np.zeros(buf.size // 16, dtype=np.uint8)
    _, _, s1 = H.cpu_decode(cpu_lib, fmt, hb[: (n_s // 16) * p0["bs"]], w_s, h_s // 16, linear=linear, threads=1, out=small)
    _, _, s1 = H.cpu_decode(cpu_lib, fmt, hb[: (n_s // 16) * p0["bs"]], w_s, h_s // 16, linear=linear, threads=1, out=small)
    base = pre_ref
    if base is None:
        base = {"value": n_s / s_once, "unit": "blocks/s", "cores": threads, "kind": kind,
                "sample": "%s %dx%d, %d blocks, one pass in this process" % (fmt, w_s, h_s, n_s)}
    base["in_process_single_pass_value"] = n_s / s_once
    base["single_thread_value"] = (n_s // 16) / s1
    return base, (want if linear else None)


def gather_leg(B, parts, timing, args):
    """optional final-image gather, outside the decode numbers: every rank's decoded slab of the first
    part all-gathered into the whole image on every rank, NCCL over NVLink / NVSwitch; and the same gather fused into
    the decode (kernels store straight into rank 0's image through a peer mapping)"""
    from detex_b200 import sharding as S
    torch, dist, D = B.torch, B.dist, B.D
    world, rank, dev, stream = B.world, B.rank, B.dev, B.stream
    slab = parts[0]["sets"][0][1]
    image = torch.empty(slab.numel() * world, dtype=torch.uint8, device=dev)
    for _ in range(3):
        S.gather_image(slab, out=image)
    torch.cuda.synchronize()
    dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g_steps = 10
    g0.record(stream)
    for _ in range(g_steps):
        S.gather_image(slab, out=image)
    g1.record(stream)
    torch.cuda.synchronize()
    g_ms = B.reduce_max(g0.elapsed_time(g1) / g_steps)
    gather = {"op": "all_gather (NCCL) of the decoded slabs into the whole image on every rank",
              "slab_bytes_per_rank": int(slab.numel()), "image_bytes": int(image.numel()), "ms": g_ms,
              "received_gbs_per_gpu": slab.numel() * (world - 1) / g_ms / 1e6,
              "decode_ms_per_step": timing["ms_per_step"]}
    try:
        p = parts[0]
        peer = S.PeerImage(slab.numel() * world, dev)
        dst = peer.slab_pointer(0, rank * slab.numel())

        def decode_to_root():
            D.decode_device(p["tf"], p["sets"][0][0].data_ptr(), p["w"], p["h"], dst, layout=p["layout"],
                            d_invalid=B.flag.data_ptr(), stream=B.sp, wib=p["wib"], hib=p["hib"])
        for _ in range(3):
            decode_to_root()
        torch.cuda.synchronize()
        dist.barrier()
        g0.record(stream)
        for _ in range(g_steps):
            decode_to_root()
        g1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        f_ms = B.reduce_max(g0.elapsed_time(g1) / g_steps)
        same = B.reduce_all(rank != 0 or torch.equal(peer.local, image))
        gather["fused_peer_store"] = {
            "op": "decode kernels store their slab into rank 0's image through peer-mapped memory",
            "ms": f_ms, "image_equals_all_gather": same,
            "root_received_gbs": slab.numel() * (world - 1) / f_ms / 1e6}
    except Exception as exc:        # symmetric memory unavailable on this build / topology
        gather["fused_peer_store"] = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    return gather


if __name__ == "__main__":
    main()
